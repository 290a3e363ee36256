"""Mirror of src/eval/inference_functions.py (regression-era forecasting helpers).

inference_onestep is the reference's function as it is.  inference_multistep is the regression-era multi-step
forecast ("stop resampling at past_len, roll the model forward feeding back the last prediction plus
N(0, Sigma_obs)" — recovered from the shipped .pyc; the checkout's version is NLP-ised and re-runs the whole
forward for every future step, O(S^2)).  Here the roll-out is incremental: one smct_cell_step per future
step on the particles' own K/V cache, O(S).
"""
import math

import numpy as np
import torch

from .. import functional as F
from ..models.classic_layers import _device


def split_input_target(data):
    """inference_functions.py:4-7."""
    return data[:, :, :-1, :], data[:, :, 1:, :]


def inference_onestep(smc_transformer, inputs, targets, save_path, past_len=40):
    """inference_functions.py:9-18: forward with resampling stopped at past_len; particle-mean predictions."""
    smc_transformer.cell.add_stop_resampling(past_len)
    (preds, _), (K, V, _), _ = smc_transformer(inputs, targets)
    mean_preds = preds.mean(dim=1)
    preds_future = preds[:, :, past_len:, :]
    if save_path is not None:
        np.save(save_path, mean_preds.cpu().numpy())
    return preds_future, mean_preds


def inference_multistep(smc_transformer, inputs, targets, past_len=40, future_len=20, seed=0, eps=None, u=None,
                        obs_noise=None):
    """Multi-step forecast for the regression SMC-T.  inputs/targets (B,1,past_len,F) observed past (F_x == F_y).
    Returns (preds_future (B,P,future_len,F_y), mean over particles (B,future_len,F_y)).  Each future step feeds
    x_{t+1}[p] = pred_t[p] + sqrt(Sigma_obs) * N(0,1) back as a per-particle input; resampling stops at past_len.
    Tests inject the draws: eps (4,S,B,P,D), u (S,B,P), obs_noise (S,B,P,F_y) with S = past_len + future_len."""
    m = smc_transformer
    c = m.cell
    dev = _device()
    inputs = torch.as_tensor(inputs).to(dev, torch.float32)
    targets = torch.as_tensor(targets).to(dev, torch.float32)
    B, _, S0, Fx = inputs.shape
    Fy = targets.shape[-1]
    if S0 < past_len:
        raise ValueError("need at least past_len observed steps")
    if Fx != Fy:
        raise ValueError("multi-step forecasting feeds predictions back as inputs: F_x must equal F_y")
    P, D, S = c.num_particles, m.d_model, past_len + future_len
    params = m._params("regression", Fx)
    shape = F.make_shape(B, P, S, D, H=m.num_heads, dff=m.dff, F_x=Fx, F_y=Fy, full_model=m.full_model,
                         attn_window=c.attention_smc.attn_window, len_resampling=past_len, weights="gaussian",
                         noise_z_mode=m.noise_z_mode or "pyc", input_mode="linear", noise=c.noise,
                         logvar_dim=c.attention_smc.logvar_dim(), sigma_obs=c.Sigma_obs)
    K = torch.zeros((B, P, S, D), device=dev)
    V = torch.zeros_like(K)
    R = torch.zeros_like(K)
    # encoded past inputs, one row per (b,s) (shared by the particles)
    enc_shape = F.make_shape(B, 1, past_len, D, F_x=Fx, F_y=Fy, input_mode="linear", dff=m.dff)
    X_past = F.encode(enc_shape, params, inputs[:, 0, :past_len].contiguous())                  # (B,past_len,D)
    preds = []
    x_next = None
    cvt = lambda a, dt: None if a is None else torch.as_tensor(np.asarray(a), device=dev).to(dt).contiguous()
    eps, u, obs_noise = cvt(eps, torch.float32), cvt(u, torch.float64), cvt(obs_noise, torch.float32)
    for t in range(S):
        inj = dict(eps4=None if eps is None else eps[:, t].contiguous(), u=None if u is None else u[t].contiguous())
        if t < past_len:
            st = F.cell_step(shape, params, t, X_past[:, t].contiguous(), targets[:, 0, t].contiguous(), K, V, R,
                             seed=seed, want_attn=False, **inj)
        else:
            encp = F.make_shape(B, 1, P, D, F_x=Fx, F_y=Fy, input_mode="linear", dff=m.dff)     # rows = particles
            X_t = F.encode(encp, params, x_next.contiguous())                                   # (B,P,D)
            dummy_y = torch.zeros((B, Fy), device=dev)
            st = F.cell_step(shape, params, t, X_t, dummy_y, K, V, R, seed=seed, want_attn=False, **inj)
            preds.append(st["pred"])
        # next input of every particle: its own prediction plus observation noise
        # observation noise from the library's own generator (stream 0 of a derived seed; (B,P,F_y) normals)
        n_obs = F.fill_noise(int(seed) ^ 0x0B5E, 0, t, 0, B, P, Fy, device=dev) if obs_noise is None else obs_noise[t]
        x_next = st["pred"] + math.sqrt(c.Sigma_obs) * n_obs
    preds_future = torch.stack(preds, dim=2)                                                     # (B,P,future_len,F_y)
    return preds_future, preds_future.mean(dim=1)


def get_empirical_distrib(mean_NP, sigma_obs, N_est, P, rng=None):
    """inference_functions.py:57-66: N_est draws from the particle mixture with N(0, sigma_obs) observation noise."""
    rng = rng or np.random.default_rng()
    mean_NP = np.asarray(mean_NP.cpu() if isinstance(mean_NP, torch.Tensor) else mean_NP)
    out = []
    for _ in range(N_est):
        ind_p = rng.integers(0, P)
        sampled_mean = mean_NP[:, ind_p, :]
        out.append(sampled_mean + sigma_obs * rng.standard_normal(sampled_mean.shape))
    return np.stack(out, axis=0)


def get_distrib_all_timesteps(preds, sigma_obs, P, save_path_distrib, N_est=10, len_future=20, rng=None):
    """inference_functions.py:43-55."""
    preds = np.asarray(preds.cpu() if isinstance(preds, torch.Tensor) else preds)
    d = np.stack([get_empirical_distrib(preds[:, :, t, :], sigma_obs, N_est, P, rng) for t in range(len_future)], axis=0)
    d = np.transpose(d, axes=[2, 1, 0, 3])
    if save_path_distrib is not None:
        np.save(save_path_distrib, d)
    return d
