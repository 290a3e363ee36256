"""The CUDA path against fixtures written by the reference's OWN Python sources (tests/golden/ref_*.npz,
generator oracle/gen_golden_from_reference.py) — through the C-ABI and through the Python mirror of the
reference call surface (SMC_Transformer, cell, compute_SMC_loss)."""
import os

import numpy as np
import pytest
import torch

import smct_b200.functional as F
from tests import ref_golden, util

pytestmark = pytest.mark.gpu


def _w_ref_form(cfg, dev):
    if cfg.weights == "classification":
        return dev["w"]
    lw = dev["logw"].astype(np.float64)                      # (S,B,P) raw log-weights
    m = lw.max(-1, keepdims=True)
    return (lw - (m + np.log(np.exp(lw - m).sum(-1, keepdims=True)))).transpose(1, 2, 0)


@pytest.mark.parametrize("path", ref_golden.ref_files())
def test_cabi_forward_matches_reference_sources(path):
    g, cfg, p, x_in, y = ref_golden.load(path)
    algos = ["staged"] + (["fused"] if (cfg.D == 64 and cfg.dff == 256 and cfg.weights == "gaussian") else [])
    for algo in algos:
        dev = util.run_device(cfg, p, x_in, y, g["eps"], g["u"], algo=algo,
                              want=("i_out", "logw", "noise_q", "noise_z", "noise_K", "noise_V", "loss_sums"))
        dev["w_as_returned_by_reference"] = _w_ref_form(cfg, dev)
        ref_golden.compare(os.path.basename(path) + ":" + algo, dev, g, util.assert_close)


def _build_mirror(g, cfg, p):
    from smct_b200.models.SMC_Transformer.SMC_Transformer import SMC_Transformer
    model = SMC_Transformer(d_model=cfg.D, output_size=cfg.F_y, seq_len=cfg.S, full_model=True, dff=cfg.dff,
                            num_layers=1, num_heads=cfg.H, attn_window=cfg.attn_window)
    lv = p["logvar_k"]
    sig = lambda n: [float(v) for v in np.exp(p["logvar_" + n])] if lv.size > 1 else float(np.exp(p["logvar_" + n]))
    model.cell.add_SMC_parameters(dict_sigmas={n: sig(n) for n in ("k", "q", "v", "z")}, num_particles=cfg.P,
                                  sigma_obs=cfg.sigma_obs)
    if cfg.len_resampling is not None:
        model.cell.add_stop_resampling(cfg.len_resampling)
    model.noise_z_mode = "checkout"   # the fixtures were written by the checkout's cell code (noise_z = z - z_)
    att, c = model.cell.attention_smc, model.cell
    for layer, (W, b) in ((att.wq, ("Wq", "bq")), (att.wk, ("Wk", "bk")), (att.wv, ("Wv", "bv")), (att.dense, ("Wz", "bz")),
                          (c.ffn.dense1, ("W1", "b1")), (c.ffn.dense2, ("W2", "b2"))):
        layer.kernel.assign(p[W]); layer.bias.assign(p[b])
    c.output_layer.kernel.assign(p["Wo"])
    c.layernorm1.gamma.assign(p["ln1_g"]); c.layernorm1.beta.assign(p["ln1_b"])
    c.layernorm2.gamma.assign(p["ln2_g"]); c.layernorm2.beta.assign(p["ln2_b"])
    return model


@pytest.mark.parametrize("path", ref_golden.ref_files())
def test_python_mirror_matches_reference_sources(path):
    """Same call as the reference: model(inputs (B,1,S,F), targets) -> ((pred, pred_resampl), (K,V,R), w);
    model.compute_SMC_loss(inputs, targets, predictions) -> (total, classic)."""
    g, cfg, p, x_in, y = ref_golden.load(path)
    model = _build_mirror(g, cfg, p)
    inputs, targets = torch.from_numpy(g["x_in"]), torch.from_numpy(g["y"])      # the reference's own (B,1,S,F) tensors
    if cfg.weights == "classification":
        inputs, targets = inputs.to(torch.int32), targets.to(torch.int32)
        model.embedding.build(np.random.default_rng(0))
        model.embedding.embeddings.assign(p["W_in"])
    else:
        model.projection_layer_ts.build(cfg.F_x, np.random.default_rng(0))
        model.projection_layer_ts.kernel.assign(p["W_in"]); model.projection_layer_ts.bias.assign(p["b_in"])
    (pred, pred_r), (K, V, R), w = model(inputs, targets, eps=g["eps"], u=g["u"])
    assert tuple(pred.shape) == (cfg.B, cfg.P, cfg.S, cfg.F_y) and tuple(K.shape) == (cfg.B, cfg.P, cfg.S, cfg.D)
    assert tuple(w.shape) == (cfg.B, cfg.P, cfg.S)
    got = dict(pred=pred.cpu().numpy(), pred_resampl=pred_r.cpu().numpy(), K=K.cpu().numpy(), V=V.cpu().numpy(),
               R=R.cpu().numpy(), noise_q=model.noise_q.cpu().numpy(), noise_z=model.noise_z.cpu().numpy(),
               noise_K=model.noise_K_resampled.cpu().numpy())
    wn = w.cpu().numpy()
    got["w_as_returned_by_reference"] = wn if cfg.weights == "classification" else np.log(np.maximum(wn.astype(np.float64), 1e-300))
    if cfg.weights == "classification":
        ref_golden.compare(os.path.basename(path), got, g, util.assert_close)
    else:
        # log of a tiny fp32 weight loses relative precision: compare weights in probability space instead
        util.assert_close("w", wn, np.exp(g["w"].astype(np.float64)))
        got["w_as_returned_by_reference"] = g["w"]
        ref_golden.compare(os.path.basename(path), got, g, util.assert_close)
    total, classic = model.compute_SMC_loss(inputs=inputs, targets=targets, predictions=pred_r)
    assert np.isclose(float(total), float(g["loss_total"]), rtol=2e-4), (float(total), float(g["loss_total"]))
    assert np.isclose(float(classic), float(g["loss_classic"]), rtol=2e-4)
    assert model.cell.dec_timestep == 0 and model.cell.cell_count == 0


def test_mirror_cell_and_layer_stepwise():
    """Drive the mirrored SMC_Transf_Cell the way tf.keras.backend.rnn drives the reference cell (one dummy
    call at t=0, then S steps) and the mirrored Self_Attention_SMC layer once; compare with the forward."""
    from smct_b200.models.SMC_Transformer.SMC_TransformerCell import NestedInput
    from smct_b200.models.SMC_Transformer.transformer_utils import resample
    path = [f for f in ref_golden.ref_files() if "pyc_smoke" in f][0]
    g, cfg, p, x_in, y = ref_golden.load(path)
    model = _build_mirror(g, cfg, p)
    model.projection_layer_ts.build(cfg.F_x, np.random.default_rng(0))
    model.projection_layer_ts.kernel.assign(p["W_in"]); model.projection_layer_ts.bias.assign(p["b_in"])
    inputs, targets = torch.from_numpy(g["x_in"]), torch.from_numpy(g["y"])
    X = model.get_encoded_input(inputs)                                           # (B,P,S,D)
    assert tuple(X.shape) == (cfg.B, cfg.P, cfg.S, cfg.D)
    yt = targets.expand(cfg.B, cfg.P, cfg.S, cfg.F_y).cuda()
    z = torch.zeros((cfg.B, cfg.P, cfg.S, cfg.D), device="cuda")
    states = (z, z.clone(), z.clone())
    cell = model.cell
    eps, u = g["eps"], g["u"]
    cell(NestedInput(x=X[:, :, 0].contiguous(), y=yt[:, :, 0].contiguous()), states, eps4=eps[:, 0], u=u[0])  # K.rnn dummy call
    assert cell.dec_timestep == 0 and cell.cell_count == 1
    for t in range(cfg.S):
        out, states = cell(NestedInput(x=X[:, :, t].contiguous(), y=yt[:, :, t].contiguous()), states, eps4=eps[:, t], u=u[t])
        assert len(out) == 4 and tuple(out[0].shape) == (cfg.B, cfg.P, 1, cfg.D) and tuple(out[1].shape) == (cfg.B, cfg.P, cfg.H, 1, cfg.S)
    assert cell.dec_timestep == cfg.S
    util.assert_close("K", states.K.cpu().numpy(), g["K"])
    util.assert_close("R", states.R.cpu().numpy(), g["R"])
    cell.dec_timestep = 0; cell.cell_count = 0
    # resample(params, i_t) mirror
    i_t = torch.randint(0, cfg.P, (cfg.B, cfg.P))
    got = resample(states.K, i_t).cpu().numpy()
    ref = np.take_along_axis(states.K.cpu().numpy(), i_t.numpy()[:, :, None, None], axis=1)
    assert np.array_equal(got, ref)
    with pytest.raises(IndexError):
        resample(states.K, i_t + cfg.P)
    # compute_w_regression mirror: normalised weights
    pred = torch.randn(cfg.B, cfg.P, 1, cfg.F_y)
    yy = torch.randn(cfg.B, 1, 1, cfg.F_y).expand(cfg.B, cfg.P, 1, cfg.F_y)
    wreg = cell.compute_w_regression(pred, yy).cpu().numpy()
    lw = -0.5 * ((yy - pred) ** 2).sum(-1).squeeze(-1).numpy().astype(np.float64) / cfg.sigma_obs
    ref = np.exp(lw - lw.max(1, keepdims=True)); ref /= ref.sum(1, keepdims=True)
    util.assert_close("w_regression", wreg, ref)


def test_train_step_mirror_forward_loss_and_em():
    from smct_b200.train.train_step_functions import train_step_SMC_T
    path = [f for f in ref_golden.ref_files() if "pyc_smoke" in f][0]
    g, cfg, p, x_in, y = ref_golden.load(path)
    model = _build_mirror(g, cfg, p)
    inputs, targets = torch.from_numpy(g["x_in"]), torch.from_numpy(g["y"])
    model.seed = 5
    loss, metric = train_step_SMC_T(inputs, targets, model, None, it=1)
    assert np.isfinite(float(loss)) and np.isfinite(float(metric))
    lv0 = float(model.cell.attention_smc.logvar_k.numpy())
    loss2, _ = train_step_SMC_T(inputs, targets, model, None, it=2, EM_param=0.6)
    assert float(model.cell.attention_smc.logvar_k.numpy()) != lv0          # EM moved the variance
    # a real optimisation step: gradients through smct_backward, Adam through smct_adam_step
    from smct_b200.train.optimizers import Adam, CustomSchedule
    model.noise_z_mode = None                                                # regression-era noise_z (needed by the backward)
    opt = Adam(CustomSchedule(cfg.D), beta_1=0.9, beta_2=0.98, epsilon=1e-9)
    w_before = model.cell.ffn.dense1.kernel.numpy().copy()
    model.seed = 9
    losses = [float(train_step_SMC_T(inputs, targets, model, opt, it=i + 3)[0]) for i in range(3)]
    assert np.all(np.isfinite(losses)) and not np.array_equal(model.cell.ffn.dense1.kernel.numpy(), w_before)


def test_smct_algo_wiring():
    from smct_b200.algos.run_SMC_T import SMCTAlgo, algos, default_args
    assert algos["smc_t"] is SMCTAlgo
    args = default_args(d_model=8, dff=8, particles=10, sigmas=0.5)
    algo = SMCTAlgo(dataset=None, args=args, output_size=1, seq_len=24)
    m = algo.smc_transformer
    assert m.cell.noise and m.cell.num_particles == 10 and m.cell.attention_smc.noise
    x = torch.randn(32, 1, 24, 1)
    (pred, pred_r), (K, V, R), w = m(x, x)
    assert tuple(pred.shape) == (32, 10, 24, 1) and tuple(K.shape) == (32, 10, 24, 8)
    assert np.allclose(w.sum(1).cpu().numpy(), 1.0, atol=1e-4)


def test_forecasting_helpers():
    """Regression-era inference helpers (src/eval/inference_functions.py): one-step with resampling stopped,
    incremental multi-step roll-out on the particles' K/V cache."""
    from smct_b200.eval.inference_functions import (get_distrib_all_timesteps, inference_multistep, inference_onestep,
                                                     split_input_target)
    from smct_b200.models.SMC_Transformer.SMC_Transformer import SMC_Transformer
    from smct_b200 import utils
    x, _ = utils.synthetic_ar1(6, 31, 1, alpha=0.8, var=0.5)
    data = torch.from_numpy(x).unsqueeze(1)                       # (B,1,31,1)
    inputs, targets = split_input_target(data)                    # (B,1,30,1)
    model = SMC_Transformer(d_model=8, output_size=1, seq_len=30, full_model=True, dff=8)
    model.cell.add_SMC_parameters(dict_sigmas=dict(k=0.1, q=0.1, v=0.1, z=0.1), num_particles=20, sigma_obs=0.5)
    preds_future, mean_preds = inference_onestep(model, inputs, targets, None, past_len=20)
    assert tuple(preds_future.shape) == (6, 20, 10, 1) and tuple(mean_preds.shape) == (6, 30, 1)
    assert model.cell.len_resampling == 20
    # after past_len the genealogy is frozen: K rows of slots >= past_len are the particles' own rows
    pf, mean = inference_multistep(model, inputs[:, :, :20], targets[:, :, :20], past_len=20, future_len=10, seed=3)
    assert tuple(pf.shape) == (6, 20, 10, 1) and tuple(mean.shape) == (6, 10, 1) and torch.isfinite(pf).all()
    pf2, _ = inference_multistep(model, inputs[:, :, :20], targets[:, :, :20], past_len=20, future_len=10, seed=3)
    assert torch.equal(pf, pf2)                                   # seeded roll-out is reproducible
    d = get_distrib_all_timesteps(pf, sigma_obs=0.5, P=20, save_path_distrib=None, N_est=7, len_future=10,
                                  rng=np.random.default_rng(0))
    assert d.shape == (6, 7, 10, 1)
