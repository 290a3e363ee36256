"""Backward pass (smct_backward) against torch autograd through the eager restatement of the reference
(oracle/ref_eager_torch.loss_and_grads, fp64), on the same injected draws and the device's own ancestors."""
import numpy as np
import pytest
import torch

import smct_b200.functional as F
from oracle import ref_eager_torch as E
from oracle import smct_oracle as O
from tests import util

pytestmark = pytest.mark.gpu

BWD_CASES = {
    "c1": O.Config(B=4, P=10, S=24, D=8, dff=8, noise_z_mode="pyc"),
    "c3": O.Config(B=2, P=60, S=12, D=32, dff=32, F_x=5, F_y=5, noise_z_mode="pyc"),
    "d64_p70": O.Config(B=2, P=70, S=10, D=64, dff=256, F_x=2, F_y=2, noise_z_mode="pyc"),
    "window_stop": O.Config(B=2, P=33, S=9, D=16, dff=24, attn_window=3, len_resampling=5, noise_z_mode="pyc"),
    "d128": O.Config(B=1, P=20, S=6, D=128, dff=64, noise_z_mode="pyc"),
    # the nlp checkout's observation model: Embedding input, softmax(pred)[y] weights, sparse cross-entropy
    # (with the post-projection noise_z of the regression era: the checkout's z - z_ quirk is forward-only)
    "classification": O.Config(B=3, P=12, S=9, D=16, dff=16, F_x=30, F_y=30, weights="classification",
                               input_mode="embedding", noise_z_mode="pyc"),
}


def _run(cfg, variant, algo="staged", seed=0, multi=False):
    p, x_in, y, eps, u = util.make_case(cfg, seed=seed, sigma2=0.5, random_bias=True, multi=multi)
    shape = F.shape_from_config(cfg, logvar_dim=util.logvar_dim(p), algo=algo)
    dp = util.dev_params(p)
    tx, ty = torch.as_tensor(x_in, device="cuda"), torch.as_tensor(y, device="cuda")
    te, tu = torch.as_tensor(eps, device="cuda"), torch.as_tensor(u, device="cuda")
    want = ("i_out", "loss_sums") + (("noise_K", "noise_q", "noise_V", "noise_z") if multi else ())
    fwd = F.forward(shape, dp, tx, ty, te, tu, want=want)
    grads = F.backward(shape, dp, tx, ty, fwd, eps=te, loss_variant=variant)
    torch.cuda.synchronize()
    i_out = fwd["i_out"].cpu().numpy()
    loss, ref = E.loss_and_grads(cfg, p, x_in, y, eps, i_out, variant=variant)
    return {k: v.cpu().numpy() for k, v in grads.items()}, ref, p


def _compare(name, got, ref, rtol=1e-4):
    """Per-parameter tolerance rtol*max|grad_k| plus a floor of 1e-6 of the largest gradient of the model:
    d/d bk is exactly zero (softmax is invariant to a shift of all keys, and K - wk(X) cancels the bias), so its
    computed value is pure fp32 cancellation noise."""
    gscale = max(np.abs(np.asarray(v)).max() for v in ref.values())
    for k in sorted(ref):
        r = np.asarray(ref[k], np.float64).reshape(-1)
        g = np.asarray(got[k], np.float64).reshape(-1)
        scale = np.abs(r).max()
        err = np.abs(g - r).max()
        assert err <= rtol * scale + 1e-6 * gscale, "%s: grad %s max err %.3g (scale %.3g)" % (name, k, err, scale)


@pytest.mark.parametrize("variant", ["pyc", "checkout"])
@pytest.mark.parametrize("name", sorted(BWD_CASES))
def test_backward_matches_autograd(name, variant):
    cfg = BWD_CASES[name]
    got, ref, p = _run(cfg, variant)
    _compare(name, got, ref)


@pytest.mark.parametrize("variant", ["pyc", "checkout"])
@pytest.mark.parametrize("name", ["c3", "window_stop", "classification", "d64_p70"])
def test_backward_per_dimension_logvars(name, variant):
    """noise_dim = "multi" (self_attention_SMC.py:57-61,98-107): per-dimension noise scales in the replayed forward, per-
    dimension 1/sigma^2 in the K / V loss terms, and (D,) log-variance gradients from the forward's recorded noises."""
    cfg = BWD_CASES[name]
    got, ref, p = _run(cfg, variant, multi=True, algo="fused" if name == "d64_p70" else "staged")
    assert got["logvar_k"].shape == (cfg.D,) and np.abs(ref["logvar_k"]).max() > 0
    _compare(name + ":multi", got, ref)


def test_backward_from_fused_forward_and_chunk_accumulation():
    """Backward is driven by forward OUTPUTS only: it must work after the fused (tensor-core) forward, and
    gradients of two batch shards accumulated with the global inv_n must equal the unsharded gradients."""
    cfg = O.Config(B=4, P=64, S=8, D=64, dff=256, noise_z_mode="pyc")
    p, x_in, y, eps, u = util.make_case(cfg, sigma2=0.5)
    dp = util.dev_params(p)
    tx, ty = torch.as_tensor(x_in, device="cuda"), torch.as_tensor(y, device="cuda")
    te, tu = torch.as_tensor(eps, device="cuda"), torch.as_tensor(u, device="cuda")
    shape = F.shape_from_config(cfg, logvar_dim=1, algo="fused")
    fwd = F.forward(shape, dp, tx, ty, te, tu, want=("i_out", "loss_sums"))
    full = F.backward(shape, dp, tx, ty, fwd, eps=te)
    inv_n = 1.0 / (cfg.B * cfg.P * cfg.S)
    acc = None
    for b0 in (0, 2):
        sc = O.Config(B=2, P=cfg.P, S=cfg.S, D=cfg.D, dff=cfg.dff, noise_z_mode="pyc")
        sh = F.shape_from_config(sc, logvar_dim=1, algo="fused", b_global0=b0)
        e2 = te[:, :, b0:b0 + 2].contiguous()
        u2 = tu[:, b0:b0 + 2].contiguous()
        f2 = F.forward(sh, dp, tx[b0:b0 + 2].contiguous(), ty[b0:b0 + 2].contiguous(), e2, u2, want=("i_out", "loss_sums"))
        assert torch.equal(f2["i_out"], fwd["i_out"][:, b0:b0 + 2])
        acc = F.backward(sh, dp, tx[b0:b0 + 2].contiguous(), ty[b0:b0 + 2].contiguous(), f2, eps=e2, inv_n=inv_n, grads=acc)
    i_out = fwd["i_out"].cpu().numpy()
    _, ref = E.loss_and_grads(cfg, p, x_in, y, eps, i_out)
    _compare("full", {k: v.cpu().numpy() for k, v in full.items()}, ref)
    _compare("accumulated shards", {k: v.cpu().numpy() for k, v in acc.items()}, ref)


def test_adam_step_matches_keras_formula():
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(1000).astype(np.float32)
    g = rng.standard_normal(1000).astype(np.float32)
    p = torch.as_tensor(p0.copy(), device="cuda")
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    b1, b2, eps, lr = 0.9, 0.98, 1e-9, 1e-3
    pm, pv, pp = np.zeros(1000), np.zeros(1000), p0.astype(np.float64)
    for t in range(1, 4):
        lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        F.adam_step(p, torch.as_tensor(g, device="cuda"), m, v, lr_t, b1, b2, eps)
        pm = b1 * pm + (1 - b1) * g; pv = b2 * pv + (1 - b2) * g.astype(np.float64) ** 2
        pp = pp - lr_t * pm / (np.sqrt(pv) + eps)
    assert np.allclose(p.cpu().numpy(), pp, rtol=1e-5, atol=1e-6)


def test_training_reduces_the_loss():
    """train_step_SMC_T with the mirrored Adam: a few hundred steps on a small AR(1) batch with FIXED draws
    (model.seed) must lower the SMC loss — gradients, optimizer and parameter plumbing end to end."""
    from smct_b200.models.SMC_Transformer.SMC_Transformer import SMC_Transformer
    from smct_b200.train.optimizers import Adam
    from smct_b200.train.train_step_functions import train_step_SMC_T
    from smct_b200 import utils
    x, y = utils.synthetic_ar1(16, 24, 1, alpha=0.8, var=0.5)
    inputs, targets = torch.from_numpy(x).unsqueeze(1), torch.from_numpy(y).unsqueeze(1)
    model = SMC_Transformer(d_model=8, output_size=1, seq_len=24, full_model=True, dff=8)
    model.cell.add_SMC_parameters(dict_sigmas=dict(k=0.5, q=0.5, v=0.5, z=0.5), num_particles=10, sigma_obs=0.5)
    model.seed = 1234
    opt = Adam(learning_rate=5e-3, beta_1=0.9, beta_2=0.98, epsilon=1e-9)
    first = float(train_step_SMC_T(inputs, targets, model, opt, it=1)[0])
    for it in range(2, 150):
        loss, metric = train_step_SMC_T(inputs, targets, model, opt, it=it)
    assert float(loss) < 0.9 * first, (first, float(loss))
