"""smct_backward, the EM update and the SMC loss against fixtures written by the REFERENCE'S OWN SOURCES
(tests/golden/ref_grad_*.npz: reference forward + compute_SMC_loss + tf.GradientTape over the TF shim, and
src/train/utils.py:EM; generator oracle/gen_golden_grads_from_reference.py)."""
import os

import numpy as np
import pytest
import torch

import smct_b200.functional as F
from tests import ref_grad_golden as RG
from tests import util

pytestmark = pytest.mark.gpu


def _algos(cfg):
    return ["staged"] + (["fused", "fused_v1"] if (cfg.D == 64 and cfg.dff == 256 and cfg.weights == "gaussian") else [])


@pytest.mark.parametrize("path", RG.files())
def test_backward_matches_reference_tape_gradient(path):
    """tape.gradient(loss, trainable_variables) of train_step_SMC_T (train_step_functions.py:43-63)."""
    g, cfg, p, ref, x_in, y, variant = RG.load(path)
    name = os.path.basename(path)
    dp = util.dev_params(p)
    tx, ty = torch.as_tensor(x_in, device="cuda"), torch.as_tensor(y, device="cuda")
    te, tu = torch.as_tensor(g["eps"], device="cuda"), torch.as_tensor(g["u"], device="cuda")
    classic_w = None
    if "mask" in g.files:   # attention_mask: weight of the classic-loss term of row (b,s) relative to the uniform mean
        m = g["mask"].reshape(cfg.B, cfg.S).astype(np.float32)
        classic_w = torch.as_tensor(m * (cfg.B * cfg.S / m.sum()), device="cuda").contiguous()
    for algo in _algos(cfg):
        ld = util.logvar_dim(p)    # 1, or d_model for the noise_dim = "multi" fixtures
        shape = F.shape_from_config(cfg, logvar_dim=ld, algo=algo)
        want = ("i_out", "loss_sums") + (("noise_K", "noise_q", "noise_V", "noise_z") if ld > 1 else ())
        fwd = F.forward(shape, dp, tx, ty, te, tu, want=want)
        # the forward is the one the reference ran: same resampled predictions
        util.assert_close(name + ":pred_resampl", fwd["pred_resampl"].cpu().numpy(), g["pred_resampl"])
        grads = F.backward(shape, dp, tx, ty, fwd, eps=te, loss_variant=variant, classic_w=classic_w)
        torch.cuda.synchronize()
        got = {k: v.cpu().numpy() for k, v in grads.items()}
        worst = RG.compare_grads(name + ":" + algo, got, ref)
        print("[ref-grad] %s %s: max rel err %s" % (name, algo, {k: "%.1e" % v for k, v in sorted(worst.items())}))


@pytest.mark.parametrize("path", RG.files())
def test_mirror_EM_matches_reference(path):
    """EM(smc_transformer, it, EM_param) — src/train/utils.py:22-51 — on the noises of the same forward."""
    from tests.test_gpu_reference_goldens import _build_mirror
    from smct_b200.train.utils import EM
    g, cfg, p, _, x_in, y, _ = RG.load(path, noise_z_mode="checkout")   # the reference recorded the checkout's noise_z
    if "lv_after" not in g.files:
        pytest.skip("the reference's EM mode is scalar-only (no EM record in the multi fixtures)")
    model = _build_mirror(g, cfg, p)
    for n in "kqvz":
        getattr(model.cell.attention_smc, "logvar_" + n).trainable = False
    inputs, targets = torch.from_numpy(g["x_in"]), torch.from_numpy(g["y"])
    if cfg.weights == "classification":
        inputs, targets = inputs.to(torch.int32), targets.to(torch.int32)
        model.embedding.build(np.random.default_rng(0))
        model.embedding.embeddings.assign(p["W_in"])
    else:
        model.projection_layer_ts.build(cfg.F_x, np.random.default_rng(0))
        model.projection_layer_ts.kernel.assign(p["W_in"]); model.projection_layer_ts.bias.assign(p["b_in"])
    model(inputs, targets, eps=g["eps"], u=g["u"])
    att = model.cell.attention_smc
    before = np.array([float(getattr(att, "logvar_" + n).numpy()) for n in "kqvz"])
    assert np.allclose(before, g["lv_before"], atol=1e-6)
    it, em_param = int(g["em"][0]), float(g["em"][1])
    EM(model, it=it, EM_param=em_param)
    after = np.array([float(getattr(att, "logvar_" + n).numpy()) for n in "kqvz"])
    assert np.allclose(after, g["lv_after"], rtol=2e-4, atol=2e-5), (after, g["lv_after"])


@pytest.mark.parametrize("path", RG.files())
def test_mirror_loss_matches_reference(path):
    from tests.test_gpu_reference_goldens import _build_mirror
    g, cfg, p, _, x_in, y, variant = RG.load(path, noise_z_mode="checkout")
    model = _build_mirror(g, cfg, p)
    inputs, targets = torch.from_numpy(g["x_in"]), torch.from_numpy(g["y"])
    if cfg.weights == "classification":
        inputs, targets = inputs.to(torch.int32), targets.to(torch.int32)
        model.embedding.build(np.random.default_rng(0))
        model.embedding.embeddings.assign(p["W_in"])
    else:
        model.projection_layer_ts.build(cfg.F_x, np.random.default_rng(0))
        model.projection_layer_ts.kernel.assign(p["W_in"]); model.projection_layer_ts.bias.assign(p["b_in"])
    mask = torch.from_numpy(g["mask"]) if "mask" in g.files else None
    (_, pred_r), _, _ = model(inputs, targets, attention_mask=mask, eps=g["eps"], u=g["u"])
    total, classic = model.compute_SMC_loss(inputs=inputs, targets=targets, predictions=pred_r, attention_mask=mask)
    assert np.isclose(float(total), float(g["loss_total"]), rtol=2e-4), (float(total), float(g["loss_total"]))
    assert np.isclose(float(classic), float(g["loss_classic"]), rtol=2e-4)
