"""Shared helpers of the parity tests: oracle <-> device plumbing and tolerances."""
import numpy as np
import torch

import smct_b200  # noqa: F401  (registers the package)
import smct_b200.functional as F
from oracle import smct_oracle as O

# Parity bar (BASELINE.json north_star): ancestor indices bit-exact; everything floating point within
# 1e-4 relative in fp32.  "Relative" is taken against the tensor's scale so that entries near zero are
# not held to an impossible standard: |a-b| <= RTOL*|b| + RTOL*max|b|*ATOL_FRAC.
RTOL = 1e-4
ATOL_FRAC = 0.1


# worst error per tensor name over the whole session, in units of the tolerance and as |err| / max|ref|
# (printed by tests/conftest.py:pytest_terminal_summary so that the GPU log shows how much margin the bar has)
PARITY_LOG = {}


def _log_parity(name, err, tol, scale):
    key = name.split(":")[-1].split("[")[0].strip()
    r_tol = float((err / tol).max()) if err.size else 0.0
    r_scale = float(err.max() / (scale + 1e-30)) if err.size else 0.0
    old = PARITY_LOG.get(key, (0.0, 0.0, 0))
    PARITY_LOG[key] = (max(old[0], r_tol), max(old[1], r_scale), old[2] + 1)


def assert_close(name, got, ref, rtol=RTOL, atol_frac=ATOL_FRAC):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (name, got.shape, ref.shape)
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    tol = rtol * np.abs(ref) + rtol * atol_frac * scale + 1e-30
    err = np.abs(got - ref)
    _log_parity(name, err, tol, scale)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err / tol), err.shape)
        raise AssertionError("%s: %d/%d outside tolerance; worst at %s got=%r ref=%r (scale %.3g)" % (
            name, int(bad.sum()), bad.size, i, got[i], ref[i], scale))
    return float((err / (np.abs(ref) + atol_frac * scale + 1e-30)).max()) if ref.size else 0.0


def dev_params(p, device="cuda"):
    return {k: torch.as_tensor(np.asarray(v, np.float32).reshape(-1) if np.ndim(v) == 0 else np.asarray(v, np.float32),
                               device=device).contiguous() for k, v in p.items()}


def logvar_dim(p):
    return int(np.asarray(p["logvar_k"]).size)


from tests.util_cpu import make_case, screen_uniforms  # noqa: E402,F401


def run_device(cfg, p, x_in, y, eps=None, u=None, i_forced=None, seed=0, want=F.OPTIONAL_OUTPUTS, algo="staged",
               b_global0=0):
    shape = F.shape_from_config(cfg, logvar_dim=logvar_dim(p), algo=algo, b_global0=b_global0)
    dp = dev_params(p)
    tx = torch.as_tensor(x_in, device="cuda").contiguous()
    ty = torch.as_tensor(y, device="cuda").contiguous()
    te = None if eps is None else torch.as_tensor(eps, device="cuda").contiguous()
    tu = None if u is None else torch.as_tensor(u, device="cuda").contiguous()
    ti = None if i_forced is None else torch.as_tensor(np.asarray(i_forced, np.int32), device="cuda").contiguous()
    out = F.forward(shape, dp, tx, ty, te, tu, ti, seed=seed, want=tuple(want))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def compare_forward(cfg, p, dev, ora, y, check_indices=True, loss_variant="checkout"):
    worst = {}
    if check_indices and cfg.noise:
        assert np.array_equal(dev["i_out"].astype(np.int64), ora["i_t"]), "ancestor indices differ"
    for k in ("pred", "pred_resampl", "K", "V", "R", "w", "noise_q", "noise_z", "attn"):
        if k in dev:
            worst[k] = assert_close(k, dev[k], ora[k])
    if "noise_K" in dev:
        worst["noise_K"] = assert_close("noise_K", dev["noise_K"], ora["noise_K_resampled"])
    if "noise_V" in dev:
        worst["noise_V"] = assert_close("noise_V", dev["noise_V"], ora["noise_V_resampled"])
    if "logw" in dev and cfg.noise:
        worst["logw"] = assert_close("logw", dev["logw"], ora["logw"])
    if "loss_sums" in dev and cfg.noise:
        ls = dev["loss_sums"]
        names = ["noise_K_resampled", "noise_q", "noise_V_resampled", "noise_z"]
        lvs = ["logvar_k", "logvar_q", "logvar_v", "logvar_z"]
        for i, (n, lv) in enumerate(zip(names, lvs)):
            ref = (np.square(ora[n].astype(np.float64)) / np.exp(np.broadcast_to(np.asarray(p[lv], np.float64), (cfg.D,)))).sum()
            assert abs(ls[i] - ref) <= 2e-4 * abs(ref) + 1e-12, "loss_sums[%d] %r vs %r" % (i, ls[i], ref)
        if cfg.weights == "gaussian":
            yy = np.asarray(y, np.float64).reshape(cfg.B, 1, cfg.S, cfg.F_y)
            ref = np.square(yy - ora["pred_resampl"].astype(np.float64)).sum()
            assert abs(ls[4] - ref) <= 2e-4 * abs(ref) + 1e-12, "loss_sums[4] %r vs %r" % (ls[4], ref)
    return worst
