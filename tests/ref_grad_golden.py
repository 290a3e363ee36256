"""Loader for the gradient / EM fixtures written by the reference's own sources
(oracle/gen_golden_grads_from_reference.py -> tests/golden/ref_grad_*.npz)."""
import ast
import glob
import os

import numpy as np

from oracle.smct_oracle import Config

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def files():
    return sorted(glob.glob(os.path.join(GOLDEN, "ref_grad_*.npz")))


def load(path, noise_z_mode="pyc"):
    g = np.load(path, allow_pickle=True)
    kw = ast.literal_eval(str(g["meta"]))
    reg = kw.get("regression")
    p = {k[2:]: g[k] for k in g.files if k.startswith("p_")}
    grads = {k[2:]: g[k] for k in g.files if k.startswith("g_")}
    B, P, S, D = kw["B"], kw["P"], kw["S"], kw["D"]
    common = dict(B=B, P=P, S=S, D=D, H=1, dff=kw["dff"], attn_window=kw["attn_window"], noise_z_mode=noise_z_mode,
                  len_resampling=kw.get("len_resampling"))
    if reg is None:
        cfg = Config(F_x=kw["V"], F_y=kw["V"], weights="classification", input_mode="embedding", **common)
        x_in, y = g["x_in"].reshape(B, S).astype(np.int32), g["y"].reshape(B, S).astype(np.int32)
        variant = "checkout"
    else:
        cfg = Config(F_x=reg["F_x"], F_y=kw["V"], weights="gaussian", input_mode="linear", sigma_obs=reg["sigma_obs"], **common)
        x_in, y = g["x_in"].reshape(B, S, reg["F_x"]).astype(np.float32), g["y"].reshape(B, S, kw["V"]).astype(np.float32)
        variant = "pyc"
    return g, cfg, p, grads, x_in, y, variant


def compare_grads(name, got, ref, rtol=2e-4):
    """Per-parameter tolerance rtol*max|grad_k| plus 2e-6 of the largest gradient of the model (d/d bk is exactly zero —
    softmax shift invariance and K - wk(X) cancel the bias — so its value is fp32 cancellation noise on both sides)."""
    gscale = max(float(np.abs(v).max()) for v in ref.values())
    worst = {}
    for k in sorted(ref):
        r = np.asarray(ref[k], np.float64).reshape(-1)
        q = np.asarray(got[k], np.float64).reshape(-1)
        assert q.shape == r.shape, (name, k, q.shape, r.shape)
        scale = float(np.abs(r).max())
        err = float(np.abs(q - r).max())
        worst[k] = err / (scale + 1e-30)
        assert err <= rtol * scale + 2e-6 * gscale, "%s: grad %s max err %.3g (scale %.3g)" % (name, k, err, scale)
    return worst
