"""Loader for the fixtures written by the reference's own sources (oracle/gen_golden_from_reference.py)."""
import ast
import glob
import os

import numpy as np

from oracle.smct_oracle import Config

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def ref_files():
    return sorted(f for f in glob.glob(os.path.join(GOLDEN, "ref_*.npz")) if not os.path.basename(f).startswith("ref_grad_"))


def load(path):
    g = np.load(path, allow_pickle=True)
    kw = ast.literal_eval(str(g["meta"]))
    reg = kw.get("regression")
    p = {k[2:]: g[k] for k in g.files if k.startswith("p_")}
    B, P, S, D = kw["B"], kw["P"], kw["S"], kw["D"]
    if reg is None:   # the checkout: embedding + classification weights passed as logits, noise_z = z - z_
        cfg = Config(B=B, P=P, S=S, D=D, H=kw["H"], dff=kw["dff"], F_x=kw["V"], F_y=kw["V"], attn_window=kw["attn_window"],
                     weights="classification", input_mode="embedding", noise_z_mode="checkout",
                     len_resampling=kw.get("len_resampling"))
        x_in = g["x_in"].reshape(B, S).astype(np.int32)
        y = g["y"].reshape(B, S).astype(np.int32)
    else:
        cfg = Config(B=B, P=P, S=S, D=D, H=kw["H"], dff=kw["dff"], F_x=reg["F_x"], F_y=kw["V"], attn_window=kw["attn_window"],
                     weights="gaussian", input_mode="linear", noise_z_mode="checkout", sigma_obs=reg["sigma_obs"],
                     len_resampling=kw.get("len_resampling"))
        x_in = g["x_in"].reshape(B, S, reg["F_x"]).astype(np.float32)
        y = g["y"].reshape(B, S, kw["V"]).astype(np.float32)
    return g, cfg, p, x_in, y


def compare(name, got, g, close, loss_fn=None):
    """got: dict with pred, pred_resampl, K, V, R, w, noise_q, noise_z, noise_K (or noise_K_resampled)."""
    for k in ("pred", "pred_resampl", "K", "R", "noise_q", "noise_z"):
        close(name + ":" + k, got[k], g[k])
    # the regression fixtures record what the (patched) weight function returns and what is passed to
    # tf.random.categorical: the normalised LOG-weights; the checkout records softmax probabilities.
    close(name + ":w", got["w_as_returned_by_reference"], g["w"])
    close(name + ":V_last", got["V"][:, :, -1], g["V_last"])
    nk = got["noise_K"] if "noise_K" in got else got["noise_K_resampled"]
    close(name + ":noise_K", nk, g["noise_K"])
