"""torch-free helpers shared by CPU tests, GPU tests and the golden generator."""
import numpy as np

from oracle import smct_oracle as O


def make_case(cfg, seed=0, sigma2=0.5, multi=False, random_bias=True):
    p = O.init_params(cfg, seed=7 + seed, sigma2=sigma2, multi=multi, random_bias=random_bias)
    rng = np.random.default_rng(1234 + seed)
    if cfg.input_mode == "embedding":
        x_in = rng.integers(0, cfg.F_x, size=(cfg.B, cfg.S)).astype(np.int32)
    elif cfg.input_mode == "encoded":
        x_in = rng.standard_normal((cfg.B, cfg.S, cfg.D)).astype(np.float32)
    else:
        x_in, _ = O.synthetic_series(cfg.B, cfg.S, cfg.F_x, seed=1234 + seed)
    if cfg.weights == "classification":
        y = rng.integers(0, cfg.F_y, size=(cfg.B, cfg.S)).astype(np.int32)
    else:
        y = (0.5 * rng.standard_normal((cfg.B, cfg.S, cfg.F_y))).astype(np.float32)
    eps, u = O.draw_randomness(cfg, seed=99 + seed)
    return p, x_in, y, eps, u


def screen_uniforms(cfg, p, x_in, y, eps, u, margin=1e-4, seed=5, max_iter=200):
    """Free-running parity protocol (L4): re-draw every uniform whose target sits within `margin`
    (relative to the total mass) of a CDF boundary of the oracle's own trajectory.  Log-weights of the CUDA
    paths agree with the oracle to ~1e-6 (fp32 paths) / ~1e-5 (tensor-core path, bf16x3 operands), so a draw
    that close to a boundary may legitimately land on the other side and fork the trajectory."""
    rng = np.random.default_rng(seed)
    u = u.copy()
    for _ in range(max_iter):
        out = O.forward(cfg, p, x_in, y, eps, u)
        bad_any = False
        for t in range(cfg.S):
            m = O.cdf_margin(out["logits"][t], u[t])
            bad = m < margin
            if bad.any():
                u[t][bad] = rng.random(int(bad.sum()))
                bad_any = True
                break  # the trajectory after t changes: recompute
        if not bad_any:
            return u, out
    raise RuntimeError("could not screen uniforms")
