"""Host-side utilities of the product path: Keras-default parameter initialisation and the synthetic
series of the benchmark configs.  numpy only; no dependency on oracle/."""
import math

import numpy as np


def glorot_uniform(rng, fan_in, fan_out):
    """tf.keras 'glorot_uniform' (default kernel initialiser of Dense): U(-l, l), l = sqrt(6/(fan_in+fan_out))."""
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


def init_parameters(d_model, dff, F_x, F_y, sigma2=0.5, input_mode="linear", seed=7, logvar_dim=1):
    """Parameters of a 1-layer SMC_Transformer with Keras defaults: Glorot-uniform kernels (in,out), zero
    biases, LayerNorm gamma=1/beta=0, Embedding U(-0.05,0.05); log-variances = log(sigma^2)
    (self_attention_SMC.py:46-67)."""
    rng = np.random.default_rng(seed)
    D = d_model
    p = {}
    for n in ("q", "k", "v", "z"):
        p["W" + n] = glorot_uniform(rng, D, D)
        p["b" + n] = np.zeros(D, np.float32)
    for n in ("ln1", "ln2"):
        p[n + "_g"] = np.ones(D, np.float32)
        p[n + "_b"] = np.zeros(D, np.float32)
    f = max(dff, 1)
    p["W1"], p["b1"] = glorot_uniform(rng, D, f), np.zeros(f, np.float32)
    p["W2"], p["b2"] = glorot_uniform(rng, f, D), np.zeros(D, np.float32)
    p["Wo"] = glorot_uniform(rng, D, F_y)
    if input_mode == "embedding":
        p["W_in"] = rng.uniform(-0.05, 0.05, size=(F_x, D)).astype(np.float32)
    else:
        p["W_in"] = glorot_uniform(rng, F_x, D)
    p["b_in"] = np.zeros(D, np.float32)
    for n in ("k", "q", "v", "z"):
        p["logvar_" + n] = np.full((logvar_dim,), math.log(sigma2), np.float32)
    return p


def synthetic_ar1(B, S, F, alpha=0.9, var=0.3, seed=1234):
    """AR(1) series x_{t+1} = alpha x_t + sqrt(var) eps; returns inputs x[:, :-1] and targets x[:, 1:]
    (the split of data_provider/class_datasets.py:19-22)."""
    rng = np.random.default_rng(seed)
    x = np.empty((B, S + 1, F), np.float32)
    cur = rng.standard_normal((B, F)).astype(np.float32)
    x[:, 0] = cur
    sd = np.float32(math.sqrt(var))
    for t in range(S):
        cur = np.float32(alpha) * cur + sd * rng.standard_normal((B, F)).astype(np.float32)
        x[:, t + 1] = cur
    return np.ascontiguousarray(x[:, :-1]), np.ascontiguousarray(x[:, 1:])


def algorithmic_bytes_per_particle_step(D, S):
    """fp32, live rows only: attention reads rows 0..t-1 of K,V; k_t,v_t,r_t stores; ancestor
    gather reads+writes rows 0..t of K,V,R.  bytes(t) = 4D(8t+9); average over a trajectory = 4D(4S+5)."""
    return 4 * D * (4 * S + 5)
