"""CPU oracle for the SMC-Transformer particle-filter forward pass.

*** TEST INFRASTRUCTURE — NOT PRODUCT CODE. ***
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``smc-t-v2_b200/``) never imports anything under ``oracle/`` and has no CPU fallback.

What it is: a numpy (fp32 arithmetic, fp64 CDF) restatement of the reference's
algorithm, with every random draw an explicit input.  Paths below are relative to
``/root/reference``.

Pinning status (see DESIGN.md "Oracle"):
  * TensorFlow / tensorflow-probability are not installed and not vendored, so the
    reference cannot be executed as-is.  The oracle is pinned by
      (1) the reference's own known-answer tables for the ancestor gather
          (``src/models/SMC_Transformer/transformer_utils.py:66-88``) and the
          squared-norm KAT (``src/models/SMC_Transformer/SMC_Transformer.py:361-365``);
      (2) the outputs of the reference's *unmodified Python sources* executed in this
          container over a small TensorFlow-API shim (``oracle/tf_shim``; generator
          ``oracle/gen_golden_from_reference.py``; fixtures ``tests/golden/ref_*.npz``).
    The arithmetic that lives inside TensorFlow itself (``tf.random.categorical``'s
    inverse-CDF sampler, Keras LayerNormalization, Dense) is restated from its
    published algorithm: **parity unpinned at the TensorFlow boundary**.
  * The Gaussian observation weight ``compute_w_regression`` is *called* in the
    reference (``SMC_TransformerCell.py:206``) but not defined in the checkout; its form
    here is the north-star specification, not a quotation.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional

import numpy as np

F32 = np.float32
NEG_MASK = F32(-1e9)  # self_attention_SMC.py:126,131


# --------------------------------------------------------------------------------------
# configuration / parameters
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class Config:
    B: int
    P: int
    S: int
    D: int
    H: int = 1
    dff: int = 0
    F_x: int = 1          # input features (linear mode) or vocabulary size (embedding mode)
    F_y: int = 1          # target features (gaussian) or number of classes (classification)
    full_model: bool = True
    attn_window: Optional[int] = None
    len_resampling: Optional[int] = None
    weights: str = "gaussian"        # "gaussian" (north star) | "classification" (checkout)
    noise_z_mode: str = "checkout"   # "checkout": z - z_  (self_attention_SMC.py:190) | "pyc": z - dense(z_)
    input_mode: str = "linear"       # "linear" (pyc projection_layer_ts) | "embedding" (SMC_Transformer.py:56,170) | "encoded"
    noise: bool = True
    ln_eps: float = 1e-6             # SMC_TransformerCell.py:36-37
    sigma_obs: float = 0.5           # pyc smoke default (variance of the observation model)


PARAM_NAMES = ("Wq", "bq", "Wk", "bk", "Wv", "bv", "Wz", "bz", "ln1_g", "ln1_b", "ln2_g", "ln2_b",
               "W1", "b1", "W2", "b2", "Wo", "W_in", "b_in", "logvar_k", "logvar_q", "logvar_v", "logvar_z")


def _glorot(rng, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(F32)


def init_params(cfg: Config, seed: int = 7, sigma2: float = 0.5, multi: bool = False,
                random_bias: bool = False) -> Dict[str, np.ndarray]:
    """Keras defaults: Glorot-uniform kernels (in,out), zero biases, LN gamma=1 beta=0
    (self_attention_SMC.py:21-24, SMC_TransformerCell.py:36-40, classic_layers.py:19-23).
    ``random_bias`` perturbs biases / LN affine so tests exercise them.
    logvar = log(sigma^2) (self_attention_SMC.py:64-67); ``multi`` = diagonal of a (D,D) matrix (:59-62)."""
    rng = np.random.default_rng(seed)
    D, dff = cfg.D, max(cfg.dff, 1)
    p = {}
    for n in ("q", "k", "v", "z"):
        p["W" + n] = _glorot(rng, D, D)
        p["b" + n] = np.zeros(D, F32)
    p["ln1_g"], p["ln1_b"] = np.ones(D, F32), np.zeros(D, F32)
    p["ln2_g"], p["ln2_b"] = np.ones(D, F32), np.zeros(D, F32)
    p["W1"], p["b1"] = _glorot(rng, D, dff), np.zeros(dff, F32)
    p["W2"], p["b2"] = _glorot(rng, dff, D), np.zeros(D, F32)
    p["Wo"] = _glorot(rng, D, cfg.F_y)
    if cfg.input_mode == "embedding":
        # Keras Embedding default initializer: uniform(-0.05, 0.05)
        p["W_in"] = rng.uniform(-0.05, 0.05, size=(cfg.F_x, D)).astype(F32)
        p["b_in"] = np.zeros(D, F32)
    else:
        p["W_in"], p["b_in"] = _glorot(rng, cfg.F_x, D), np.zeros(D, F32)
    if random_bias:
        for n in ("bq", "bk", "bv", "bz", "b1", "b2", "b_in", "ln1_b", "ln2_b"):
            p[n] = (0.1 * rng.standard_normal(p[n].shape)).astype(F32)
        for n in ("ln1_g", "ln2_g"):
            p[n] = (1.0 + 0.1 * rng.standard_normal(p[n].shape)).astype(F32)
    for i, n in enumerate(("k", "q", "v", "z")):
        if multi:
            s2 = sigma2 * (1.0 + 0.25 * np.cos(np.arange(D) + i))
            p["logvar_" + n] = np.log(s2).astype(F32)
        else:
            p["logvar_" + n] = np.asarray(math.log(sigma2), F32)
    return p


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def layer_norm(x, g, b, eps):
    """Keras LayerNormalization(epsilon=1e-6) takes the non-fused path (eps < 1.001e-5):
    tf.nn.moments (biased variance) then tf.nn.batch_normalization:
    inv = rsqrt(var + eps) * gamma ; out = x * inv + (beta - mean * inv).
    Call sites: SMC_TransformerCell.py:140,147,150."""
    x = x.astype(F32, copy=False)
    mean = x.mean(axis=-1, keepdims=True, dtype=F32)
    var = np.square(x - mean).mean(axis=-1, keepdims=True, dtype=F32)
    inv = (F32(1.0) / np.sqrt(var + F32(eps))) * g
    return (x * inv + (b - mean * inv)).astype(F32)


def dense(x, W, b=None):
    """tf.keras.layers.Dense: x @ kernel(in,out) + bias."""
    y = np.matmul(x.astype(F32, copy=False), W)
    if b is not None:
        y = y + b
    return y.astype(F32)


def encode_input(cfg: Config, p, x_in):
    """SMC_Transformer.get_encoded_input (SMC_Transformer.py:166-172): the raw input is
    tiled to P particles *before* the embedding/projection, so every particle of a
    sequence carries the identical encoded vector; we keep one copy (B,S,D).
    checkout: Embedding(tokens) * sqrt(D)   (:56,:170-172)
    pyc (regression era): Dense(D, name='projection_layer_ts')(x) * sqrt(D)."""
    if cfg.input_mode == "encoded":
        return np.asarray(x_in, F32)
    if cfg.input_mode == "embedding":
        tok = np.asarray(x_in).reshape(cfg.B, cfg.S).astype(np.int64)
        X = p["W_in"][tok]
    else:
        X = dense(np.asarray(x_in, F32).reshape(cfg.B, cfg.S, cfg.F_x), p["W_in"], p["b_in"])
    return (X * F32(math.sqrt(cfg.D))).astype(F32)


def noise_std(logvar):
    """self_attention_SMC.py:98-107: std = exp(0.5*logvar); scalar, or the diagonal of a (D,D) matrix."""
    return np.exp(np.asarray(logvar, F32) * F32(0.5)).astype(F32)


def attention_step(cfg: Config, q, K, V, t):
    """self_attention_SMC.py:110-138 after k_t, v_t were inserted at slot t.
    q (B,P,D); K,V (B,P,S,D).  Returns z_ (B,P,D) and attention weights (B,P,H,S).
    Computed over all S slots with -1e9 on disallowed ones, exactly like the reference."""
    B, P, S, D, H = cfg.B, cfg.P, cfg.S, cfg.D, cfg.H
    Dh = D // H
    qh = q.reshape(B, P, H, Dh)
    Kh = K.reshape(B, P, S, H, Dh)
    Vh = V.reshape(B, P, S, H, Dh)
    logits = np.einsum("bphd,bpshd->bphs", qh, Kh, dtype=F32) / F32(math.sqrt(Dh))  # :119-122
    masked = np.full_like(logits, NEG_MASK)
    w = cfg.attn_window
    lo = (t - w) if (w is not None and t > w) else 0  # :124-133
    masked[..., lo:t + 1] = logits[..., lo:t + 1]
    m = masked.max(axis=-1, keepdims=True)
    e = np.exp(masked - m, dtype=F32)
    att = (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)  # :135
    z = np.einsum("bphs,bpshd->bphd", att, Vh, dtype=F32).reshape(B, P, D)  # :136, concat_heads :72-79
    return z.astype(F32), att


def categorical_tf_cpu(logits, u):
    """tf.random.categorical(logits, num_samples) as executed by TensorFlow's CPU kernel
    (tensorflow/core/kernels/multinomial_op.cc, TF 2.6 — third-party, not vendored; restated
    from the published algorithm; call site SMC_TransformerCell.py:166):
      max over finite logits; cdf = running sum (sequential, fp64) of exp(double(logit) - double(max));
      per draw: target = U * cdf[-1]; index = upper_bound(cdf, target).
    The uniform U is injected (u, fp64 in [0,1)).  The index is clamped to P-1 (TF does not
    clamp; U*total == total can only happen through rounding).  logits (B,P) fp32 -> (B,P) int64."""
    logits = np.asarray(logits, F32)
    u = np.asarray(u, np.float64)
    B, P = logits.shape
    finite = np.isfinite(logits)
    mx = np.where(finite, logits, -np.inf).max(axis=1, keepdims=True).astype(np.float64)
    e = np.where(finite, np.exp(logits.astype(np.float64) - mx), 0.0)
    cdf = np.cumsum(e, axis=1)  # ufunc.accumulate: strictly sequential in fp64
    target = u * cdf[:, -1:]
    idx = np.empty((B, u.shape[1]), np.int64)
    for b in range(B):
        idx[b] = np.searchsorted(cdf[b], target[b], side="right")
    return np.minimum(idx, P - 1)


def cdf_margin(logits, u):
    """Relative distance of each target to the nearest CDF boundary (free-running parity
    protocol L4): tests reject draws that sit closer than a margin."""
    logits = np.asarray(logits, F32)
    mx = logits.max(axis=1, keepdims=True).astype(np.float64)
    cdf = np.cumsum(np.exp(logits.astype(np.float64) - mx), axis=1)
    tgt = np.asarray(u, np.float64) * cdf[:, -1:]
    d = np.abs(tgt[:, :, None] - cdf[:, None, :]).min(axis=2) / cdf[:, -1:]
    return d


def resample(params, i_t):
    """transformer_utils.py:39-47: tf.gather(params, i_t, axis=1, batch_dims=1):
    out[b, m, ...] = params[b, i_t[b, m], ...]."""
    i_t = np.asarray(i_t)
    idx = i_t.reshape(i_t.shape + (1,) * (params.ndim - 2))
    return np.take_along_axis(params, idx, axis=1)


def resample_rows_le_t(params, i_t, t):
    """Older semantics pinned by the reference's KAT tables (transformer_utils.py:66-88, pyc
    resample(params, i_t, t)): gather rows <= t, keep rows > t."""
    out = params.copy()
    out[:, :, :t + 1] = resample(params[:, :, :t + 1], i_t)
    return out


def compute_w_classification(pred, y_tok):
    """SMC_TransformerCell.py:78-84: softmax over classes, probability of the target token
    (unnormalised over particles, *not* a log).  pred (B,P,V), y_tok (B,) or (B,P)."""
    m = pred.max(axis=-1, keepdims=True)
    e = np.exp(pred - m, dtype=F32)
    prob = (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)
    y_tok = np.asarray(y_tok).astype(np.int64)
    if y_tok.ndim == 1:
        y_tok = np.broadcast_to(y_tok[:, None], pred.shape[:2])
    return np.take_along_axis(prob, y_tok[..., None], axis=-1)[..., 0]


def compute_w_regression(pred, y, sigma_obs):
    """North-star Gaussian observation weight (evidence: pyc loss
    0.5*(1/Sigma_obs)*einsum('bijk,bijk->bij', diff, diff)):
      logw[b,p] = -0.5 * ||y[b] - pred[b,p]||^2 / Sigma_obs ;  logw_hat = logw - logsumexp_p(logw).
    Returns (logw, logw_hat, w_hat = exp(logw_hat)).  pred (B,P,F), y (B,F)."""
    diff = (y[:, None, :] - pred).astype(F32)
    sq = np.einsum("bpf,bpf->bp", diff, diff, dtype=F32)
    logw = (F32(-0.5) * sq / F32(sigma_obs)).astype(F32)
    m = logw.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(logw - m, dtype=F32).sum(axis=1, keepdims=True, dtype=F32), dtype=F32)
    logw_hat = (logw - lse).astype(F32)
    return logw, logw_hat, np.exp(logw_hat, dtype=F32)


# --------------------------------------------------------------------------------------
# the forward pass
# --------------------------------------------------------------------------------------
def forward(cfg: Config, p, x_in, y, eps=None, u=None, i_forced=None, trace: bool = False,
            fast_gather: bool = True):
    """SMC_Transformer.call (SMC_Transformer.py:183-248) driving SMC_Transf_Cell.call
    (SMC_TransformerCell.py:129-184) for t = 0..S-1.

    x_in : (B,S,F_x) float | (B,S) int tokens | (B,S,D) pre-encoded
    y    : (B,S,F_y) float (gaussian) | (B,S) int tokens (classification)
    eps  : (4,S,B,P,D) fp32 standard normals in the order k,q,v,z (self_attention_SMC.py:160-162,189)
    u    : (S,B,P) fp64 uniforms for the categorical draw (SMC_TransformerCell.py:166)
    i_forced : optional (S,B,P) ancestor indices that replace the draw (teacher forcing, L3).
    fast_gather : gather only rows <= t (result-identical: rows > t are zero for every
                  particle, SMC_Transformer.py:197-202) — keeps the oracle usable at S=128.
    """
    B, P, S, D = cfg.B, cfg.P, cfg.S, cfg.D
    X = encode_input(cfg, p, x_in)                      # (B,S,D)   SMC_Transformer.py:192
    if cfg.weights == "classification":
        y_arr = np.asarray(y).reshape(B, S).astype(np.int64)
    else:
        y_arr = np.asarray(y, F32).reshape(B, S, cfg.F_y)
    K = np.zeros((B, P, S, D), F32)                     # :197-202
    V = np.zeros((B, P, S, D), F32)
    R = np.zeros((B, P, S, D), F32)
    sk, sq, sv, sz = (noise_std(p["logvar_" + n]) for n in ("k", "q", "v", "z"))
    out = dict(
        r_unres=np.zeros((B, P, S, D), F32), pred=np.zeros((B, P, S, cfg.F_y), F32),
        w=np.ones((B, P, S), F32), logw=np.zeros((S, B, P), F32), logits=np.zeros((S, B, P), F32),
        i_t=np.zeros((S, B, P), np.int64), noise_q=np.zeros((B, P, S, D), F32),
        noise_z=np.zeros((B, P, S, D), F32), attn=np.zeros((B, P, cfg.H, S, S), F32),
    )
    steps: List[dict] = []
    for t in range(S):
        x = layer_norm(X[:, t, :], p["ln1_g"], p["ln1_b"], cfg.ln_eps)[:, None, :]   # Cell:140  (B,1,D)
        k_ = dense(x, p["Wk"], p["bk"])                                              # attn:151-153
        q_ = dense(x, p["Wq"], p["bq"])
        v_ = dense(x, p["Wv"], p["bv"])
        if cfg.noise:                                                                # attn:155-165
            k = (k_ + sk * eps[0, t]).astype(F32)
            q = (q_ + sq * eps[1, t]).astype(F32)
            v = (v_ + sv * eps[2, t]).astype(F32)
        else:
            k, q, v = (np.broadcast_to(a, (B, P, D)).astype(F32) for a in (k_, q_, v_))
        out["noise_q"][:, :, t] = q - q_
        K[:, :, t, :] = k                                                            # attn:111-116
        V[:, :, t, :] = v
        z_, att = attention_step(cfg, q, K, V, t)                                    # attn:119-136
        out["attn"][:, :, :, t, :] = att
        z_proj = dense(z_, p["Wz"], p["bz"])                                         # attn:187
        if cfg.noise:
            z = (z_proj + sz * eps[3, t]).astype(F32)                                # attn:189
            out["noise_z"][:, :, t] = (z - z_) if cfg.noise_z_mode == "checkout" else (z - z_proj)  # attn:190
        else:
            z = z_proj
        if cfg.full_model:                                                           # Cell:145-150
            o = layer_norm(z + x, p["ln1_g"], p["ln1_b"], cfg.ln_eps)                # same LN1, normalised x
            h = np.maximum(dense(o, p["W1"], p["b1"]), F32(0))
            r = layer_norm(dense(h, p["W2"], p["b2"]) + o, p["ln2_g"], p["ln2_b"], cfg.ln_eps)
        else:
            r = z                                                                    # Cell:152
        pred_t = dense(r, p["Wo"])                                                   # Cell:154
        R[:, :, t, :] = r                                                            # Cell:159-161
        out["r_unres"][:, :, t] = r
        out["pred"][:, :, t] = pred_t
        st = None
        if trace:
            st = dict(x=x[:, 0], k_=k_[:, 0], q_=q_[:, 0], v_=v_[:, 0], k=k, q=q, v=v, z_=z_, z=z, r=r, pred=pred_t)
        if cfg.noise:                                                                # Cell:164-175
            if cfg.weights == "classification":
                w = compute_w_classification(pred_t, y_arr[:, t])
                logits = w                       # passed AS LOGITS (Cell:165-166)
                out["w"][:, :, t] = w
                out["logw"][t] = w
            else:
                logw, logw_hat, w_hat = compute_w_regression(pred_t, y_arr[:, t], cfg.sigma_obs)
                logits = logw_hat
                out["w"][:, :, t] = w_hat
                out["logw"][t] = logw
            out["logits"][t] = logits
            i_t = categorical_tf_cpu(logits, u[t]) if i_forced is None else np.asarray(i_forced[t], np.int64)
            out["i_t"][t] = i_t
            if cfg.len_resampling is None or t < cfg.len_resampling:                 # Cell:169-172
                if fast_gather:
                    K[:, :, :t + 1] = resample(K[:, :, :t + 1], i_t)
                    V[:, :, :t + 1] = resample(V[:, :, :t + 1], i_t)
                    R[:, :, :t + 1] = resample(R[:, :, :t + 1], i_t)
                else:
                    KVR = resample(np.concatenate([K, V, R], axis=-1), i_t)
                    K, V, R = (np.ascontiguousarray(a) for a in np.split(KVR, 3, axis=-1))
            if trace:
                st.update(i_t=i_t, logits=logits.copy())
        if trace:
            steps.append(st)
    # after the loop: SMC_Transformer.py:218-248
    out["pred_resampl"] = dense(R, p["Wo"])                                          # :231
    kx = dense(X, p["Wk"], p["bk"])                                                  # :235 (un-normalised X)
    vx = dense(X, p["Wv"], p["bv"])                                                  # :236
    out["noise_K_resampled"] = (K - kx[:, None]).astype(F32)
    out["noise_V_resampled"] = (V - vx[:, None]).astype(F32)
    out["K"], out["V"], out["R"] = K, V, R
    out["X"] = X
    if not cfg.noise:
        out["w"][:] = 1.0                                                            # :246
    if trace:
        out["steps"] = steps
    return out


def forecast_multistep(cfg: Config, p, x_past, y_past, eps, u, obs_noise, past_len, future_len):
    """Regression-era multi-step forecast (src/eval/inference_functions.py:19-39 in its time-series form, recovered from
    the shipped .pyc `inference_multistep`): run the particle filter on the observed past with resampling stopped at
    past_len, then roll forward feeding every particle its own last prediction plus N(0, Sigma_obs) observation noise.
    With injected draws indexed by the absolute timestep, re-running the whole forward for every future step (what the
    reference does) and continuing it step by step give the same states, so this loop is one pass.

    x_past (B,past_len,F), y_past (B,past_len,F_y); eps (4,S,B,P,D), u (S,B,P) with S = past_len + future_len;
    obs_noise (S,B,P,F_y) standard normals (entry t is added to the prediction of step t).
    Returns preds_future (B,P,future_len,F_y)."""
    B, P, D = cfg.B, cfg.P, cfg.D
    S = past_len + future_len
    assert cfg.F_x == cfg.F_y and cfg.weights == "gaussian" and cfg.full_model and cfg.noise
    K = np.zeros((B, P, S, D), F32)
    V = np.zeros((B, P, S, D), F32)
    R = np.zeros((B, P, S, D), F32)
    sk, sq, sv, sz = (noise_std(p["logvar_" + n]) for n in ("k", "q", "v", "z"))
    cfgS = Config(**{**cfg.__dict__, "S": S})
    W_in, b_in = np.asarray(p["W_in"], F32), np.asarray(p["b_in"], F32)
    sqrtD = F32(math.sqrt(D))
    preds = []
    x_next = None
    sd_obs = F32(math.sqrt(cfg.sigma_obs))
    for t in range(S):
        if t < past_len:
            xin = np.broadcast_to(np.asarray(x_past, F32)[:, None, t, :], (B, P, cfg.F_x))
        else:
            xin = x_next
        X = ((dense(xin, W_in, b_in)) * sqrtD).astype(F32)                                  # get_encoded_input (pyc)
        x = layer_norm(X, p["ln1_g"], p["ln1_b"], cfg.ln_eps)
        k = (dense(x, p["Wk"], p["bk"]) + sk * eps[0, t]).astype(F32)
        q = (dense(x, p["Wq"], p["bq"]) + sq * eps[1, t]).astype(F32)
        v = (dense(x, p["Wv"], p["bv"]) + sv * eps[2, t]).astype(F32)
        K[:, :, t, :] = k
        V[:, :, t, :] = v
        z_, _ = attention_step(cfgS, q, K, V, t)
        z = (dense(z_, p["Wz"], p["bz"]) + sz * eps[3, t]).astype(F32)
        o = layer_norm(z + x, p["ln1_g"], p["ln1_b"], cfg.ln_eps)
        h = np.maximum(dense(o, p["W1"], p["b1"]), F32(0))
        r = layer_norm(dense(h, p["W2"], p["b2"]) + o, p["ln2_g"], p["ln2_b"], cfg.ln_eps)
        pred_t = dense(r, p["Wo"])
        R[:, :, t, :] = r
        if t < past_len:                                                                    # resampling stops at past_len
            _, logw_hat, _ = compute_w_regression(pred_t, np.asarray(y_past, F32)[:, t], cfg.sigma_obs)
            i_t = categorical_tf_cpu(logw_hat, u[t])
            K[:, :, :t + 1] = resample(K[:, :, :t + 1], i_t)
            V[:, :, :t + 1] = resample(V[:, :, :t + 1], i_t)
            R[:, :, :t + 1] = resample(R[:, :, :t + 1], i_t)
        else:
            preds.append(pred_t)
        x_next = (pred_t + sd_obs * np.asarray(obs_noise[t], F32)).astype(F32)
    return np.stack(preds, axis=2)


# --------------------------------------------------------------------------------------
# losses / metrics / EM
# --------------------------------------------------------------------------------------
def _neg_log_gaussian(noise, logvar):
    """compute_log_gaussian_density (SMC_Transformer.py:91-99): -log N(noise; 0, diag(exp(logvar))).
    = 0.5*sum(n^2/sigma^2) + 0.5*D*log(2pi) + 0.5*sum(logvar)."""
    D = noise.shape[-1]
    lv = np.broadcast_to(np.asarray(logvar, F32), (D,)).astype(F32)
    quad = (np.square(noise) / np.exp(lv)).sum(axis=-1, dtype=F32)
    return (F32(0.5) * quad + F32(0.5 * D * math.log(2 * math.pi)) + F32(0.5) * lv.sum(dtype=F32)).astype(F32)


def smc_loss(cfg: Config, p, out, y, variant: str = "checkout"):
    """compute_SMC_loss (SMC_Transformer.py:123-146) + compute_classic_loss (:148-164).
    variant "checkout": full -log N terms + sparse CE of pred_resampl (classification) —
                        with gaussian weights the classic term is the pyc MSE form.
    variant "pyc":      0.5*||n||^2/sigma^2 (no log-det terms) + 0.5*||y-pred_resampl||^2/Sigma_obs."""
    noises = [out["noise_K_resampled"], out["noise_q"], out["noise_V_resampled"], out["noise_z"]]
    logvars = [p["logvar_k"], p["logvar_q"], p["logvar_v"], p["logvar_z"]]
    if cfg.noise:
        if variant == "checkout":
            parts = [_neg_log_gaussian(n, lv) for n, lv in zip(noises, logvars)]
        else:
            parts = [(F32(0.5) * (np.square(n) / np.exp(np.broadcast_to(np.asarray(lv, F32), (cfg.D,)))).sum(-1, dtype=F32))
                     for n, lv in zip(noises, logvars)]
        smc = np.float64(np.sum(parts, axis=0, dtype=np.float64).mean())
    else:
        smc = 0.0
    pr = out["pred_resampl"]
    if cfg.weights == "classification":
        yt = np.asarray(y).reshape(cfg.B, 1, cfg.S).astype(np.int64)
        yt = np.broadcast_to(yt, pr.shape[:3])
        m = pr.max(-1, keepdims=True)
        lse = (m[..., 0] + np.log(np.exp(pr - m).sum(-1)))
        ce = lse - np.take_along_axis(pr, yt[..., None], -1)[..., 0]
        classic = np.float64(ce.astype(np.float64).mean())
    else:
        yy = np.asarray(y, F32).reshape(cfg.B, 1, cfg.S, cfg.F_y)
        diff = yy - pr
        classic = np.float64((0.5 / cfg.sigma_obs) * np.square(diff).sum(-1).astype(np.float64).mean())
    return smc + classic, classic


def mse_metric(out, y, cfg: Config):
    """pyc train_step_SMC_T metric: MSE of the particle-mean prediction (regression era)."""
    yy = np.asarray(y, F32).reshape(cfg.B, cfg.S, cfg.F_y)
    return float(np.square(out["pred"].mean(axis=1) - yy).astype(np.float64).mean())


def em_update(sigma2, noise, it, a):
    """train/utils.py:22-51: per batch element j in order,
    sigma^2 <- (1 - it^-a) sigma^2 + it^-a * mean_{p,s,d}(noise[j]^2)."""
    err = np.square(noise).mean(axis=(1, 2, 3))
    s2 = float(sigma2)
    g = float(it) ** (-a)
    for j in range(err.shape[0]):
        s2 = (1 - g) * s2 + g * float(err[j])
    return s2


# --------------------------------------------------------------------------------------
# synthetic data
# --------------------------------------------------------------------------------------
def synthetic_series(B, S, F, alpha=0.8, var=0.5, seed=1234):
    """AR(1) x_{t+1} = alpha x_t + sqrt(var) eps; inputs = x[:, :-1], targets = x[:, 1:]
    (data_provider/class_datasets.py:19-22)."""
    rng = np.random.default_rng(seed)
    x = np.zeros((B, S + 1, F), np.float64)
    x[:, 0] = rng.standard_normal((B, F))
    for t in range(S):
        x[:, t + 1] = alpha * x[:, t] + math.sqrt(var) * rng.standard_normal((B, F))
    x = x.astype(F32)
    return x[:, :-1].copy(), x[:, 1:].copy()


def draw_randomness(cfg: Config, seed=99):
    rng = np.random.default_rng(seed)
    eps = rng.standard_normal((4, cfg.S, cfg.B, cfg.P, cfg.D)).astype(F32)
    u = rng.random((cfg.S, cfg.B, cfg.P))
    return eps, u
