"""Reference-equivalent CPU restatement with torch-CPU ops — the `cpu_baseline` / `--impl reference` arm.

*** TEST / MEASUREMENT INFRASTRUCTURE — NOT PRODUCT CODE. ***

TensorFlow is not installed in this image (nor on the GPU box), so the reference's TF path cannot
be timed.  This file restates, op for op, what the reference executes eagerly per timestep — same
data-movement pattern, not an optimised one:
  * inputs/targets tiled to P particles *before* the projection  (SMC_Transformer.py:166-172,193-194)
  * 3 Dense on (B,P,1,D), randn noise, k - k_                     (self_attention_SMC.py:151-165)
  * slice + concat to insert k_t, v_t (copies all of K and V)     (self_attention_SMC.py:111-116)
  * matmul(q, K^T) over the full padded length S, -1e9 mask by concat, softmax, matmul(att, V) (:119-136)
  * out Dense + noise, LN1/FFN/LN2, output Dense                  (SMC_TransformerCell.py:140-154)
  * slice + concat to store r_t in R                              (SMC_TransformerCell.py:159-161)
  * weights, categorical draw, concat[K,V,R] -> gather -> split   (SMC_TransformerCell.py:164-172)
torch.set_num_threads(n) gives it all host cores, like TF's Eigen thread pool.
With injected draws it reproduces oracle/smct_oracle.py (tests/test_oracle_cpu.py), which makes it a
second, independently written restatement of the same reference lines.
"""
import math

import torch


def _ln(x, g, b, eps):
    mean = x.mean(-1, keepdim=True)
    var = (x - mean).square().mean(-1, keepdim=True)
    inv = torch.rsqrt(var + eps) * g
    return x * inv + (b - mean * inv)


def _categorical(logits, u):
    """TF CPU multinomial: fp64 running sum of exp(logit - max), upper_bound(u * total)."""
    l64 = logits.double()
    cdf = torch.cumsum(torch.exp(l64 - l64.max(dim=1, keepdim=True).values), dim=1)
    tgt = u * cdf[:, -1:]
    return torch.searchsorted(cdf, tgt, right=True).clamp_(max=logits.shape[1] - 1)


def forward(cfg, p, x_in, y, eps=None, u=None, generator=None, i_forced=None):
    """cfg: oracle Config.  p: dict of torch CPU fp32 tensors.
    x_in (B,S,F_x) floats | (B,S) tokens (input_mode embedding); y (B,S,F_y) floats | (B,S) class ids (classification
    weights).  eps (4,S,B,P,D) / u (S,B,P) optional injected draws."""
    B, P, S, D, H = cfg.B, cfg.P, cfg.S, cfg.D, cfg.H
    Dh = D // H
    # get_encoded_input: tile first, then project (SMC_Transformer.py:166-172)
    classification = cfg.weights == "classification"
    if cfg.input_mode == "embedding":                                    # checkout: Embedding(tokens) * sqrt(D) (:56,:170)
        inp = x_in.reshape(B, 1, S).repeat(1, P, 1).long()
        X = p["W_in"][inp] * math.sqrt(D)
    else:                                                                # regression era (pyc): Dense projection_layer_ts
        inp = x_in.reshape(B, 1, S, cfg.F_x).repeat(1, P, 1, 1)
        X = (inp @ p["W_in"] + p["b_in"]) * math.sqrt(D)                 # (B,P,S,D)
    tgt = y.reshape(B, 1, S, -1).repeat(1, P, 1, 1)                       # (B,P,S,F_y) floats | (B,P,S,1) class ids
    K = torch.zeros(B, P, S, D)
    V = torch.zeros(B, P, S, D)
    R = torch.zeros(B, P, S, D)
    sk, sq, sv, sz = (torch.exp(0.5 * p["logvar_" + n]) for n in ("k", "q", "v", "z"))
    r_list, w_list, i_list, nq_list, nz_list = [], [], [], [], []

    def split_heads(a):  # self_attention_SMC.py:81-87
        return a.reshape(B, P, -1, H, Dh).permute(0, 1, 3, 2, 4)

    def concat_heads(a):  # :72-79
        return a.permute(0, 1, 3, 2, 4).reshape(B, P, -1, D)

    for t in range(S):
        x = X[:, :, t:t + 1, :]
        yt = tgt[:, :, t:t + 1, :]
        x = _ln(x, p["ln1_g"], p["ln1_b"], cfg.ln_eps)
        k_ = x @ p["Wk"] + p["bk"]
        q_ = x @ p["Wq"] + p["bq"]
        v_ = x @ p["Wv"] + p["bv"]
        ek, eq, ev, ez = ((eps[i, t].unsqueeze(2) if eps is not None else torch.randn(B, P, 1, D, generator=generator))
                          for i in range(4))
        k = k_ + sk.detach() * ek      # stop_gradient on the noise scale (self_attention_SMC.py:98-104)
        q = q_ + sq.detach() * eq
        v = v_ + sv.detach() * ev
        noise_q = q - q_
        kh, qh, vh, Kh, Vh = split_heads(k), split_heads(q), split_heads(v), split_heads(K), split_heads(V)
        Kh = torch.cat([Kh[:, :, :, :t], kh, Kh[:, :, :, t + 1:]], dim=3)
        Vh = torch.cat([Vh[:, :, :, :t], vh, Vh[:, :, :, t + 1:]], dim=3)
        logits = (qh @ Kh.transpose(-1, -2)) / math.sqrt(Dh)            # (B,P,H,1,S)
        w_ = cfg.attn_window
        if w_ is not None and t > w_:
            logits = torch.cat([-1e9 * torch.ones(B, P, H, 1, t - w_), logits[..., t - w_:t + 1],
                                -1e9 * torch.ones(B, P, H, 1, S - (t + 1))], dim=-1)
        else:
            logits = torch.cat([logits[..., :t + 1], -1e9 * torch.ones(B, P, H, 1, S - (t + 1))], dim=-1)
        att = torch.softmax(logits, dim=-1)
        zh = att @ Vh
        z_ = concat_heads(zh)
        K = concat_heads(Kh)
        V = concat_heads(Vh)
        z_proj = z_ @ p["Wz"] + p["bz"]
        z = z_proj + sz.detach() * ez
        noise_z = (z - z_) if cfg.noise_z_mode == "checkout" else (z - z_proj)
        if cfg.full_model:
            o = _ln(z + x, p["ln1_g"], p["ln1_b"], cfg.ln_eps)
            r = torch.relu(o @ p["W1"] + p["b1"]) @ p["W2"] + p["b2"]
            r = _ln(r + o, p["ln2_g"], p["ln2_b"], cfg.ln_eps)
        else:
            r = z
        pred = r @ p["Wo"]
        R = torch.cat([R[:, :, :t], r, R[:, :, t + 1:]], dim=2)
        if classification:                                                         # compute_w_classification (:78-84):
            probs = torch.softmax(pred.squeeze(2), dim=-1)                         # the probabilities go in as logits (:165-166)
            logw = torch.gather(probs, -1, yt.reshape(B, P, 1).long()).squeeze(-1)
        else:
            diff = yt - pred
            logw = (-0.5 * (diff * diff).sum(-1).squeeze(-1)) / cfg.sigma_obs     # (B,P)
            logw = logw - torch.logsumexp(logw, dim=1, keepdim=True)
        logw = logw.detach()                                                       # w, i_t are stop_gradient (:167)
        if i_forced is not None:
            i_t = i_forced[t].long()
        else:
            ut = u[t] if u is not None else torch.rand(B, P, dtype=torch.float64, generator=generator)
            i_t = _categorical(logw, ut)
        if cfg.len_resampling is None or t < cfg.len_resampling:
            KVR = torch.cat([K, V, R], dim=-1)                                     # SMC_TransformerCell.py:170-172
            KVR = torch.gather(KVR, 1, i_t.reshape(B, P, 1, 1).expand(B, P, S, 3 * D))
            K, V, R = torch.split(KVR, D, dim=-1)
            K, V, R = K.contiguous(), V.contiguous(), R.contiguous()
        r_list.append(r)
        w_list.append(logw if classification else torch.exp(logw))
        i_list.append(i_t)
        nq_list.append(noise_q)
        nz_list.append(noise_z)
    R_unres = torch.cat(r_list, dim=2)
    pred = R_unres @ p["Wo"]
    pred_resampl = R @ p["Wo"]
    noise_K = K - (X @ p["Wk"] + p["bk"])
    noise_V = V - (X @ p["Wv"] + p["bv"])
    return dict(pred=pred, pred_resampl=pred_resampl, K=K, V=V, R=R, w=torch.stack(w_list, dim=-1),
                i_t=torch.stack(i_list), noise_q=torch.cat(nq_list, 2), noise_z=torch.cat(nz_list, 2),
                noise_K_resampled=noise_K, noise_V_resampled=noise_V)


def loss_and_grads(cfg, p_np, x_in, y, eps, i_forced, variant="pyc", mask=None):
    """Gradient oracle: torch autograd through this restatement (fp64), with injected draws and
    teacher-forced ancestor indices.  Returns (loss, {name: grad ndarray}).  The reference takes
    tape.gradient(loss, trainable_variables) of exactly this graph (train_step_functions.py:43-63)."""
    import numpy as np
    tp = {k: torch.tensor(np.asarray(v, np.float64).reshape(()) if k.startswith("logvar") and np.size(v) == 1
                          else np.asarray(v, np.float64), dtype=torch.float64, requires_grad=True) for k, v in p_np.items()}
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        tx = torch.as_tensor(np.asarray(x_in)) if cfg.input_mode == "embedding" else torch.as_tensor(np.asarray(x_in, np.float64))
        classification = cfg.weights == "classification"
        ty = torch.as_tensor(np.asarray(y)) if classification else torch.as_tensor(np.asarray(y, np.float64))
        out = forward(cfg, tp, tx, ty, eps=torch.as_tensor(np.asarray(eps, np.float64)),
                      i_forced=torch.as_tensor(np.asarray(i_forced)))
        B, S = cfg.B, cfg.S
        tot = 0.0
        for n, lv in (("noise_K_resampled", "logvar_k"), ("noise_q", "logvar_q"), ("noise_V_resampled", "logvar_v"),
                      ("noise_z", "logvar_z")):
            part = 0.5 * (out[n].square() / torch.exp(tp[lv])).sum(-1)
            if variant == "checkout":
                D = out[n].shape[-1]
                part = part + 0.5 * D * math.log(2 * math.pi) + (0.5 * D * tp[lv] if tp[lv].dim() == 0 else 0.5 * tp[lv].sum())
            tot = tot + part.mean()
        if classification:   # compute_classic_loss (SMC_Transformer.py:148-164): sparse CE of the resampled logits
            pr = out["pred_resampl"]
            tgt = ty.reshape(B, 1, S).repeat(1, cfg.P, 1).long()
            ce = torch.nn.functional.cross_entropy(pr.reshape(-1, pr.shape[-1]), tgt.reshape(-1), reduction="none").reshape(B, cfg.P, S)
            if mask is None:
                loss = tot + ce.mean()
            else:   # masked mean of compute_classic_loss (SMC_Transformer.py:158-163): the mask is tiled over the particles
                m = torch.as_tensor(np.asarray(mask, np.float64)).reshape(B, 1, S).expand(B, cfg.P, S)
                loss = tot + (ce * m).sum() / m.sum()
        else:
            diff = ty.reshape(B, 1, S, cfg.F_y) - out["pred_resampl"]
            se = (0.5 / cfg.sigma_obs) * diff.square().sum(-1)
            if mask is None:
                loss = tot + se.mean()
            else:
                m = torch.as_tensor(np.asarray(mask, np.float64)).reshape(B, 1, S).expand(B, cfg.P, S)
                loss = tot + (se * m).sum() / m.sum()
        names = list(tp)
        grads = torch.autograd.grad(loss, [tp[n] for n in names], allow_unused=True)
    finally:
        torch.set_default_dtype(old)
    return float(loss), {n: (np.zeros_like(np.asarray(p_np[n], np.float64)) if g is None else g.detach().numpy().reshape(np.shape(p_np[n])))
                         for n, g in zip(names, grads)}


def loss_pyc(cfg, p, out, y):
    """Regression-era loss (pyc): mean_{b,p,s} sum_n 0.5||n||^2/sigma_n^2 + mean 0.5||y-pred_resampl||^2/Sigma_obs."""
    B, S = cfg.B, cfg.S
    tot = 0.0
    for n, lv in (("noise_K_resampled", "logvar_k"), ("noise_q", "logvar_q"), ("noise_V_resampled", "logvar_v"),
                  ("noise_z", "logvar_z")):
        tot = tot + (0.5 * (out[n].square() / torch.exp(p[lv])).sum(-1)).mean()
    diff = y.reshape(B, 1, S, cfg.F_y) - out["pred_resampl"]
    return float(tot + (0.5 / cfg.sigma_obs) * diff.square().sum(-1).mean())
