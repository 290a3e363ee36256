"""Gradient, EM and loss fixtures written by the REFERENCE'S OWN PYTHON SOURCES (TEST INFRASTRUCTURE).

    python -m oracle.gen_golden_grads_from_reference      # needs /root/reference (this container only)

Same set-up as oracle/gen_golden_from_reference.py (the reference modules imported unmodified from /root/reference
over the TensorFlow-API shim, injected draws), plus what `train_step_SMC_T` does with the forward
(src/train/train_step_functions.py:43-63):

  * `with tf.GradientTape() as tape:` forward -> `compute_SMC_loss` -> `tape.gradient(loss, trainable_variables)`.
    The shim's ops are torch ops, so the tape is torch autograd over the reference's own forward and loss code
    (oracle/tf_shim.py: GradientTape).  Classification specs use the checkout's `compute_SMC_loss` as it is; regression
    specs use the regression-era (pyc) loss form, like the forward fixtures.
    One documented deviation: the recorded `noise_z` enters the loss DETACHED.  The checkout records noise_z = z - z_
    (self_attention_SMC.py:190), whose gradient involves particles outside the final genealogy; the CUDA backward serves
    the regression-era noise_z = z - dense(z_) (a constant of the parameters).  All parameter gradients are therefore
    comparable; the gradient of logvar_z is not (it depends on which noise_z is recorded) and is not stored.
  * `EM(smc_transformer, it, EM_param)` of src/train/utils.py:22-51 executed on the model after the forward: the four
    updated log-variances.

Fixtures: tests/golden/ref_grad_*.npz.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tf_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run_reference(B, P, S, D, dff, V, attn_window, sigma, seed, regression=None, len_resampling=None, em=(3, 0.6), multi=False,
                  masked=False):
    tf_shim.TRACK_GRADIENTS = True
    try:
        inst = tf_shim.install()
        tf, rng = inst.tf, inst.rng
        rng.np_rng = np.random.default_rng(seed)
        from src.models.SMC_Transformer.SMC_Transformer import SMC_Transformer
        from src.train.utils import EM
        model = SMC_Transformer(d_model=D, output_size=V, seq_len=S, full_model=True, dff=dff, num_layers=1, num_heads=1,
                                attn_window=attn_window)
        base = [sigma * 1.0, sigma * 1.3, sigma * 0.8, sigma * 1.1]
        if multi:   # noise_dim = "multi" (run_SMC_T.py:75-82): a list of d_model variances per noise -> (D,D) diagonal variables
            sig = [[float(sv * (1.0 + 0.07 * i)) for i in range(D)] for sv in base]
        else:
            sig = base
        model.cell.add_SMC_parameters(dict_sigmas=dict(zip(["k", "q", "v", "z"], sig)), num_particles=P)
        if len_resampling is not None:
            model.cell.add_stop_resampling(len_resampling)
        g = np.random.default_rng(seed + 1)
        eps = g.standard_normal((4, S, B, P, D)).astype(np.float32)
        u = g.random((S, B, P))
        proj = None
        if regression is None:
            x_in = g.integers(0, V, size=(B, 1, S, 1)).astype(np.int64)
            y = g.integers(0, V, size=(B, 1, S, 1)).astype(np.int64)
        else:
            F_x, sigma_obs = regression["F_x"], regression["sigma_obs"]
            x_in = g.standard_normal((B, 1, S, F_x)).astype(np.float32)
            y = (0.5 * g.standard_normal((B, 1, S, V))).astype(np.float32)
            proj = tf.keras.layers.Dense(D, name="projection_layer_ts")

            def get_encoded_input(inputs, attention_mask=None):   # pyc: Dense(d_model)(x) * sqrt(d_model), after tiling
                if tf.shape(inputs)[1] == 1:
                    inputs = tf.tile(inputs, multiples=[1, model.cell.num_particles, 1, 1])
                return proj(inputs) * tf.math.sqrt(tf.cast(model.d_model, tf.float32))

            def compute_w_regression(predictions, y):             # north-star Gaussian weight, log form, normalised
                diff = tf.squeeze(y - predictions, axis=-2)
                logw = -0.5 * tf.einsum("bpf,bpf->bp", diff, diff) / sigma_obs
                return logw - torch.logsumexp(logw, dim=1, keepdim=True)
            model.get_encoded_input = get_encoded_input
            model.cell.compute_w_classification = compute_w_regression
        cnt = {"n": 0, "c": 0}
        picked = []

        def normal_fn(shape):
            which, call = cnt["n"] % 4, cnt["n"] // 4
            cnt["n"] += 1
            return torch.from_numpy(eps[which, max(call - 1, 0)]).reshape(shape)

        def uniform_fn(Bn, ns):
            t = max(cnt["c"] - 1, 0)
            cnt["c"] += 1
            return torch.from_numpy(u[t])
        rng.normal_fn, rng.uniform_fn = normal_fn, uniform_fn
        tx, ty = torch.from_numpy(x_in), torch.from_numpy(y)
        mask = None
        if masked:   # NLP padding mask (B,1,S,1): ones up to a per-sequence length, zeros behind it
            lens = g.integers(max(1, S // 2), S + 1, size=B)
            mask = (np.arange(S)[None, :] < lens[:, None]).astype(np.int32).reshape(B, 1, S, 1)
        tmask = None if mask is None else torch.from_numpy(mask)
        with tf.GradientTape() as tape:
            (pred, pred_r), (K, Vv, R), w = model(inputs=tx, targets=ty, attention_mask=tmask)
            assert cnt["n"] == 4 * (S + 1) and cnt["c"] == S + 1
            # the recorded noise_z enters the loss detached (module docstring)
            model.noise_z = model.noise_z.detach()
            model.internal_noises = [model.noise_K_resampled, model.noise_q, model.noise_V_resampled, model.noise_z]
            att = model.cell.attention_smc
            if regression is None:
                total, classic = model.compute_SMC_loss(inputs=None, targets=ty, predictions=pred_r, attention_mask=tmask)
            else:
                parts = []
                for noise, lv in zip(model.internal_noises, [att.logvar_k, att.logvar_q, att.logvar_v, att.logvar_z]):
                    var = torch.exp(lv.value) if lv.value.dim() == 0 else torch.exp(torch.diagonal(lv.value))
                    parts.append(0.5 * (noise * noise / var).sum(-1))
                classic = (0.5 / regression["sigma_obs"]) * ((ty - pred_r) ** 2).sum(-1).mean()
                total = torch.stack(parts).sum(0).mean() + classic
        c = model.cell
        named = dict(Wq=att.wq.kernel, bq=att.wq.bias, Wk=att.wk.kernel, bk=att.wk.bias, Wv=att.wv.kernel, bv=att.wv.bias,
                     Wz=att.dense.kernel, bz=att.dense.bias, ln1_g=c.layernorm1.gamma, ln1_b=c.layernorm1.beta,
                     ln2_g=c.layernorm2.gamma, ln2_b=c.layernorm2.beta, W1=c.ffn.layers[0].kernel, b1=c.ffn.layers[0].bias,
                     W2=c.ffn.layers[1].kernel, b2=c.ffn.layers[1].bias, Wo=c.output_layer.kernel,
                     logvar_k=att.logvar_k, logvar_q=att.logvar_q, logvar_v=att.logvar_v, logvar_z=att.logvar_z)
        if regression is None:
            named["W_in"] = model.embedding.embeddings
        else:
            named["W_in"], named["b_in"] = proj.kernel, proj.bias
        names = sorted(named)
        grads = tape.gradient(total, [named[n] for n in names])
        out_g = {}
        for n, gr in zip(names, grads):
            if n == "logvar_z":
                continue
            gr = torch.zeros_like(named[n].value) if gr is None else gr
            if n.startswith("logvar_") and gr.dim() == 2:
                gr = torch.diagonal(gr)          # only the diagonal of the (D,D) variable is used (tf.linalg.diag_part)
            out_g[n] = gr.detach().numpy().astype(np.float32).copy()
        params = {n: (np.diagonal(v.numpy()) if (n.startswith("logvar_") and v.numpy().ndim == 2) else v.numpy()).astype(np.float32).copy()
                  for n, v in named.items()}
        if regression is None:
            params["b_in"] = np.zeros(D, np.float32)
        # ---- EM on the model after the forward (src/train/utils.py:22-51) ----
        if multi:   # the reference's EM mode is scalar-only
            return dict(params=params, grads=out_g, x_in=x_in, y=y, eps=eps, u=u, loss_total=float(total), loss_classic=float(classic),
                        pred_resampl=pred_r.detach().numpy(), noise_z=model.noise_z.numpy(), em=None, lv_before=None, lv_after=None,
                        mask=mask)
        it, em_param = em
        lv_before = {n: float(getattr(att, "logvar_" + n).numpy()) for n in "kqvz"}
        with torch.no_grad():
            model.noise_K_resampled, model.noise_q = model.noise_K_resampled.detach(), model.noise_q.detach()
            model.noise_V_resampled = model.noise_V_resampled.detach()
            for n in "kqvz":   # EM mode keeps plain tensors (self_attention_SMC.py:48-52)
                setattr(att, "logvar_" + n, getattr(att, "logvar_" + n).value.detach().clone())
            EM(model, it=it, EM_param=em_param)
        lv_after = {n: float(getattr(att, "logvar_" + n)) for n in "kqvz"}
        return dict(params=params, grads=out_g, x_in=x_in, y=y, eps=eps, u=u, loss_total=float(total), loss_classic=float(classic),
                    pred_resampl=pred_r.detach().numpy(), noise_z=model.noise_z.numpy(),
                    em=np.array([it, em_param], np.float64), lv_before=lv_before, lv_after=lv_after, mask=mask)
    finally:
        tf_shim.TRACK_GRADIENTS = False


SPECS = {
    "ref_grad_regression_small": dict(B=3, P=12, S=8, D=16, dff=32, V=2, attn_window=None, sigma=0.3, seed=21,
                                      regression=dict(F_x=2, sigma_obs=0.5)),
    "ref_grad_regression_window_stop": dict(B=2, P=10, S=9, D=8, dff=8, V=1, attn_window=3, sigma=0.5, seed=22,
                                            regression=dict(F_x=1, sigma_obs=0.4), len_resampling=5),
    "ref_grad_regression_d64": dict(B=1, P=24, S=6, D=64, dff=256, V=1, attn_window=None, sigma=0.5, seed=23,
                                    regression=dict(F_x=1, sigma_obs=0.5)),
    "ref_grad_checkout_classification": dict(B=3, P=8, S=6, D=16, dff=16, V=20, attn_window=None, sigma=0.2, seed=24),
    # noise_dim = "multi": per-dimension variances, the configuration of the reference's published ROC runs
    "ref_grad_checkout_multi": dict(B=2, P=10, S=7, D=16, dff=16, V=15, attn_window=None, sigma=0.3, seed=25, multi=True),
    # attention_mask: the masked mean of compute_classic_loss (SMC_Transformer.py:158-163)
    "ref_grad_checkout_masked": dict(B=4, P=6, S=8, D=16, dff=16, V=12, attn_window=None, sigma=0.2, seed=27, masked=True),
    "ref_grad_regression_multi": dict(B=2, P=12, S=6, D=8, dff=16, V=2, attn_window=None, sigma=0.4, seed=26, multi=True,
                                      regression=dict(F_x=2, sigma_obs=0.5)),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, kw in SPECS.items():
        r = run_reference(**kw)
        arrays = dict(name=np.array(name), meta=np.array(repr(kw)), x_in=r["x_in"], y=r["y"], eps=r["eps"], u=r["u"],
                      loss_total=np.array(r["loss_total"]), loss_classic=np.array(r["loss_classic"]),
                      pred_resampl=r["pred_resampl"])
        if r.get("mask") is not None:
            arrays["mask"] = r["mask"]
        if r["em"] is not None:
            arrays.update(em=r["em"], lv_before=np.array([r["lv_before"][n] for n in "kqvz"]),
                          lv_after=np.array([r["lv_after"][n] for n in "kqvz"]))
        arrays.update({"p_" + k: v for k, v in r["params"].items()})
        arrays.update({"g_" + k: v for k, v in r["grads"].items()})
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **arrays)
        gn = {k: float(np.abs(v).max()) for k, v in r["grads"].items()}
        print(path, os.path.getsize(path) // 1024, "KiB loss", r["loss_total"], "max|grad|", {k: "%.2e" % v for k, v in gn.items()})
        if r["em"] is not None:
            print("   EM", r["lv_before"], "->", r["lv_after"])


if __name__ == "__main__":
    main()
