"""Golden vectors written by the REFERENCE'S OWN PYTHON SOURCES (TEST INFRASTRUCTURE).

    python -m oracle.gen_golden_from_reference        # needs /root/reference (this container only)

The reference modules are imported unmodified from /root/reference over the TensorFlow-API shim
(oracle/tf_shim.py: TF is not installed) and driven with injected N(0,1) / U[0,1) draws.
Fixtures (tests/golden/ref_*.npz) hold weights, inputs, draws and the reference's outputs; they travel
to the GPU box, /root/reference does not.

  ref_checkout_*  : the checkout exactly as it is (nlp22 branch): Embedding input, softmax weights passed
                    as logits to tf.random.categorical, sparse-CE loss.  No patch of any kind.
  ref_regression_*: the same reference cell / attention / model code with the three documented
                    regression-era substitutions (SURVEY.md §0 finding 1, §8 a9): Dense
                    'projection_layer_ts' encoding (pyc), Gaussian observation weight compute_w_regression
                    (called at SMC_TransformerCell.py:206 but absent from the checkout; north-star form)
                    and the pyc MSE-type classic loss.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tf_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run_reference(B, P, S, D, dff, H, V, attn_window, sigma, multi, seed, regression=None, len_resampling=None):
    inst = tf_shim.install()
    tf, rng = inst.tf, inst.rng
    rng.np_rng = np.random.default_rng(seed)
    from src.models.SMC_Transformer.SMC_Transformer import SMC_Transformer
    model = SMC_Transformer(d_model=D, output_size=V, seq_len=S, full_model=True, dff=dff, num_layers=1, num_heads=H,
                            attn_window=attn_window)
    sig = [sigma * (1 + 0.1 * i) for i in range(D)] if multi else sigma
    model.cell.add_SMC_parameters(dict_sigmas=dict(zip(["k", "q", "v", "z"], [sig] * 4)), num_particles=P)
    if len_resampling is not None:
        model.cell.add_stop_resampling(len_resampling)
    g = np.random.default_rng(seed + 1)
    eps = g.standard_normal((4, S, B, P, D)).astype(np.float32)
    u = g.random((S, B, P))
    if regression is None:
        x_in = g.integers(0, V, size=(B, 1, S, 1)).astype(np.int64)
        y = g.integers(0, V, size=(B, 1, S, 1)).astype(np.int64)
    else:
        F_x, F_y = regression["F_x"], V
        x_in = g.standard_normal((B, 1, S, F_x)).astype(np.float32)
        y = (0.5 * g.standard_normal((B, 1, S, F_y))).astype(np.float32)
        sigma_obs = regression["sigma_obs"]
        proj = tf.keras.layers.Dense(D, name="projection_layer_ts")

        def get_encoded_input(inputs, attention_mask=None):   # pyc: Dense(d_model)(x) * sqrt(d_model), after tiling
            if tf.shape(inputs)[1] == 1:
                inputs = tf.tile(inputs, multiples=[1, model.cell.num_particles, 1, 1])
            return proj(inputs) * tf.math.sqrt(tf.cast(model.d_model, tf.float32))

        def compute_w_regression(predictions, y):             # north-star Gaussian weight, log form, normalised
            diff = tf.squeeze(y - predictions, axis=-2)       # (B,P,F)
            logw = -0.5 * tf.einsum("bpf,bpf->bp", diff, diff) / sigma_obs
            return logw - torch.logsumexp(logw, dim=1, keepdim=True)
        model.get_encoded_input = get_encoded_input
        model.cell.compute_w_classification = compute_w_regression
    # injected draws, consumed in call order; K.rnn's dummy first call consumes one set (it re-uses t = 0)
    cnt = {"n": 0, "c": 0}

    def normal_fn(shape):
        which, call = cnt["n"] % 4, cnt["n"] // 4
        cnt["n"] += 1
        return torch.from_numpy(eps[which, max(call - 1, 0)]).reshape(shape)

    def uniform_fn(Bn, ns):
        t = max(cnt["c"] - 1, 0)
        cnt["c"] += 1
        return torch.from_numpy(u[t])
    rng.normal_fn, rng.uniform_fn = normal_fn, uniform_fn
    (pred, pred_r), (K, Vv, R), w = model(inputs=torch.from_numpy(x_in), targets=torch.from_numpy(y), attention_mask=None)
    assert cnt["n"] == 4 * (S + 1) and cnt["c"] == S + 1
    if regression is None:
        total, classic = model.compute_SMC_loss(inputs=None, targets=torch.from_numpy(y), predictions=pred_r)
    else:   # pyc: mean 0.5*sum_n ||n||^2/sigma_n^2 + mean 0.5*||y-pred||^2/Sigma_obs
        att = model.cell.attention_smc
        parts = []
        for noise, lv in zip(model.internal_noises, [att.logvar_k, att.logvar_q, att.logvar_v, att.logvar_z]):
            var = torch.exp(lv.value) if lv.value.dim() == 0 else torch.exp(torch.diagonal(lv.value))
            parts.append(0.5 * (noise * noise / var).sum(-1))
        classic = (0.5 / regression["sigma_obs"]) * ((torch.from_numpy(y) - pred_r) ** 2).sum(-1).mean()
        total = torch.stack(parts).sum(0).mean() + classic
    att = model.cell.attention_smc
    c = model.cell

    def lv(v):
        v = v.value
        return (torch.diagonal(v) if v.dim() == 2 else v).numpy().astype(np.float32)
    params = dict(Wq=att.wq.kernel.numpy(), bq=att.wq.bias.numpy(), Wk=att.wk.kernel.numpy(), bk=att.wk.bias.numpy(),
                  Wv=att.wv.kernel.numpy(), bv=att.wv.bias.numpy(), Wz=att.dense.kernel.numpy(), bz=att.dense.bias.numpy(),
                  ln1_g=c.layernorm1.gamma.numpy(), ln1_b=c.layernorm1.beta.numpy(),
                  ln2_g=c.layernorm2.gamma.numpy(), ln2_b=c.layernorm2.beta.numpy(),
                  W1=c.ffn.layers[0].kernel.numpy(), b1=c.ffn.layers[0].bias.numpy(),
                  W2=c.ffn.layers[1].kernel.numpy(), b2=c.ffn.layers[1].bias.numpy(), Wo=c.output_layer.kernel.numpy(),
                  logvar_k=lv(att.logvar_k), logvar_q=lv(att.logvar_q), logvar_v=lv(att.logvar_v), logvar_z=lv(att.logvar_z))
    if regression is None:
        params["W_in"], params["b_in"] = model.embedding.embeddings.numpy(), np.zeros(D, np.float32)
    else:
        params["W_in"], params["b_in"] = proj.kernel.numpy(), proj.bias.numpy()
    return dict(params=params, x_in=x_in, y=y, eps=eps, u=u, pred=pred.numpy(), pred_resampl=pred_r.numpy(), K=K.numpy(),
                V=Vv.numpy(), R=R.numpy(), w=w.numpy(), noise_q=model.noise_q.numpy(), noise_z=model.noise_z.numpy(),
                noise_K=model.noise_K_resampled.numpy(), noise_V=model.noise_V_resampled.numpy(),
                loss_total=float(total), loss_classic=float(classic))


SPECS = {
    # name: kwargs.  ref_checkout_main mirrors the checkout's own smoke block (SMC_Transformer.py:252-259):
    # b=8, S=10, d_model=6... (vocabulary 50 here), P=4, attention window 4.
    "ref_checkout_main": dict(B=8, P=4, S=10, D=6, dff=24, H=1, V=50, attn_window=4, sigma=0.1, multi=False, seed=11),
    "ref_checkout_heads": dict(B=3, P=10, S=7, D=12, dff=24, H=3, V=30, attn_window=None, sigma=0.5, multi=True, seed=12),
    # regression-era smoke (pyc __main__): b=8, S=10, F=1, d_model=6, dff=24, attn_window=4, P=20, sigma=0.1, sigma_obs=0.5
    "ref_regression_pyc_smoke": dict(B=8, P=20, S=10, D=6, dff=24, H=1, V=1, attn_window=4, sigma=0.1, multi=False, seed=13,
                                     regression=dict(F_x=1, sigma_obs=0.5)),
    "ref_regression_c3": dict(B=2, P=60, S=24, D=32, dff=32, H=1, V=5, attn_window=None, sigma=0.5, multi=False, seed=14,
                              regression=dict(F_x=5, sigma_obs=0.5)),
    "ref_regression_d64_stop": dict(B=1, P=40, S=12, D=64, dff=256, H=1, V=2, attn_window=None, sigma=0.5, multi=False, seed=15,
                                    regression=dict(F_x=3, sigma_obs=0.5), len_resampling=6),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, kw in SPECS.items():
        r = run_reference(**kw)
        meta = {k: v for k, v in kw.items() if k != "regression"}
        arrays = dict(name=np.array(name), meta=np.array(repr(kw)), x_in=r["x_in"], y=r["y"], eps=r["eps"], u=r["u"],
                      pred=r["pred"], pred_resampl=r["pred_resampl"], w=r["w"], K=r["K"], R=r["R"],
                      V_last=r["V"][:, :, -1], noise_q=r["noise_q"], noise_z=r["noise_z"], noise_K=r["noise_K"],
                      noise_V_last=r["noise_V"][:, :, -1],
                      loss_total=np.array(r["loss_total"]), loss_classic=np.array(r["loss_classic"]))
        arrays.update({"p_" + k: v for k, v in r["params"].items()})
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(path, os.path.getsize(path) // 1024, "KiB", "loss", r["loss_total"], meta)


if __name__ == "__main__":
    main()
