"""CPU tests of the oracle itself: reference KATs, golden fixtures, internal consistency."""
import glob
import os

import numpy as np
import pytest

from oracle import philox as PH
from oracle import smct_oracle as O
from tests import util_cpu
from tests.cases import CASES

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_resample_kat_tables():
    """src/models/SMC_Transformer/transformer_utils.py:66-88 (B=2,P=4,S=3,D=1)."""
    ind = np.array([[[1, 1, 2, 2], [0, 0, 0, 0], [1, 1, 1, 0]], [[0, 1, 3, 2], [3, 3, 2, 0], [1, 2, 3, 1]]]).transpose(0, 2, 1)
    K = np.array([[[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]],
                  [[13, 14, 15, 16], [17, 18, 19, 20], [21, 22, 23, 24]]]).transpose(0, 2, 1)[..., None].astype(np.float32)
    truth = {0: [[[2, 5, 9], [2, 6, 10], [3, 7, 11], [3, 8, 12]], [[13, 17, 21], [14, 18, 22], [16, 19, 23], [15, 20, 24]]],
             1: [[[2, 5, 9], [2, 5, 10], [2, 5, 11], [2, 5, 12]], [[15, 20, 21], [15, 20, 22], [16, 19, 23], [13, 17, 24]]],
             2: [[[2, 5, 10], [2, 5, 10], [2, 5, 10], [2, 5, 9]], [[15, 20, 22], [16, 19, 23], [13, 17, 24], [15, 20, 22]]]}
    cur = K
    for t in range(3):
        cur = O.resample_rows_le_t(cur, ind[:, :, t], t)
        assert np.array_equal(cur[..., 0], np.array(truth[t], np.float32))


def test_reference_squared_norm_kat():
    """src/models/SMC_Transformer/SMC_Transformer.py:361-365."""
    x = np.array([[.1, 1], [.2, 2], [.3, 3], [.4, 4], [.5, 5]], np.float32)[None, None]
    assert np.allclose(np.einsum("bijk,bijk->bij", x, x)[0, 0], [1.01, 4.04, 9.09, 16.16, 25.25], rtol=1e-6)
    # the gaussian log-weight uses exactly this reduction
    logw, _, _ = O.compute_w_regression(np.zeros((1, 5, 2), np.float32), -x[0, 0][None, 0], 0.5)
    assert np.allclose(logw[0], -1.01 * np.ones(5), rtol=1e-6)


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    assert [int(v) for v in PH.philox4x32_10(0, 0, 0, 0, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert [int(v) for v in PH.philox4x32_10(f, f, f, f, f, f)] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(v) for v in PH.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_categorical_matches_sequential_definition():
    rng = np.random.default_rng(0)
    logits = (2 * rng.standard_normal((5, 37))).astype(np.float32)
    u = rng.random((5, 37))
    got = O.categorical_tf_cpu(logits, u)
    for b in range(5):
        mx = float(logits[b].max())
        run, cdf = 0.0, []
        for j in range(37):
            run += np.exp(np.float64(logits[b, j]) - mx)
            cdf.append(run)
        for m in range(37):
            tgt = u[b, m] * cdf[-1]
            idx = next((j for j, c in enumerate(cdf) if c > tgt), 36)
            assert got[b, m] == idx


def test_full_row_gather_equals_rows_le_t_gather():
    """rows > t are zero for every particle (SMC_Transformer.py:197-202), so the reference's full-row
    tf.gather and the live-row gather give identical states."""
    c = CASES["c3_b8"]
    cfg = c["cfg"]
    p, x_in, y, eps, u = util_cpu.make_case(cfg, sigma2=c["sigma2"])
    a = O.forward(cfg, p, x_in, y, eps, u, fast_gather=True)
    b = O.forward(cfg, p, x_in, y, eps, u, fast_gather=False)
    for k in ("K", "V", "R", "pred", "pred_resampl", "w", "i_t"):
        assert np.array_equal(a[k], b[k]), k


def test_quirks_are_reproduced():
    c = CASES["c2_b4"]
    cfg = c["cfg"]
    p, x_in, y, eps, u = util_cpu.make_case(cfg, sigma2=c["sigma2"])
    out = O.forward(cfg, p, x_in, y, eps, u, trace=True)
    st = out["steps"][3]
    # noise_z = z - z_ (pre-projection), self_attention_SMC.py:190
    assert np.allclose(out["noise_z"][:, :, 3], st["z"] - st["z_"], atol=1e-6)
    # noise_K_resampled uses the un-normalised X, SMC_Transformer.py:235
    assert np.allclose(out["noise_K_resampled"], out["K"] - O.dense(out["X"], p["Wk"], p["bk"])[:, None], atol=1e-6)
    # weights are normalised over particles
    assert np.allclose(out["w"].sum(1), 1, atol=1e-5)
    # the last slot of K is never resampled away from what was written at t = S-1 other than by i_{S-1}
    i_last = out["i_t"][-1]
    assert np.allclose(out["K"][:, :, -1], np.take_along_axis(out["steps"][-1]["k"], i_last[..., None], 1))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "oracle_*.npz"))))
def test_oracle_reproduces_golden(path):
    g = np.load(path, allow_pickle=True)
    c = CASES[str(g["case"])]
    cfg = c["cfg"]
    p = {k[2:]: g[k] for k in g.files if k.startswith("p_")}
    out = O.forward(cfg, p, g["x_in"], g["y"], g["eps"], g["u"])
    assert np.array_equal(out["i_t"], g["i_t"])
    for k in ("pred", "pred_resampl", "w"):
        assert np.allclose(out[k], g[k], rtol=1e-5, atol=1e-6), k
    total, classic = O.smc_loss(cfg, p, out, g["y"], "checkout")
    assert np.isclose(total, float(g["loss_checkout"]), rtol=1e-5)
    assert np.isclose(classic, float(g["loss_classic"]), rtol=1e-5)


def test_layer_norm_matches_torch():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(1)
    x = rng.standard_normal((7, 33)).astype(np.float32)
    g = rng.standard_normal(33).astype(np.float32)
    b = rng.standard_normal(33).astype(np.float32)
    ref = torch.nn.functional.layer_norm(torch.from_numpy(x), (33,), torch.from_numpy(g), torch.from_numpy(b), 1e-6).numpy()
    assert np.allclose(O.layer_norm(x, g, b, 1e-6), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["c3_b8", "heads3", "window2", "pyc_smoke"])
def test_eager_torch_restatement_agrees_with_numpy_oracle(name):
    """Two independently written restatements of the same reference lines (numpy oracle vs the
    op-for-op torch-CPU port used as the CPU baseline) must agree on injected draws."""
    torch = pytest.importorskip("torch")
    from oracle import ref_eager_torch as E
    c = CASES[name]
    cfg = c["cfg"]
    p, x_in, y, eps, u = util_cpu.make_case(cfg, seed=4, sigma2=c["sigma2"])
    u, ora = util_cpu.screen_uniforms(cfg, p, x_in, y, eps, u, margin=1e-5)
    tp = {k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in p.items()}
    out = E.forward(cfg, tp, torch.from_numpy(x_in), torch.from_numpy(y), torch.from_numpy(eps), torch.from_numpy(u))
    assert np.array_equal(out["i_t"].numpy(), ora["i_t"])
    for k in ("pred", "pred_resampl", "K", "V", "R", "w", "noise_q", "noise_z", "noise_K_resampled"):
        assert np.allclose(out[k].numpy(), ora[k], rtol=2e-4, atol=2e-5), k
    assert np.isclose(E.loss_pyc(cfg, tp, out, torch.from_numpy(y)), O.smc_loss(cfg, p, ora, y, "pyc")[0], rtol=1e-4)


def test_eager_torch_restatement_classification_variant():
    """The nlp checkout's variant (Embedding input, softmax(pred)[y] weights passed as logits, sparse cross-entropy): the
    torch restatement (gradient oracle of the backward tests) against the numpy oracle, forward and loss."""
    torch = pytest.importorskip("torch")
    from oracle import ref_eager_torch as E
    c = CASES["classification"]
    cfg = c["cfg"]
    p, x_in, y, eps, u = util_cpu.make_case(cfg, seed=4, sigma2=c["sigma2"])
    ora = O.forward(cfg, p, x_in, y, eps, u)
    tp = {k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in p.items()}
    out = E.forward(cfg, tp, torch.from_numpy(x_in), torch.from_numpy(y), torch.from_numpy(eps), i_forced=torch.from_numpy(ora["i_t"]))
    for k in ("pred", "pred_resampl", "K", "V", "R", "w", "noise_q", "noise_z", "noise_K_resampled"):
        assert np.allclose(out[k].numpy(), ora[k], rtol=2e-4, atol=2e-5), k
    loss, grads = E.loss_and_grads(cfg, p, x_in, y, eps, ora["i_t"], variant="checkout")
    assert np.isclose(loss, O.smc_loss(cfg, p, ora, y, "checkout")[0], rtol=1e-5)
    assert all(np.isfinite(g).all() for g in grads.values()) and np.abs(grads["W_in"]).max() > 0


# ------------------------------------------------------------------------------------------------
# the oracle against the outputs of the reference's OWN Python sources (run over oracle/tf_shim.py)
# ------------------------------------------------------------------------------------------------
from tests import ref_golden  # noqa: E402


def _close(name, got, ref, rtol=1e-4):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    scale = np.abs(ref).max()
    err = np.abs(got - ref)
    assert (err <= rtol * np.abs(ref) + 0.1 * rtol * scale + 1e-30).all(), "%s: max err %g (scale %g)" % (name, err.max(), scale)


@pytest.mark.parametrize("path", ref_golden.ref_files())
def test_oracle_matches_reference_sources(path):
    """numpy restatement == the reference's unmodified SMC_Transformer / cell / attention code (executed
    over the TF shim) on the same weights and injected draws — including every quirk of the checkout."""
    g, cfg, p, x_in, y = ref_golden.load(path)
    out = O.forward(cfg, p, x_in, y, g["eps"], g["u"])
    out["w_as_returned_by_reference"] = out["w"] if cfg.weights == "classification" else out["logits"].transpose(1, 2, 0)
    ref_golden.compare(os.path.basename(path), out, g, _close)
    variant = "checkout" if cfg.weights == "classification" else "pyc"
    total, classic = O.smc_loss(cfg, p, out, y, variant)
    assert np.isclose(total, float(g["loss_total"]), rtol=1e-4), (total, float(g["loss_total"]))
    assert np.isclose(classic, float(g["loss_classic"]), rtol=1e-4)


def test_reference_fixtures_exist():
    assert len(ref_golden.ref_files()) >= 5


def test_gradient_oracle_matches_reference_tape_gradient():
    """The autograd restatement (oracle/ref_eager_torch.loss_and_grads) that the backward tests use reproduces
    tape.gradient(loss, trainable_variables) of the reference's own forward and loss code (tests/golden/ref_grad_*.npz,
    written over the TF shim by oracle/gen_golden_grads_from_reference.py)."""
    from oracle import ref_eager_torch as E
    from tests import ref_grad_golden as RG
    paths = RG.files()
    assert len(paths) >= 7
    for path in paths:
        g, cfg, p, ref, x_in, y, variant = RG.load(path)
        ora = O.forward(cfg, p, x_in, y, g["eps"], g["u"])
        mask = g["mask"] if "mask" in g.files else None
        _, grads = E.loss_and_grads(cfg, p, x_in, y, g["eps"], ora["i_t"], variant=variant, mask=mask)
        RG.compare_grads(path, grads, ref)


def test_philox_round_count_matches_the_device_header():
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(__file__)), "smc-t-v2_b200", "csrc", "smct_common.cuh")).read()
    assert int(re.search(r"SMCT_PHILOX_ROUNDS\s*=\s*(\d+)", src).group(1)) == PH.ROUNDS
    # the first ROUNDS rounds of the generator whose 10-round form is pinned by the Random123 vectors above
    a = PH.philox4x32(1, 2, 3, 4, 5, 6)
    b = PH.philox4x32(1, 2, 3, 4, 5, 6, rounds=PH.ROUNDS)
    assert [int(v) for v in a] == [int(v) for v in b] != [int(v) for v in PH.philox4x32_10(1, 2, 3, 4, 5, 6)]
