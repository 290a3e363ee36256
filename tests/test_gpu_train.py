"""The `run.py -algo smc_t` switch: SMCTAlgo(dataset, args).train() — mirror of /root/reference/src/algos/run_SMC_T.py:16-47,
121-153, src/train/train_functions.py:146-243, src/scripts/run.py:128-135 — plus the training-step corner cases
(EM with an optimiser, classification inputs)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class TinyDataset:
    """A Dataset-like object with the reference's protocol (data_provider/class_datasets.py:10-93): get_datasets() ->
    (train, val, test) arrays; data_to_dataset(...) -> iterables of (inputs, targets, attention_mask) batches."""
    name = "synthetic"

    def __init__(self, B=8, S=12, F=1, n_train=4, n_val=2, classification=False, vocab=12):
        from smct_b200 import utils
        self.seq_len, self.B, self.F = S, B, F
        self.classification = classification
        self.output_size = vocab if classification else F
        rng = np.random.default_rng(0)
        if classification:
            self.data = rng.integers(1, vocab, size=((n_train + n_val + 1) * B, S + 1, 1)).astype(np.int32)
        else:
            x, y = utils.synthetic_ar1((n_train + n_val + 1) * B, S, F, alpha=0.8, var=0.5)
            self.data = np.concatenate([x, y[:, -1:]], axis=1)
        self.n_train, self.n_val = n_train, n_val

    def get_datasets(self):
        B = self.B
        a, b = self.n_train * B, (self.n_train + self.n_val) * B
        return self.data[:a], self.data[a:b], self.data[b:]

    def _batches(self, data):
        out = []
        for i in range(0, len(data), self.B):
            d = data[i:i + self.B]
            inp, tar = d[:, None, :-1, :], d[:, None, 1:, :]       # (B,1,S,F): split of class_datasets.py:19-22
            out.append((inp, tar, None))
        return out

    def data_to_dataset(self, train_data, val_data, test_data, num_dim=4):
        assert num_dim == 4
        return self._batches(train_data), self._batches(val_data), self._batches(test_data)


def test_smct_algo_train_two_epochs(tmp_path):
    """algos["smc_t"](dataset, args).train(): 2 epochs, history / CSV / weights written, the loss drops."""
    from smct_b200.algos.run_SMC_T import algos, default_args
    ds = TinyDataset()
    args = default_args(d_model=8, dff=8, particles=10, sigmas=0.5, ep=2, bs=8, output_path=str(tmp_path))
    algo = algos["smc_t"](dataset=ds, args=args)
    assert algo.optimizer is not None and algo.EPOCHS == 2 and algo.seq_len == 12
    algo.smc_transformer.seed = 11                      # same draws in every step: the loss is a function of the weights only
    algo.optimizer.learning_rate = 5e-3                 # CustomSchedule starts at ~1e-6: too slow to see in 8 steps
    w0 = algo.smc_transformer.cell.ffn.dense1.kernel.numpy().copy()
    hist = algo.train()
    assert len(hist["train_loss"]) == 2 and len(hist["val_loss"]) == 2 and len(hist["train_mse_metric"]) == 2
    assert np.all(np.isfinite(hist["train_loss"])) and np.all(np.isfinite(hist["val_loss"]))
    assert hist["train_loss"][1] < hist["train_loss"][0], hist
    assert not np.array_equal(algo.smc_transformer.cell.ffn.dense1.kernel.numpy(), w0)
    assert algo.optimizer.iterations == 8
    out = algo.out_folder
    assert os.path.exists(os.path.join(out, "model.npz")) and os.path.exists(os.path.join(out, "smc_t_history_1.csv"))
    assert os.path.exists(os.path.join(out, "config.json")) and os.path.exists(os.path.join(out, "logvar_after_training.json"))
    with open(os.path.join(out, "logvar_after_training.json")) as f:
        assert set(json.load(f)) == {"k", "q", "v", "z"}
    with pytest.raises(NotImplementedError):
        algo.test(test_samples=None, temp=1.)


def test_smct_algo_default_schedule_runs():
    """The reference's own optimiser wiring (Adam(CustomSchedule(d_model), 0.9, 0.98, 1e-9)): first update has lr(0) = 0."""
    from smct_b200.algos.run_SMC_T import SMCTAlgo, default_args
    from smct_b200.train.optimizers import CustomSchedule
    assert CustomSchedule(8)(0) == 0.0
    assert abs(CustomSchedule(8)(4000) - 8 ** -0.5 * 4000 ** -0.5) < 1e-12
    ds = TinyDataset(n_train=2, n_val=1)
    algo = SMCTAlgo(dataset=ds, args=default_args(d_model=8, dff=8, particles=5, sigmas=0.1, ep=1, bs=8))
    w0 = algo.smc_transformer.cell.ffn.dense1.kernel.numpy().copy()
    from smct_b200.train.train_step_functions import train_step_SMC_T
    inp, tar, _ = algo.train_dataset[0]
    train_step_SMC_T(inp, tar, algo.smc_transformer, algo.optimizer, it=1)
    assert np.array_equal(algo.smc_transformer.cell.ffn.dense1.kernel.numpy(), w0)      # lr(0) = 0, like tf.keras
    train_step_SMC_T(inp, tar, algo.smc_transformer, algo.optimizer, it=2)
    assert not np.array_equal(algo.smc_transformer.cell.ffn.dense1.kernel.numpy(), w0)
    assert algo.out_folder is None


def test_train_step_with_EM_and_optimizer():
    """EM mode: the log-variances are not trainable (Adam must not touch them), EM overwrites them before the loss, and
    the backward replays the forward with the variances the forward used."""
    from smct_b200.algos.run_SMC_T import SMCTAlgo, default_args
    from smct_b200.train.optimizers import Adam
    from smct_b200.train.train_step_functions import compute_gradients, train_step_SMC_T
    ds = TinyDataset(n_train=1, n_val=1)
    algo = SMCTAlgo(dataset=ds, args=default_args(d_model=8, dff=8, particles=6, sigmas=0.5, EM_param=0.6, ep=1, bs=8))
    m = algo.smc_transformer
    att = m.cell.attention_smc
    assert not att.logvar_k.trainable and all(v.name[:6] != "logvar" for v in m.trainable_variables)
    m.seed = 3
    inp, tar, _ = ds._batches(ds.data[:8])[0]
    opt = Adam(1e-2)
    lv0 = {n: float(getattr(att, "logvar_" + n).numpy()) for n in "kqvz"}
    loss, _ = train_step_SMC_T(inp, tar, m, opt, it=2, EM_param=0.6)
    lv1 = {n: float(getattr(att, "logvar_" + n).numpy()) for n in "kqvz"}
    assert all(lv1[n] != lv0[n] for n in "kqvz")                       # moved by EM ...
    # ... and by EM only: redo the EM update by hand from the recorded noises
    from smct_b200.train.utils import EM_update
    for n, noise in (("k", m.noise_K_resampled), ("q", m.noise_q), ("v", m.noise_V_resampled), ("z", m.noise_z)):
        cur = lv0[n]
        for j in range(noise.shape[0]):
            cur = EM_update(float(noise[j].double().square().mean()), cur, 2, 0.6)
        assert abs(cur - lv1[n]) < 1e-5 * max(1.0, abs(cur)), (n, cur, lv1[n])
    assert len(opt._slots) == len(m.trainable_variables)
    # the gradient belongs to the forward that was run: the q / z noises are regenerated with the forward's variances
    fwd = m._last_forward
    assert abs(float(fwd["logvar_fwd"]["logvar_q"]) - lv0["q"]) < 1e-7
    g = compute_gradients(m, inp, tar)
    assert all(torch.isfinite(v).all() for v in g.values())


def test_classification_train_step():
    """The nlp checkout's inputs (integer tokens): training needs the regression-era noise_z; the error comes BEFORE the
    forward, and with noise_z_mode = "pyc" the step trains."""
    from smct_b200.algos.run_SMC_T import SMCTAlgo, default_args
    from smct_b200.train.optimizers import Adam
    from smct_b200.train.train_step_functions import train_step_SMC_T
    ds = TinyDataset(classification=True, n_train=1, n_val=1)
    algo = SMCTAlgo(dataset=ds, args=default_args(d_model=16, dff=16, particles=6, sigmas=0.1, ep=1, bs=8))
    m = algo.smc_transformer
    inp, tar, _ = algo.train_dataset[0]
    opt = Adam(1e-2)
    m._last_forward = None
    with pytest.raises(NotImplementedError, match="noise_z_mode"):
        train_step_SMC_T(inp, tar, m, opt, it=1)
    assert m._last_forward is None                                     # no forward was run
    m.noise_z_mode = "pyc"
    m.seed = 5
    losses = [float(train_step_SMC_T(inp, tar, m, opt, it=i + 1)[0]) for i in range(6)]
    assert np.all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def test_forecast_multistep_matches_oracle_rollout():
    """inference_multistep (incremental roll-out on the particles' K/V cache, smct_cell_step per future step) against
    the oracle's restatement of the regression-era forecast (oracle/smct_oracle.py:forecast_multistep; reference
    src/eval/inference_functions.py:19-39 + the shipped pyc) on injected draws."""
    from oracle import smct_oracle as O
    from smct_b200 import utils
    from smct_b200.eval.inference_functions import inference_multistep
    from smct_b200.models.SMC_Transformer.SMC_Transformer import SMC_Transformer
    from tests import util
    B, P, D, dff, past_len, future_len = 3, 12, 16, 16, 9, 6
    S = past_len + future_len
    x, y = utils.synthetic_ar1(B, past_len, 1, alpha=0.8, var=0.5)
    model = SMC_Transformer(d_model=D, output_size=1, seq_len=S, full_model=True, dff=dff)
    model.cell.add_SMC_parameters(dict_sigmas=dict(k=0.1, q=0.2, v=0.1, z=0.3), num_particles=P, sigma_obs=0.4)
    inputs, targets = torch.from_numpy(x).unsqueeze(1), torch.from_numpy(y).unsqueeze(1)
    rng = np.random.default_rng(5)
    eps = rng.standard_normal((4, S, B, P, D)).astype(np.float32)
    u = rng.random((S, B, P))
    obs = rng.standard_normal((S, B, P, 1)).astype(np.float32)
    params = {k: v.detach().cpu().numpy() for k, v in model._params("regression", 1).items()}
    cfg = O.Config(B=B, P=P, S=S, D=D, dff=dff, F_x=1, F_y=1, sigma_obs=0.4, noise_z_mode="pyc", len_resampling=past_len)
    # screen the uniforms of the observed past against the oracle's own CDFs (free-running protocol L4)
    for _ in range(50):
        ref = O.forecast_multistep(cfg, params, x, y, eps, u, obs, past_len, future_len)
        pf, mean = inference_multistep(model, inputs, targets, past_len=past_len, future_len=future_len, eps=eps, u=u, obs_noise=obs)
        got = pf.cpu().numpy()
        if np.allclose(got, ref, rtol=1e-3, atol=1e-3):
            break
        u = rng.random((S, B, P))          # a draw sat on a CDF boundary: the trajectories forked; redraw
    assert got.shape == ref.shape == (B, P, future_len, 1)
    util.assert_close("forecast preds_future", got, ref)
    util.assert_close("forecast mean", mean.cpu().numpy(), ref.mean(axis=1))


def test_train_per_dimension_variances():
    """-noise_dim multi (the reference's published ROC runs): the (D,D) diagonal log-variance variables train — Adam moves
    their diagonals, keeps its moments between steps, and leaves the off-diagonal entries alone."""
    from smct_b200.algos.run_SMC_T import SMCTAlgo, default_args
    from smct_b200.train.optimizers import Adam
    from smct_b200.train.train_step_functions import train_step_SMC_T
    ds = TinyDataset(n_train=1, n_val=1)
    algo = SMCTAlgo(dataset=ds, args=default_args(d_model=8, dff=8, particles=6, sigmas=0.5, noise_dim="multi", ep=1, bs=8))
    m = algo.smc_transformer
    att = m.cell.attention_smc
    assert tuple(att.logvar_k.shape) == (8, 8) and att.logvar_dim() == 8
    m.seed = 4
    inp, tar, _ = algo.train_dataset[0]
    opt = Adam(1e-2)
    lv0 = att.logvar_k.numpy().copy()
    losses = [float(train_step_SMC_T(inp, tar, m, opt, it=i + 1)[0]) for i in range(5)]
    n_slots = len(opt._slots)
    losses.append(float(train_step_SMC_T(inp, tar, m, opt, it=6)[0]))
    assert len(opt._slots) == n_slots == len(m.trainable_variables)          # moments are kept, not re-created
    lv1 = att.logvar_k.numpy()
    off = ~np.eye(8, dtype=bool)
    assert np.array_equal(lv1[off], lv0[off]) and np.all(np.diagonal(lv1) != np.diagonal(lv0))
    assert np.all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def test_train_step_with_attention_mask():
    """(inputs, targets, attention_mask) batches of the NLP datasets: the mask weights the classic loss, its gradient and
    the metric (SMC_Transformer.py:158-163, train/utils.py:15-19); a fully padded-out row gets no classic-loss gradient."""
    from smct_b200.algos.run_SMC_T import SMCTAlgo, default_args
    from smct_b200.train.optimizers import Adam
    from smct_b200.train.train_step_functions import compute_gradients, train_step_SMC_T
    ds = TinyDataset(classification=True, n_train=1, n_val=1)
    algo = SMCTAlgo(dataset=ds, args=default_args(d_model=16, dff=16, particles=6, sigmas=0.1, ep=1, bs=8))
    m = algo.smc_transformer
    m.noise_z_mode, m.seed = "pyc", 5
    inp, tar, _ = algo.train_dataset[0]
    B, S = inp.shape[0], inp.shape[2]
    mask = np.ones((B, 1, S, 1), np.int32)
    mask[:, :, S - 3:, :] = 0
    loss_m, metric_m = train_step_SMC_T(inp, tar, m, None, it=1, attention_mask=mask)
    loss_u, metric_u = train_step_SMC_T(inp, tar, m, None, it=1)
    assert np.isfinite(float(loss_m)) and float(loss_m) != float(loss_u) and float(metric_m) != float(metric_u)
    # classic loss by hand from the returned logits
    (pred, pred_r), _, _ = m(inp, tar, attention_mask=mask)
    lp = torch.log_softmax(pred_r.double(), -1)
    y = torch.as_tensor(tar).reshape(B, 1, S, 1).to(lp.device, torch.int64).expand(B, pred_r.shape[1], S, 1)
    ce = -torch.gather(lp, -1, y).squeeze(-1)
    mm = torch.as_tensor(mask, device=lp.device).double().reshape(B, 1, S)
    want = float((ce * mm).sum() / (mm.sum() * pred_r.shape[1]))
    assert abs(float(m.compute_classic_loss(tar, pred_r, attention_mask=mask)) - want) < 1e-5 * abs(want)
    opt = Adam(1e-2)
    losses = [float(train_step_SMC_T(inp, tar, m, opt, it=i + 1, attention_mask=mask)[0]) for i in range(5)]
    assert np.all(np.isfinite(losses)) and losses[-1] < losses[0], losses


@pytest.mark.parametrize("masked", [False, True])
def test_metric_kernel_matches_torch(masked):
    """smct_metric (the training metric, src/train/utils.py:3-20 / the regression-era MSE of the particle mean) against a
    torch fp64 restatement."""
    from smct_b200.train.utils import compute_categorical_cross_entropy, compute_mse_metric
    rng = np.random.default_rng(3)
    B, P, S, V = 5, 7, 9, 33
    mask = (rng.random((B, 1, S, 1)) > 0.3).astype(np.int32) if masked else None
    m = None if mask is None else torch.as_tensor(mask).double().reshape(B, S)
    # classification
    pred = torch.as_tensor(rng.standard_normal((B, P, S, V)).astype(np.float32) * 3, device="cuda")
    tar = rng.integers(0, V, size=(B, 1, S, 1)).astype(np.int32)
    got = float(compute_categorical_cross_entropy(tar, pred, P, attention_mask=mask))
    probs = torch.softmax(pred.double().cpu(), -1).mean(1)                                    # (B,S,V)
    ce = -torch.log(torch.gather(probs, -1, torch.as_tensor(tar).long().reshape(B, S, 1)).squeeze(-1).clamp_min(1e-7))
    want = float(ce.mean()) if m is None else float((ce * m).sum() / m.sum())
    assert abs(got - want) < 1e-5 * abs(want), (got, want)
    # regression
    Fy = 3
    pred = torch.as_tensor(rng.standard_normal((B, P, S, Fy)).astype(np.float32), device="cuda")
    tar = rng.standard_normal((B, 1, S, Fy)).astype(np.float32)
    got = float(compute_mse_metric(tar, pred, attention_mask=mask))
    se = (pred.double().cpu().mean(1) - torch.as_tensor(tar).double().reshape(B, S, Fy)).square()
    want = float(se.mean()) if m is None else float((se * m.unsqueeze(-1)).sum() / (m.sum() * Fy))
    assert abs(got - want) < 1e-5 * abs(want), (got, want)
