"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes (no GPU, no compute kernels).
Checks the shard partition, the global-index keying and the loss-sum reduction."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_the_batch():
    import smct_b200  # noqa: F401
    from smct_b200.distributed import shard_range
    for B in (1, 7, 32, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                b0, b1 = shard_range(B, r, world)
                assert 0 <= b0 <= b1 <= B
                cover += list(range(b0, b1))
            assert cover == list(range(B))
            sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import smct_b200  # noqa: F401
    from smct_b200.distributed import sharded_forward, shard_range
    from oracle import philox as PH
    x = torch.arange(B * 3, dtype=torch.float32).reshape(B, 3)
    y = torch.ones(B, 2)
    seen = {}

    def fake_forward(xs, ys, b_global0, seed):
        # stands in for the CUDA forward: "loss sums" that depend on the global sequence index only,
        # through the same Philox stream keying the device uses (uniform stream, t=0, one particle)
        n = xs.shape[0]
        seen["b0"], seen["n"] = b_global0, n
        u = PH.uniforms(seed, 0, b_global0, n, 1)[:, 0]
        s = torch.zeros(8, dtype=torch.float64)
        s[0] = float(u.sum())
        s[4] = float(xs.sum())
        s[7] = n
        return s

    sums = sharded_forward(fake_forward, x, y, B, seed=123)
    b0, b1 = shard_range(B, rank, world)
    assert seen["b0"] == b0 and seen["n"] == b1 - b0
    np.save(os.path.join(out_dir, "sums_%d.npy" % rank), sums.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_sharded_forward(tmp_path):
    from oracle import philox as PH
    B, world = 13, 2
    mp.spawn(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    s0 = np.load(tmp_path / "sums_0.npy")
    s1 = np.load(tmp_path / "sums_1.npy")
    assert np.array_equal(s0, s1)                                  # all-reduced: identical on every rank
    ref_u = PH.uniforms(123, 0, 0, B, 1)[:, 0].sum()
    assert np.isclose(s0[0], ref_u, rtol=1e-12)                    # independent of the number of ranks
    assert np.isclose(s0[4], np.arange(B * 3, dtype=np.float64).sum())
    assert s0[7] == B


def test_loss_from_sums_matches_oracle_loss():
    import smct_b200  # noqa: F401
    from smct_b200.distributed import loss_from_sums
    from oracle import smct_oracle as O
    from tests import util_cpu
    from tests.cases import CASES
    cfg = CASES["c3_b8"]["cfg"]
    p, x_in, y, eps, u = util_cpu.make_case(cfg, sigma2=0.5)
    out = O.forward(cfg, p, x_in, y, eps, u)
    sums = np.zeros(8)
    for i, (n, lv) in enumerate((("noise_K_resampled", "logvar_k"), ("noise_q", "logvar_q"), ("noise_V_resampled", "logvar_v"),
                                 ("noise_z", "logvar_z"))):
        sums[i] = (np.square(out[n].astype(np.float64)) / np.exp(float(p[lv]))).sum()
    sums[4] = np.square(y.reshape(cfg.B, 1, cfg.S, cfg.F_y).astype(np.float64) - out["pred_resampl"]).sum()
    n_rows = cfg.B * cfg.P * cfg.S
    tot, cl = loss_from_sums(sums, n_rows, cfg.sigma_obs, variant="pyc")
    ref_tot, ref_cl = O.smc_loss(cfg, p, out, y, "pyc")
    assert np.isclose(tot, ref_tot, rtol=1e-5) and np.isclose(cl, ref_cl, rtol=1e-5)
    lsum = sum(float(p["logvar_" + n]) * cfg.D for n in "kqvz")
    tot2, _ = loss_from_sums(sums, n_rows, cfg.sigma_obs, D=cfg.D, logvar_sum=lsum, variant="checkout")
    assert np.isclose(tot2, O.smc_loss(cfg, p, out, y, "checkout")[0], rtol=1e-5)
