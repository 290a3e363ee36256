"""tcgen05 building blocks (smct_tc.cuh): three-term scaled-fp16 split GEMM through TMEM vs an fp64 reference."""
import numpy as np
import pytest
import torch

import smct_b200.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [0, 1])
def test_tc_selftest_matches_fp64(seed):
    rng = np.random.default_rng(seed)
    A = (rng.standard_normal((128, 64)) * (1.0 + 3.0 * rng.random((128, 1)))).astype(np.float32)
    W = rng.uniform(-0.3, 0.3, size=(64, 64)).astype(np.float32)
    if seed == 1:   # structured: identity weights expose any row/column permutation of the operand layout
        W = np.eye(64, dtype=np.float32)
        A = np.arange(128 * 64, dtype=np.float32).reshape(128, 64) / 64.0
    D = F.tc_selftest(torch.as_tensor(A, device="cuda"), torch.as_tensor(W, device="cuda")).cpu().numpy()
    ref = A.astype(np.float64) @ W.astype(np.float64)
    err = np.abs(D - ref).max()
    scale = np.abs(ref).max()
    assert err <= 1e-6 * scale, (err, scale)   # fp16 x2 split with exact power-of-two scaling: ~22 mantissa bits
