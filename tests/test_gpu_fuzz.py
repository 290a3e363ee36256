"""Shape fuzzing of the forward: random small shapes — particle counts that are not powers of two,
several heads, attention windows, a resampling horizon, attention-only models, multivariate inputs/targets — through
every algorithm that serves the shape, against the CPU oracle on the same injected draws (screened uniforms)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import smct_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


@st.composite
def small_shapes(draw):
    H = draw(st.sampled_from([1, 1, 2, 4]))
    D = H * draw(st.sampled_from([2, 4, 6, 8]))
    S = draw(st.integers(1, 9))
    full = draw(st.booleans())
    return dict(B=draw(st.integers(1, 3)), P=draw(st.integers(1, 37)), S=S, D=D, H=H,
                dff=draw(st.sampled_from([4, 8, 12])) if full else 0, full_model=full,
                F_x=draw(st.integers(1, 3)), F_y=draw(st.integers(1, 3)),
                attn_window=draw(st.one_of(st.none(), st.integers(1, 4))),
                len_resampling=draw(st.one_of(st.none(), st.integers(0, S))),
                noise_z_mode=draw(st.sampled_from(["checkout", "pyc"])))


@settings(max_examples=80, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(shape=small_shapes(), seed=st.integers(0, 2 ** 16), sigma2=st.sampled_from([0.05, 0.5, 1.5]))
def test_forward_fuzz_small_shapes(shape, seed, sigma2):
    cfg = O.Config(**shape)
    p, x_in, y, eps, u = util.make_case(cfg, seed=seed, sigma2=sigma2)
    u, ora = util.screen_uniforms(cfg, p, x_in, y, eps, u)
    for algo in ("staged", "staged_seq"):
        dev = util.run_device(cfg, p, x_in, y, eps, u, algo=algo)
        util.compare_forward(cfg, p, dev, ora, y)


@settings(max_examples=30, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(B=st.integers(1, 2), P=st.integers(1, 150), S=st.integers(1, 12), dff=st.sampled_from([64, 128, 192, 256]),
       F=st.integers(1, 3), window=st.one_of(st.none(), st.integers(1, 5)), seed=st.integers(0, 2 ** 16))
def test_fused_forward_fuzz(B, P, S, dff, F, window, seed):
    cfg = O.Config(B=B, P=P, S=S, D=64, dff=dff, F_x=F, F_y=F, attn_window=window)
    p, x_in, y, eps, u = util.make_case(cfg, seed=seed, sigma2=0.5)
    u, ora = util.screen_uniforms(cfg, p, x_in, y, eps, u)
    want = ("i_out", "logw", "noise_q", "noise_z", "noise_K", "noise_V", "loss_sums")
    for algo in ("fused", "fused_ffma"):
        dev = util.run_device(cfg, p, x_in, y, eps, u, want=want, algo=algo)
        util.compare_forward(cfg, p, dev, ora, y)


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(B=st.integers(1, 3), P=st.integers(1, 24), S=st.integers(1, 8), D=st.sampled_from([4, 8, 16, 40]),
       dff=st.sampled_from([4, 8, 20]), F=st.integers(1, 3), window=st.one_of(st.none(), st.integers(1, 4)),
       stop=st.one_of(st.none(), st.integers(0, 8)), variant=st.sampled_from(["pyc", "checkout"]), seed=st.integers(0, 999))
def test_backward_fuzz(B, P, S, D, dff, F, window, stop, variant, seed):
    """smct_backward (node de-duplication, multiplicity weights) against fp64 torch autograd on random small shapes."""
    from tests.test_gpu_backward import _compare, _run
    cfg = O.Config(B=B, P=P, S=S, D=D, dff=dff, F_x=F, F_y=F, attn_window=window,
                   len_resampling=None if stop is None else min(stop, S), noise_z_mode="pyc")
    got, ref, _ = _run(cfg, variant, seed=seed)
    _compare("fuzz", got, ref)
