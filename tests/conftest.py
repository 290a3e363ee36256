import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Per-tensor worst parity error of the session (tests/util.py:assert_close): fraction of the tolerance used and
    |err| / max|ref|."""
    try:
        from tests import util
    except Exception:
        return
    if not getattr(util, "PARITY_LOG", None):
        return
    tr = terminalreporter
    tr.write_line("")
    tr.write_line("parity summary (tolerance |a-b| <= %g*|b| + %g*max|b|): tensor, comparisons, worst err/tol, worst err/max|ref|"
                  % (util.RTOL, util.RTOL * util.ATOL_FRAC))
    for k, (r_tol, r_scale, n) in sorted(util.PARITY_LOG.items()):
        tr.write_line("  [parity] %-16s n=%-5d err/tol=%.3f  err/scale=%.2e" % (k, n, r_tol, r_scale))
