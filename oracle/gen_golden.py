"""Writes tests/golden/oracle_<case>.npz from the numpy oracle (TEST INFRASTRUCTURE).

    python -m oracle.gen_golden

The uniforms are margin-screened (tests/util.screen_uniforms) so that the recorded ancestor indices
do not depend on last-bit differences of the log-weights.  Fixtures written by the reference's own
Python sources (over the TF shim) come from oracle/gen_golden_from_reference.py instead."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import smct_oracle as O  # noqa: E402

GOLDEN_CASES = ("c1", "pyc_smoke", "classification", "heads3")


def main():
    from tests.cases import CASES
    from tests import util_cpu
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    for name in GOLDEN_CASES:
        c = CASES[name]
        cfg = c["cfg"]
        p, x_in, y, eps, u = util_cpu.make_case(cfg, seed=2, sigma2=c["sigma2"], multi=c.get("multi", False))
        u, out = util_cpu.screen_uniforms(cfg, p, x_in, y, eps, u, margin=1e-5)
        total, classic = O.smc_loss(cfg, p, out, y, "checkout")
        total_pyc, _ = O.smc_loss(cfg, p, out, y, "pyc")
        arrays = dict(case=np.array(name), x_in=x_in, y=y, eps=eps, u=u, i_t=out["i_t"], pred=out["pred"],
                      pred_resampl=out["pred_resampl"], w=out["w"], K_last=out["K"][:, :, -1], R_last=out["R"][:, :, -1],
                      loss_checkout=np.array(total), loss_classic=np.array(classic), loss_pyc=np.array(total_pyc))
        arrays.update({"p_" + k: v for k, v in p.items()})
        path = os.path.join(outdir, "oracle_%s.npz" % name)
        np.savez_compressed(path, **arrays)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
