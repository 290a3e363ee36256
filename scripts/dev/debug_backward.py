import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests.test_gpu_backward import BWD_CASES, _run
for name in sys.argv[1:] or ["c1"]:
    got, ref, p = _run(BWD_CASES[name], "pyc")
    print("==", name)
    for k in sorted(ref):
        r = np.asarray(ref[k], np.float64).reshape(-1); g = np.asarray(got[k], np.float64).reshape(-1)
        sc = np.abs(r).max(); err = np.abs(g - r).max()
        cos = float((g * r).sum() / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-300))
        print("%-9s scale %.3e err %.3e rel %.2e cos %.6f ratio %.4f" % (k, sc, err, err / max(sc, 1e-300), cos, float(np.linalg.norm(g) / (np.linalg.norm(r) + 1e-300))))
