"""Quick device timing of smct_forward on a C4-shaped shard (used while developing; bench.py is the contract)."""
import argparse, json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smct_b200, smct_b200.functional as F
from oracle import smct_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32); ap.add_argument("--P", type=int, default=1024)
ap.add_argument("--S", type=int, default=128); ap.add_argument("--D", type=int, default=64)
ap.add_argument("--dff", type=int, default=256); ap.add_argument("--algo", default="staged")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
cfg = O.Config(B=a.B, P=a.P, S=a.S, D=a.D, dff=a.dff)
p = O.init_params(cfg)
x_in, y = O.synthetic_series(cfg.B, cfg.S, 1, alpha=0.9, var=0.3)
shape = F.shape_from_config(cfg, algo=a.algo)
from tests import util
dp = util.dev_params(p)
tx, ty = torch.as_tensor(x_in, device="cuda"), torch.as_tensor(y, device="cuda")
out = F.forward(shape, dp, tx, ty, seed=1, want=("loss_sums",))
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
ev[0].record()
for i in range(a.iters):
    F.forward(shape, dp, tx, ty, seed=1, want=("loss_sums",), out=out)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters)]
ps = a.B * a.P * a.S
bytes_alg = ps * 4 * a.D * (4 * a.S + 5)
best = min(ms)
print(json.dumps(dict(algo=a.algo, B=a.B, P=a.P, S=a.S, D=a.D, ms=ms, psteps_per_s=ps / best * 1e3,
                      alg_GBs=bytes_alg / best / 1e6, loss_sums=out["loss_sums"].cpu().tolist()[:5])))
