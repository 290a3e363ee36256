import sys, os, json
sys.path.insert(0, os.getcwd())
import torch, smct_b200, smct_b200.functional as F
from smct_b200 import utils
import bench
for wl in ("c1", "c2", "c3"):
    B, P, S, D, dff, Fdim, alpha, var = bench.WORKLOADS[wl]
    params = {k: torch.as_tensor(v, device="cuda") for k, v in utils.init_parameters(D, dff, Fdim, Fdim).items()}
    x, y = utils.synthetic_ar1(B, S, Fdim, alpha, var)
    xd, yd = torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda")
    shape = F.make_shape(B, P, S, D, dff=dff, F_x=Fdim, F_y=Fdim, algo="auto")
    out = F.forward(shape, params, xd, yd, seed=1, want=("loss_sums",))
    for i in range(3):
        F.forward(shape, params, xd, yd, seed=2 + i, want=("loss_sums",), out=out)
    torch.cuda.synchronize()
