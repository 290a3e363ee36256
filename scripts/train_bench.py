"""C5-style measurement: one SMC-T TRAINING step (forward + backward + gradient all-reduce + Adam) on a C4-shaped
batch shard per GPU.  Run directly (1 GPU) or under torchrun (N GPUs, NCCL).  Prints one JSON line on rank 0.
Not the bench.py contract (that one is the forward metric); this documents SURVEY.md §8 config 5."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smct_b200, smct_b200.functional as F
from smct_b200 import utils
from smct_b200.train.optimizers import Adam, CustomSchedule

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=148); ap.add_argument("--P", type=int, default=1024)
ap.add_argument("--S", type=int, default=128); ap.add_argument("--steps", type=int, default=2); ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--chunks", type=int, default=1, help="resident chunks of --batch sequences per optimiser step (gradients accumulated; 28 chunks of 148 = the C5 batch of 4096+ sequences per GPU)")
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
B, P, S, D, dff, Fdim = a.batch, a.P, a.S, 64, 256, 1
params = {k: torch.as_tensor(v, device=dev) for k, v in utils.init_parameters(D, dff, Fdim, Fdim, sigma2=0.5).items()}
NC = a.chunks
x, y = utils.synthetic_ar1(B * NC, S, Fdim, 0.9, 0.3, seed=1234 + rank)
tx_all, ty_all = torch.as_tensor(x, device=dev), torch.as_tensor(y, device=dev)
shapes = [F.make_shape(B, P, S, D, dff=dff, F_x=Fdim, F_y=Fdim, noise_z_mode="pyc", algo="auto", b_global0=(rank * NC + c) * B)
          for c in range(NC)]
shape, tx, ty = shapes[0], tx_all[:B], ty_all[:B]
opt = Adam(CustomSchedule(D), beta_1=0.9, beta_2=0.98, epsilon=1e-9)
names = sorted(n for n in params)
inv_n = 1.0 / (float(B) * NC * world * P * S)
out = None

def step(i):
    global out
    grads, sums = None, torch.zeros(8, dtype=torch.float64, device=dev)
    for c in range(NC):   # chunks of the batch: forward + backward per chunk, gradients accumulated with the global 1/n
        xc, yc = tx_all[c * B:(c + 1) * B], ty_all[c * B:(c + 1) * B]
        out = F.forward(shapes[c], params, xc, yc, seed=100 + i, want=("i_out", "loss_sums"), out=out)
        grads = F.backward(shapes[c], params, xc, yc, out, seed=100 + i, inv_n=inv_n, grads=grads)
        sums += out["loss_sums"]
    flat = torch.cat([grads[n].reshape(-1) for n in names] + [sums.float()])
    if dist is not None:
        dist.all_reduce(flat)
    o = 0
    for n in names:
        k = grads[n].numel(); grads[n].copy_(flat[o:o + k].reshape(grads[n].shape)); o += k
    opt.apply_gradients((grads[n], params[n]) for n in names)
    return flat[o:o + 8]

for i in range(a.warmup):
    step(i)
if dist is not None:
    dist.barrier()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
t_f = t_b = 0.0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    sums = step(a.warmup + i)
e1.record()
if dist is not None:
    dist.barrier()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if dist is not None:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# forward-only time for the split
out2 = F.forward(shape, params, tx, ty, seed=1, want=("i_out", "loss_sums"), out=out)
torch.cuda.synchronize(); e0.record(); F.forward(shape, params, tx, ty, seed=2, want=("i_out", "loss_sums"), out=out); e1.record(); torch.cuda.synchronize()
fwd_ms = e0.elapsed_time(e1)
if rank == 0:
    n = float(B) * NC * world * P * S
    tot = sums.double().cpu().numpy()
    loss = (0.5 * tot[:4].sum() + tot[4]) / n   # sigma_obs = 0.5
    print(json.dumps({"what": "SMC-T training step (fwd + bwd + NCCL grad all-reduce + Adam)", "n_gpus": world,
                      "per_gpu_batch": B * NC, "chunks_per_step": NC, "P": P, "S": S, "d_model": D, "dff": dff, "ms_per_step": float(t.item()),
                      "forward_ms": fwd_ms, "particle_steps_per_s": n / (float(t.item()) * 1e-3), "loss": float(loss),
                      "grad_elems_allreduced": int(sum(params[k].numel() for k in names))}))
if dist is not None:
    dist.destroy_process_group()
