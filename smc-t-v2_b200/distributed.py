"""Batch sharding of the SMC forward across the GPUs of one box.

Sequences never interact in the forward (attention, weights, the categorical draw and the ancestor
gather are all per sequence), so the batch is split into contiguous shards, one process per GPU, with NO
data-path collective.  All particles of a sequence stay on one GPU.  The device RNG is keyed by the
GLOBAL sequence index (smct_shape.b_global0), so results do not depend on the number of ranks.
The only exchange is the reduction of the scalar loss sums (and, once the backward exists, one flat
gradient buffer): torch.distributed all_reduce over NCCL (NVLink) — gloo in the CPU tests."""
import torch


def shard_range(B, rank, world):
    """Contiguous shard [b0, b1) of a batch of B sequences for `rank` of `world` (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside [0, %d)" % (rank, world))
    base, rem = divmod(B, world)
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_sums(t: torch.Tensor) -> torch.Tensor:
    """Sum a (small) tensor of loss sums / counts over all ranks, in place; no-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def loss_from_sums(sums, n_rows, sigma_obs=0.5, D=None, logvar_sum=None, variant="pyc"):
    """Scalar SMC loss from the (all-reduced) sums of smct_io.loss_sums over `n_rows` = B*P*S rows.
    pyc: mean 0.5*sum_n ||n||^2/sigma_n^2 + mean 0.5*||y-pred_resampl||^2/Sigma_obs;
    checkout adds the log-determinant terms 0.5*D*log(2pi)*4 + 0.5*sum_n sum_d logvar_n[d]."""
    import math
    smc = 0.5 * float(sums[0] + sums[1] + sums[2] + sums[3]) / n_rows
    if variant == "checkout":
        smc += 4 * 0.5 * D * math.log(2 * math.pi) + 0.5 * float(logvar_sum)
    classic = (0.5 / sigma_obs) * float(sums[4]) / n_rows
    return smc + classic, classic


def sharded_forward(forward_fn, x, y, B, seed):
    """Run `forward_fn(x_shard, y_shard, b_global0, seed) -> loss_sums (8,) fp64 tensor` on this rank's shard
    and return the all-reduced sums.  `x`, `y` hold the FULL batch on every rank (host tensors)."""
    rank, ws = world()
    b0, b1 = shard_range(B, rank, ws)
    sums = forward_fn(x[b0:b1], y[b0:b1], b0, seed)
    return allreduce_sums(sums.clone())
