"""Data parallelism over collocation points: every rank owns a contiguous slice of the batch, weights
and optimizer state are replicated, the per-rank partial gradients -- computed with ``loss_scale = 2 / (N_global n_eq)``
so that they simply add up -- and the partial sums of squared residuals travel in ONE flat buffer ``[grad | sum r^2]``
that is all-reduced once per optimizer step (NCCL over NVLink on the GPUs; gloo in the CPU tests)."""
import torch


def shard_bounds(n_points, rank, world_size):
    """Half-open slice of rank ``rank`` of a batch of ``n_points`` (slices tile the batch, sizes differ by <= 1)."""
    return (n_points * rank) // world_size, (n_points * (rank + 1)) // world_size


def shard_coords(coords, rank, world_size):
    lo, hi = shard_bounds(coords[0].shape[0], rank, world_size)
    return [c[lo:hi].contiguous() for c in coords]


def all_reduce_gradbuf(gradbuf, dist=None):
    """Sum ``[grad_theta | sum r^2]`` over the ranks, in place.  No-op without an initialised process group."""
    dist = dist or torch.distributed
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(gradbuf)
    return gradbuf
