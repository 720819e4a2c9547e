"""Sequence sharding across the GPUs of one box.  The reference splits the batch with
`torch.nn.DataParallel` (main_spectrogram.py:118-119): contiguous N/G blocks, parameters
replicated, no communication inside the layer.  Here: one process per GPU, each rank runs the
fused kernel on its contiguous block; a collective is used ONLY to gather outputs for verification."""
import torch


def shard_bounds(n, world_size, rank):
    """Contiguous block [lo, hi) of rank `rank` when n sequences are split like torch.chunk /
    DataParallel.scatter does (ceil(n / world) per rank, trailing ranks may be short or empty)."""
    per = -(-n // world_size)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def sharded_forward(layer_fn, x, group=None, gather=True):
    """Run `layer_fn` on this rank's shard of x (dim 0) and, if `gather`, all-gather the outputs so
    every rank holds the full (N, n_fft, F) result.  `layer_fn` maps (n_i,3,T,V,M) -> (n_i,n_fft,F)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(x.shape[0], world, rank)
    local = layer_fn(x[lo:hi])
    if not gather or world == 1:
        return local
    per = -(-x.shape[0] // world)
    padded = local.new_zeros((per,) + tuple(local.shape[1:]))
    padded[: hi - lo] = local
    full = local.new_empty((per * world,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(full, padded, group=group)
    return full[: x.shape[0]]
