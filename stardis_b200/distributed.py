"""nu sharding across the GPUs of one box (one process per GPU, torch.distributed).

The path is embarrassingly parallel in frequency: rank r evaluates pixels [p0, p1) of the GLOBAL grid
with global window centres / half-widths / d_nu, so there is no exchange inside the kernels.  The only collective is
the final all-gather of the emergent spectrum (or of F_nu blocks): NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n, rank, world_size):
    """Contiguous, balanced pixel ranges: rank r owns [p0, p1)."""
    base, rem = divmod(int(n), int(world_size))
    p0 = rank * base + min(rank, rem)
    return p0, p0 + base + (1 if rank < rem else 0)


def all_shards(n, world_size):
    return [shard_bounds(n, r, world_size) for r in range(world_size)]


def dist_info():
    try:
        import torch.distributed as dist
    except ImportError:
        return None, 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def allgather_spectrum(local, shard, n_total, device=None):
    """Every rank contributes its (W_r,) slice of a length-``n_total`` vector; returns the full vector on every rank.
    ``local`` may be a numpy array or a torch tensor (CUDA tensors are gathered with NCCL without touching the host)."""
    import torch

    dist, rank, world = dist_info()
    if dist is None or world == 1:
        out = np.asarray(local.cpu() if hasattr(local, "cpu") else local, dtype=np.float64)
        if out.shape[0] != n_total:
            raise ValueError("a sharded result needs an initialised process group to be gathered")
        return out
    bounds = all_shards(n_total, world)
    if tuple(bounds[rank]) != tuple(int(x) for x in shard):
        raise ValueError(f"rank {rank}: shard {shard} does not match the balanced partition {bounds[rank]}")
    is_tensor = isinstance(local, torch.Tensor)
    t = local if is_tensor else torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    # equal-sized contributions (gloo and NCCL both take the fast path): pad to the widest shard, trim afterwards
    wmax = max(b - a for a, b in bounds)
    padded = torch.zeros(wmax, dtype=torch.float64, device=t.device)
    padded[: t.shape[0]] = t
    pieces = [torch.empty(wmax, dtype=torch.float64, device=t.device) for _ in bounds]
    dist.all_gather(pieces, padded)
    full = torch.cat([p[: b - a] for p, (a, b) in zip(pieces, bounds)])
    return full if is_tensor else full.cpu().numpy()


def allgather_columns(local, shard, n_total):
    """(D, W_r) column blocks -> (D, n_total) on every rank (used for F_nu / total_alphas when the caller wants the
    whole radiation field)."""
    import torch

    dist, rank, world = dist_info()
    if dist is None or world == 1:
        return local
    is_tensor = isinstance(local, torch.Tensor)
    t = local if is_tensor else torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    D = t.shape[0]
    bounds = all_shards(n_total, world)
    wmax = max(b - a for a, b in bounds)
    padded = torch.zeros((D, wmax), dtype=torch.float64, device=t.device)
    padded[:, : t.shape[1]] = t
    pieces = [torch.empty((D, wmax), dtype=torch.float64, device=t.device) for _ in bounds]
    dist.all_gather(pieces, padded)
    full = torch.cat([p[:, : b - a] for p, (a, b) in zip(pieces, bounds)], dim=1)
    return full if is_tensor else full.cpu().numpy()
