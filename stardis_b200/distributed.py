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


def line_balanced_bounds(tracing_nus, line_nus, world_size, line_weight=10.0, align=512):
    """Contiguous pixel ranges of (nearly) equal COST instead of equal width.

    With the far-field evaluation of the line wings the work of a pixel range is no longer proportional to its width:
    what remains per range is a per-pixel part (polynomial evaluation, continuum, formal solution) plus the directly
    evaluated line cores, which follow the number of lines whose centre lies in the range -- and a grid that is uniform
    in wavelength holds an order of magnitude more lines per pixel at its blue end.  cost(range) = pixels +
    ``line_weight`` * lines inside (one line core costs about ten pixels on a B200; measured with
    tools/shard_probe.py).  Cuts are multiples of ``align`` pixels (the level-0 tile of the line kernel) so that no
    tile is evaluated by two ranks.  Every rank computes the same list from the same inputs; pass it to
    ``allgather_spectrum(..., bounds=...)``."""
    nus = np.asarray(tracing_nus, dtype=np.float64)
    n, world_size = int(nus.shape[0]), int(world_size)
    if world_size <= 1:
        return [(0, n)]
    line_nus = np.asarray(line_nus, dtype=np.float64)
    if n >= 2 and nus[0] > nus[-1]:  # descending grid: position from the other end
        pos = n - np.searchsorted(nus[::-1], line_nus, side="left")
    else:
        pos = np.searchsorted(nus, line_nus, side="left")
    pos = pos[(pos > 0) & (pos < n)] if line_nus.size else np.zeros(0, dtype=np.int64)
    align = max(1, min(int(align), n // (8 * world_size)))  # short grids: finer cuts rather than empty ranges
    cuts = np.arange(0, n, align, dtype=np.int64)
    cuts = np.append(cuts, n)  # candidate cut positions (aligned) and the end of the grid
    lines_below = np.searchsorted(np.sort(pos), cuts, side="left")
    cost = cuts.astype(np.float64) + float(line_weight) * lines_below
    if n < world_size:
        raise ValueError(f"cannot split {n} pixels into {world_size} non-empty ranges")
    step = align if n >= world_size * align else 1  # smallest admissible range (one aligned block; pixels on tiny grids)
    bounds, prev = [], 0
    for r in range(1, world_size):
        k = int(np.searchsorted(cost, cost[-1] * r / world_size, side="left"))
        cut = int(cuts[min(max(k, 0), len(cuts) - 1)])
        # every rank gets at least one block (sd_set_grid needs p0 < p1; an empty range would leave its rank out of the
        # collectives): not before prev + step, and room for the ranks still to come
        cut = min(max(cut, prev + step), n - (world_size - r) * step)
        bounds.append((prev, cut))
        prev = cut
    bounds.append((prev, n))
    return bounds


def dist_info():
    try:
        import torch.distributed as dist
    except ImportError:
        return None, 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def allgather_spectrum(local, shard, n_total, device=None, bounds=None):
    """Every rank contributes its (W_r,) slice of a length-``n_total`` vector; returns the full vector on every rank.
    ``local`` may be a numpy array or a torch tensor (CUDA tensors are gathered with NCCL without touching the host).
    ``bounds``: the partition [(p0, p1)] * world all ranks agreed on (default: equal widths, ``all_shards``)."""
    import torch

    dist, rank, world = dist_info()
    if dist is None or world == 1:
        out = np.asarray(local.cpu() if hasattr(local, "cpu") else local, dtype=np.float64)
        if out.shape[0] != n_total:
            raise ValueError("a sharded result needs an initialised process group to be gathered")
        return out
    bounds = _checked_bounds(bounds, n_total, world)
    if tuple(bounds[rank]) != tuple(int(x) for x in shard):
        raise ValueError(f"rank {rank}: shard {shard} does not match the partition {bounds[rank]}")
    is_tensor = isinstance(local, torch.Tensor)
    t = local if is_tensor else torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    if len({b - a for a, b in bounds}) == 1 and t.is_cuda:  # equal widths: one collective straight into the result
        full = torch.empty(n_total, dtype=torch.float64, device=t.device)
        dist.all_gather_into_tensor(full, t.contiguous())
        return full if is_tensor else full.cpu().numpy()
    # equal-sized contributions (gloo and NCCL both take the fast path): pad to the widest shard, trim afterwards
    wmax = max(b - a for a, b in bounds)
    padded = torch.zeros(wmax, dtype=torch.float64, device=t.device)
    padded[: t.shape[0]] = t
    pieces = [torch.empty(wmax, dtype=torch.float64, device=t.device) for _ in bounds]
    dist.all_gather(pieces, padded)
    full = torch.cat([p[: b - a] for p, (a, b) in zip(pieces, bounds)])
    return full if is_tensor else full.cpu().numpy()


def _checked_bounds(bounds, n_total, world):
    if bounds is None:
        return all_shards(n_total, world)
    bounds = [(int(a), int(b)) for a, b in bounds]
    ok = len(bounds) == world and bounds[0][0] == 0 and bounds[-1][1] == n_total
    ok = ok and all(a <= b for a, b in bounds) and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
    if not ok:
        raise ValueError(f"{bounds} is not a contiguous partition of [0, {n_total}) into {world} ranges")
    return bounds


def allgather_columns(local, shard, n_total, bounds=None):
    """(D, W_r) column blocks -> (D, n_total) on every rank (used for F_nu / total_alphas when the caller wants the
    whole radiation field)."""
    import torch

    dist, rank, world = dist_info()
    if dist is None or world == 1:
        return local
    is_tensor = isinstance(local, torch.Tensor)
    t = local if is_tensor else torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    D = t.shape[0]
    bounds = _checked_bounds(bounds, n_total, world)
    wmax = max(b - a for a, b in bounds)
    padded = torch.zeros((D, wmax), dtype=torch.float64, device=t.device)
    padded[:, : t.shape[1]] = t
    pieces = [torch.empty((D, wmax), dtype=torch.float64, device=t.device) for _ in bounds]
    dist.all_gather(pieces, padded)
    full = torch.cat([p[:, : b - a] for p, (a, b) in zip(pieces, bounds)], dim=1)
    return full if is_tensor else full.cpu().numpy()


def stripe_rows(n_rows, rank, world_size):
    """Row block [r0, r1) of an (n_rows, D) host array that rank `rank` uploads in ``upload_rows_striped`` (equal blocks
    of ceil(n_rows / world) rows; the last ones may be short or empty)."""
    rows = -(-int(n_rows) // int(world_size))
    r0 = min(rank * rows, n_rows)
    return r0, min(r0 + rows, n_rows), rows


def upload_rows_striped(host_array, device):
    """(n, D) float64 HOST array, identical on every rank -> the same array in the HBM of every rank, moved over PCIe only
    once in total: every rank uploads 1/world of the rows and the blocks are exchanged with one NCCL all-gather over
    NVLink (gloo in the CPU test).  The (L, D) line-strength table is the largest host->device transfer of a run
    (134 MB at the flagship size); in a sharded run every rank needs all of it, and 8 ranks pulling it through the host
    at the same time is what keeps the end-to-end rate below the device rate.  Returns a torch tensor (n, D) on
    ``device`` (or the input when there is no process group).  Collective: every rank must call it with the same data."""
    import torch

    dist, rank, world = dist_info()
    if dist is None or world == 1:
        return host_array
    a = np.ascontiguousarray(host_array, dtype=np.float64)
    n, d = a.shape
    r0, r1, rows = stripe_rows(n, rank, world)
    local = torch.zeros((rows, d), dtype=torch.float64, device=device)
    if r1 > r0:
        local[: r1 - r0].copy_(torch.from_numpy(a[r0:r1]), non_blocking=True)
    full = torch.empty((rows * world, d), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(full, local)
    return full[:n]


# ------------------------------------------------------------------------------------------------- depth sharding
# Under nu sharding every rank repeats the per-(line, depth) preparation of the line kernel for the whole line list
# (windows, far-field levels, class lists, edge sort: independent of the rank's pixel range for the ~40 % of the pairs
# whose window spans the whole grid), which caps the 8-GPU efficiency near 0.5.  All opacity stages (K1, preparation, K2,
# K3) are independent per DEPTH POINT, so the multi-GPU driver shards those by depth instead -- rank r evaluates every
# R-th depth point (dealt in serpentine order, ``depth_indices``: neighbouring depths cost about the same, so the ranks are
# balanced) on the WHOLE grid -- and the formal solution, which couples all depths of one frequency, by nu.  Between the two sits the only
# real exchange step of the path: one all-to-all of the total opacity (every rank sends (D/R, N/R) blocks, 34 MB per
# rank at the flagship size), NCCL over NVLink.  Every depth row is computed exactly as in a single-GPU run, so the
# result is bitwise identical for every R.

def depth_indices(n_depth, rank, world_size):
    """Depth points of rank ``rank``, ascending.  Depth points are dealt to the ranks in serpentine order (0..R-1, then
    R-1..0, ...): the cost of a depth point falls smoothly from the hot, deep layers to the surface, and a plain
    round-robin would give rank 0 the most expensive member of every group of R (measured on 8 GPUs: 11 % above the
    mean)."""
    n_depth, rank, world_size = int(n_depth), int(rank), int(world_size)
    d = np.arange(n_depth)
    pos, grp = d % world_size, d // world_size
    owner = np.where(grp % 2 == 0, pos, world_size - 1 - pos)
    return d[owner == rank]


_index_cache = {}


def _depth_rows(n_depth, rank, world, device):
    """``depth_indices`` as a cached device index tensor (no host->device copy per call)."""
    import torch

    key = ("rows", int(n_depth), int(rank), int(world), str(device))
    if key not in _index_cache:
        _index_cache[key] = torch.as_tensor(depth_indices(n_depth, rank, world), dtype=torch.long, device=device)
    return _index_cache[key]


def _depth_gather_perm(n_depth, world, device):
    """perm[d] = row of depth d in the (world * D/world) rank-major stack of the ranks' depth rows (D divisible by world)."""
    import torch

    key = ("perm", int(n_depth), int(world), str(device))
    if key not in _index_cache:
        per = int(n_depth) // int(world)
        perm = np.empty(int(n_depth), dtype=np.int64)
        for s in range(world):
            perm[depth_indices(n_depth, s, world)] = s * per + np.arange(per)
        _index_cache[key] = torch.as_tensor(perm, dtype=torch.long, device=device)
    return _index_cache[key]


def exchange_depth_to_nu(local, n_depth, n_total, bounds=None):
    """All-to-all between the two decompositions: ``local`` (D_r, N) holds this rank's depth rows (``depth_indices``) over
    the whole grid; returns (D, W_r), all depth rows over this rank's pixel range ``bounds[rank]`` (default: equal
    widths).  torch tensors (CUDA -> NCCL ``all_to_all_single``; CPU/gloo -> point-to-point sends, used by the tests);
    returns a tensor on the same device.  Without a process group the input is returned unchanged.  With D divisible by
    the world size and equal-width ranges (the usual case) the whole exchange is three device operations: one transposing
    copy, the all-to-all, one row gather."""
    import torch

    dist, rank, world = dist_info()
    if dist is None or world == 1:
        return local
    bounds = _checked_bounds(bounds, n_total, world)
    n_depth = int(n_depth)
    widths = {b - a for a, b in bounds}
    if local.is_cuda and n_depth % world == 0 and len(widths) == 1:
        w = widths.pop()
        send = local.view(local.shape[0], world, w).permute(1, 0, 2).contiguous()   # (R, D_r, W): block j -> rank j
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        return recv.view(n_depth, w).index_select(0, _depth_gather_perm(n_depth, world, local.device))
    d_pad = -(-n_depth // world)
    w_pad = max(b - a for a, b in bounds)
    send = torch.zeros((world, d_pad, w_pad), dtype=torch.float64, device=local.device)
    for j, (a, b) in enumerate(bounds):
        send[j, : local.shape[0], : b - a] = local[:, a:b]
    recv = torch.empty_like(send)
    if local.is_cuda:
        dist.all_to_all_single(recv, send)
    else:  # gloo has no all-to-all: pairwise exchange
        reqs = []
        for j in range(world):
            if j == rank:
                recv[j] = send[j]
            else:
                reqs.append(dist.isend(send[j].contiguous(), j))
                reqs.append(dist.irecv(recv[j], j))
        for r in reqs:
            r.wait()
    a, b = bounds[rank]
    out = torch.empty((n_depth, b - a), dtype=torch.float64, device=local.device)
    for s in range(world):
        rows = _depth_rows(n_depth, s, world, local.device)
        out.index_copy_(0, rows, recv[s, : rows.shape[0], : b - a])
    return out


def allgather_depth_columns(local, n_depth):
    """(L, D_r) columns of this rank's depth points -> (L, D) on every rank (gammas / Doppler widths when the caller asked
    for the whole radiation field)."""
    import torch

    dist, rank, world = dist_info()
    if dist is None or world == 1:
        return local
    d_pad = -(-int(n_depth) // world)
    padded = torch.zeros((local.shape[0], d_pad), dtype=torch.float64, device=local.device)
    padded[:, : local.shape[1]] = local
    pieces = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(pieces, padded)
    out = torch.empty((local.shape[0], int(n_depth)), dtype=torch.float64, device=local.device)
    for s in range(world):
        rows = _depth_rows(n_depth, s, world, local.device)
        out.index_copy_(1, rows, pieces[s][:, : rows.shape[0]])
    return out


class DepthSlicedModel:
    """View of a stellar model restricted to some depth points (what the opacity stages read: temperatures, composition,
    microturbulence; the geometry belongs to the formal solution and is not sliced)."""

    def __init__(self, stellar_model, idx):
        self._m = stellar_model
        self.temperatures = stellar_model.temperatures[idx]
        self.no_of_depth_points = len(idx)

    def __getattr__(self, name):
        return getattr(self._m, name)


class DepthSlicedPlasma:
    """View of a plasma restricted to some depth points: every per-depth table the opacity stages read is
    sliced, everything else passes through.  Cached per plasma and depth selection (the columnar line table of the view
    is built once)."""

    _SERIES = ("electron_densities", "h_minus_density", "h2_density", "h2_plus_density")
    _FRAMES = ("ion_number_density", "level_number_density", "partition_function")
    _WITH_NU = ("alpha_line", "alpha_line_from_linelist", "molecule_alpha_line_from_linelist")

    def __init__(self, stellar_plasma, idx):
        self._p = stellar_plasma
        self._idx = np.asarray(idx)
        self._cache = {}

    @classmethod
    def of(cls, stellar_plasma, idx):
        key = tuple(int(i) for i in idx)
        try:
            views = stellar_plasma.__dict__.setdefault("_stardis_b200_depth_views", {})
        except AttributeError:
            return cls(stellar_plasma, idx)
        if key not in views:
            views[key] = cls(stellar_plasma, idx)
        return views[key]

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        cache = self.__dict__["_cache"]
        if name in cache:
            return cache[name]
        v = getattr(self._p, name)
        idx = self._idx
        if v is None:
            out = None
        elif name in self._SERIES:
            out = v.iloc[idx]
        elif name in self._FRAMES:
            out = v.iloc[:, idx]
        elif name in self._WITH_NU:  # (L, D) + a trailing "nu" column
            cols = list(v.columns)
            out = v[[cols[i] for i in idx] + ["nu"]]
        elif name == "line_table":
            out = _slice_line_table(v, idx)
        else:
            out = v
        cache[name] = out
        return out


def _slice_line_table(table, idx):
    """ColumnarLines with the depth columns ``idx`` of alpha_line and of the line-strength producer tables."""
    from .plasma.columnar import ColumnarLines, LineStrength

    if table is None:
        return None
    strength = table.strength
    if strength is not None:
        strength = LineStrength(strength.kind, strength.per_line,
                                {k: (np.ascontiguousarray(v[:, idx]) if np.ndim(v) == 2 else v) for k, v in strength.tables.items()})
    cols = {k: getattr(table, k) for k in ColumnarLines._data_fields()}
    if cols["alpha_line"] is not None:
        cols["alpha_line"] = np.ascontiguousarray(cols["alpha_line"][:, idx])
    return ColumnarLines(**cols, strength=strength)


def upload_columns_striped(columns, device):
    """Per-line columns (equal-length 1-D float64 / int64 host arrays, identical on every rank) -> device tensors on every
    rank, moved over PCIe only once in total: every rank uploads 1/world of the lines of all columns in ONE copy and the
    blocks are exchanged with one all-gather over NVLink (gloo in the CPU test).  In a multi-GPU run every rank needs the
    whole line table (24 MB at the flagship size); R ranks pulling it through the host at the same time is what separates
    the end-to-end rate from the device rate.  Returns the input (a dict) unchanged without a process group."""
    import torch

    dist, rank, world = dist_info()
    if dist is None or world == 1:
        return columns
    names = list(columns)
    n = len(columns[names[0]])
    r0, r1, rows = stripe_rows(n, rank, world)
    block = np.zeros((len(names), rows), dtype=np.int64)  # 8-byte cells: float64 columns travel as their bit patterns
    for k, name in enumerate(names):
        a = np.ascontiguousarray(columns[name])
        if a.dtype.itemsize != 8 or a.shape != (n,):
            raise ValueError(f"column {name!r}: need a 1-D 8-byte column of length {n}")
        block[k, : r1 - r0] = a[r0:r1].view(np.int64)
    local = torch.from_numpy(block).to(device, non_blocking=True)
    gathered = torch.empty((world * len(names), rows), dtype=torch.int64, device=device)  # rank-major blocks along dim 0
    dist.all_gather_into_tensor(gathered, local)
    full = gathered.view(world, len(names), rows).permute(1, 0, 2).reshape(len(names), world * rows)[:, :n].contiguous()
    return {name: (full[k].view(torch.float64) if np.asarray(columns[name]).dtype.kind == "f" else full[k])
            for k, name in enumerate(names)}
