"""Multi-GPU partitioning of the gridding path: independent units, NO data-path collective.

  time steps : unit = one (day, hour, variable) raster; round-robin over ranks, rasters replicated once per GPU
  row bands  : unit = a band of DEM rows; every raster entry point of the C ABI takes (rowBegin, rowEnd) and writes only
               its band of the output grid; bands are balanced on the number of VALID cells per row

The only inter-rank traffic is the host-side gather of the finished bands (and, in bench.py, a barrier plus the
max-over-ranks reduction of the timings). torch.distributed is plumbing here: NCCL on the GPU box, gloo in the CPU tests.
"""
import numpy as np


def time_steps_for_rank(n_steps, rank, world):
    """indices of the time steps rank `rank` of `world` grids (round-robin keeps consecutive hours on different GPUs)."""
    return list(range(rank, n_steps, world))


def row_bands(valid_per_row, world):
    """contiguous row bands [(r0, r1)] * world with (nearly) equal numbers of valid cells; valid_per_row: (rows,) counts."""
    valid_per_row = np.asarray(valid_per_row, dtype=np.int64)
    rows = valid_per_row.shape[0]
    cum = np.concatenate([[0], np.cumsum(valid_per_row)])
    total = int(cum[-1])
    bounds = [0]
    for k in range(1, world):
        target = total * k / world
        r = int(np.searchsorted(cum, target, side="left"))
        bounds.append(min(max(r, bounds[-1]), rows))
    bounds.append(rows)
    return [(bounds[k], bounds[k + 1]) for k in range(world)]


def valid_cells_per_row(dem_data, flag):
    """cells with !isEqual(z, flag) per row (interpolationCmd.cpp:377)."""
    return np.count_nonzero(np.abs(dem_data.astype(np.float64) - float(flag)) >= 1e-5, axis=1)


def gather_bands(local_grid, band, dist, rank, world, flag):
    """host gather of the per-rank output bands into the full grid on rank 0 (no reduction: bands are disjoint).
    `dist` is torch.distributed (any backend able to move CPU tensors, or NCCL with device tensors moved by the caller)."""
    import torch
    r0, r1 = band
    mine = torch.from_numpy(np.ascontiguousarray(local_grid[r0:r1]))
    meta = [None] * world
    dist.all_gather_object(meta, (r0, r1))
    if rank == 0:
        full = np.full(local_grid.shape, np.float32(flag), dtype=np.float32)
        full[r0:r1] = local_grid[r0:r1]
        for src in range(1, world):
            a, b = meta[src]
            buf = torch.empty((b - a, local_grid.shape[1]), dtype=torch.float32)
            if b > a:
                dist.recv(buf, src=src)
            full[a:b] = buf.numpy()
        return full
    if r1 > r0:
        dist.send(mine, dst=0)
    return None


def max_over_ranks(value, dist, world, device=None):
    """timing rule of the bench contract: the job takes as long as its slowest rank."""
    import torch
    if world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
