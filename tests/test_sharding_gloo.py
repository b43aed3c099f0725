"""CPU, world_size = 2 over gloo: the N > 1 host logic — row-band partition balanced on valid cells, per-rank gridding of
its band only, host gather on rank 0, max-over-ranks timing, round-robin time steps. The per-band computation here is
the CPU oracle (test infrastructure) standing in for the GPU engine, which takes the same (rowBegin, rowEnd)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from praga_b200 import _abi as A
from praga_b200 import sharding
from praga_b200 import synthetic as syn


def test_time_steps_round_robin_cover_everything_once():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in sharding.time_steps_for_rank(37, r, world))
        assert seen == list(range(37))


def test_row_bands_balance_valid_cells():
    dem = syn.make_dem(300, 200)
    v = sharding.valid_cells_per_row(dem.data, dem.flag)
    for world in (2, 4, 8):
        bands = sharding.row_bands(v, world)
        assert bands[0][0] == 0 and bands[-1][1] == 300
        assert all(bands[k][1] == bands[k + 1][0] for k in range(world - 1))
        per = [int(v[a:b].sum()) for a, b in bands]
        assert sum(per) == int(v.sum())
        assert max(per) - min(per) <= 2 * int(v.max())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding
    port_oracle = binding.Oracle("port")
    dem = syn.make_dem(96, 80, seed=4)
    st = syn.make_stations(dem, 60, variable="humidity", seed=5)
    s = A.default_settings()
    bands = sharding.row_bands(sharding.valid_cells_per_row(dem.data, dem.flag), world)
    r0, r1 = bands[rank]
    # this rank grids ONLY its band: a DEM view whose other rows are masked stands in for (rowBegin, rowEnd)
    masked = dem.data.copy()
    masked[:r0] = dem.flag
    masked[r1:] = dem.flag
    band_dem = A.Raster(masked, dem.llx, dem.lly, dem.cell_size, dem.flag)
    ok, grid, mn, mx, nv, sec = port_oracle.interpolation_raster(st, s, A.airRelHumidity, band_dem, ())
    full = sharding.gather_bands(grid, (r0, r1), dist, rank, world, dem.flag)
    slowest = sharding.max_over_ranks(1.0 + rank, dist, world)
    assert slowest == float(world)
    steps = sharding.time_steps_for_rank(10, rank, world)
    counts = [None] * world
    dist.all_gather_object(counts, len(steps))
    assert sum(counts) == 10
    if rank == 0:
        okf, want, *_ = port_oracle.interpolation_raster(st, s, A.airRelHumidity, dem, ())
        np.save(result_path, np.array([np.array_equal(full, want)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_row_band_sharding_world_size_2_gloo(tmp_path):
    result = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(2, _free_port(), result), nprocs=2, join=True)
    assert bool(np.load(result)[0])
