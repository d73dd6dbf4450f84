"""world_size-2 (and 3) gloo runs of the sharding logic: contiguous slices + one all-gather reproduce the full result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qsextra_b200.engine import dist as qdist


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, size, port, n_units, rows_per_unit, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=size)
    try:
        full = torch.arange(n_units * rows_per_unit * 3, dtype=torch.float64).reshape(n_units * rows_per_unit, 3)
        full = torch.complex(full, -2 * full)
        lo, hi = qdist.shard_bounds(n_units)
        local = full[lo * rows_per_unit: hi * rows_per_unit].clone()
        got = qdist.all_gather_rows(local, n_units, rows_per_unit)
        real = qdist.all_gather_rows(local.real.contiguous(), n_units, rows_per_unit)
        ok = torch.equal(got, full) and torch.equal(real, full.real)
        np.save(os.path.join(out_dir, 'ok%d.npy' % rank), np.array([ok, lo, hi]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('size,n_units,rows', [(2, 7, 1), (2, 5, 4), (3, 4, 2), (2, 1, 3), (2, 8, 3), (3, 6, 1)])   # (the last two: equal shares, gathered straight into the result)
def test_shard_and_gather(tmp_path, size, n_units, rows):
    mp.spawn(_worker, args=(size, _free_port(), n_units, rows, str(tmp_path)), nprocs=size, join=True)
    bounds = []
    for r in range(size):
        ok, lo, hi = np.load(tmp_path / ('ok%d.npy' % r))
        assert ok
        bounds.append((int(lo), int(hi)))
    assert bounds[0][0] == 0 and bounds[-1][1] == n_units
    assert all(bounds[i][1] == bounds[i + 1][0] for i in range(size - 1))
    sizes = [hi - lo for lo, hi in bounds]
    assert max(sizes) - min(sizes) <= 1


def test_shard_bounds_single_process():
    assert qdist.world() == (0, 1)
    assert qdist.shard_bounds(10) == (0, 10)
    assert [qdist.shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
