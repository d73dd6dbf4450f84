"""Sharding of independent units (parameter sets, delay-grid chains) over the ranks of one node.

One process per GPU (``torch.distributed``, backend NCCL on the GPU, gloo in CPU tests).  The path has no exchange
step: every rank evolves a contiguous slice of the units and one all-gather assembles the result.
"""
from __future__ import annotations

import contextlib
import threading

import torch
import torch.distributed as dist

_LOCAL = threading.local()


@contextlib.contextmanager
def local_only():
    """Inside: ``world()`` answers (0, 1).  For work whose units were split over the ranks by the CALLER (an ensemble whose
    members each rank lowered for itself): the tree must not split them again; the caller gathers afterwards."""
    previous = getattr(_LOCAL, 'on', False)
    _LOCAL.on = True
    try:
        yield
    finally:
        _LOCAL.on = previous


def world():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised or inside ``local_only()``."""
    if getattr(_LOCAL, 'on', False):
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_units: int, rank: int | None = None, size: int | None = None):
    """Contiguous slice [lo, hi) of ``n_units`` owned by ``rank`` (balanced to within one unit)."""
    if rank is None or size is None:
        rank, size = world()
    base, extra = divmod(n_units, size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_rows(local: torch.Tensor, n_units: int, rows_per_unit: int = 1) -> torch.Tensor:
    """Concatenate the row-slices of all ranks into [n_units * rows_per_unit, ...].  Rank r holds the rows of the
    units ``shard_bounds(n_units, r)``, ``rows_per_unit`` consecutive rows each.  A single collective: equal slices
    are gathered straight into the result, unequal ones are padded to the largest slice so that
    ``all_gather_into_tensor`` applies."""
    rank, size = world()
    if size == 1:
        return local
    rows = [tuple(v * rows_per_unit for v in shard_bounds(n_units, r, size)) for r in range(size)]
    width = max(hi - lo for lo, hi in rows)
    if all(hi - lo == width for lo, hi in rows):        # equal shares: no padding, no reassembly
        source = local.contiguous()
        gathered = local.new_empty((size * width,) + tuple(local.shape[1:]))
        if source.is_complex():
            dist.all_gather_into_tensor(torch.view_as_real(gathered), torch.view_as_real(source))
        else:
            dist.all_gather_into_tensor(gathered, source)
        return gathered
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    gathered = local.new_empty((size * width,) + tuple(local.shape[1:]))
    if padded.is_complex():
        dist.all_gather_into_tensor(torch.view_as_real(gathered), torch.view_as_real(padded.contiguous()))
    else:
        dist.all_gather_into_tensor(gathered, padded.contiguous())
    return torch.cat([gathered[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(rows)], dim=0)
