"""Regroup a lowered step program into the engine's fused passes (include/qsx.h, QSX_PASS_*).

The result is the same map on sigma as the plain operation list (operations are only regrouped where they commute):

  QSX_PASS_CFG   [diagonal table] -> [exciton hops of both sides] -> [dephasing table]      one sweep
  QSX_PASS_MODE  per pseudomode bit: conditioned X-rotation of both sides, then amplitude damping; 1 or 2 bits per sweep
  QSX_PASS_OP    anything else (multi-qubit pseudomode registers, resets): the generic sweep of that operation
"""
from __future__ import annotations

import numpy as np

from .program import Damp, _Diag, _Hop, _Prot
from .sectors import Block, sector_rank

PASS_CFG, PASS_MODE, PASS_OP = 0, 1, 2
_CFG_SHAPES = {(1, 1), (1, 2), (2, 1), (2, 2), (1, 3), (3, 1), (3, 3), (2, 3), (3, 2), (1, 4), (4, 1), (4, 4),
               (4, 6), (6, 4), (1, 6), (6, 1)}


def default_team(block: Block) -> int:
    """Threads per instance: one 16-element pseudomode tile (or ~4 elements without pseudomodes) per thread."""
    per_thread = 8 if block.n_bits >= 2 else 4
    t = 1
    while t < 256 and t * per_thread < block.size:
        t <<= 1
    return t


def _is_simple_rot(op) -> bool:
    return isinstance(op, _Prot) and bin(op.xb).count('1') == 1 and op.zb == 0


def build_passes(lowered, op_slots, block: Block, bit_site: dict, team: int, kb: int | None = None):
    """lowered: list of lowered ops; op_slots[i] = (index of op i in the qsx_op array, its parameter slots dict).
    Returns (passes [n, 12] int32, tiles int32, hops int32)."""
    nb, DA, DB = block.n_bits, block.dim_a, block.dim_b
    passes, tiles, hops = [], [], []
    rank_a, rank_b = sector_rank(block.n_sites, block.ka), sector_rank(block.n_sites, block.kb)
    shape_ok = (block.n_a, block.n_b) in _CFG_SHAPES

    def tile_table(entries):
        off = len(tiles)
        for base, aux in entries:
            tiles.extend((int(base), int(aux)))
        return off, len(entries)

    cfg_tiles = None

    def emit_cfg(pre, hop_ops, post):
        nonlocal cfg_tiles
        if cfg_tiles is None:
            cfg_tiles = tile_table([(ma * DB + mb, ma | (mb << 16)) for ma in range(1 << nb) for mb in range(1 << nb)])
        h0 = len(hops) // 4
        for idx in hop_ops:
            op = lowered[idx]
            flip = (1 << op.i) | (1 << op.j)
            slot = op_slots[idx][1]['cs']
            for side, cfgs, rank in ((0, block.cfg_a, rank_a), (1, block.cfg_b, rank_b)):
                for p, mask in enumerate(cfgs):
                    if (mask >> op.i) & 1 and not (mask >> op.j) & 1:
                        q = int(rank[mask ^ flip])
                        hops.extend((side, min(p, q), max(p, q), slot))
        h1 = len(hops) // 4
        slots_pre = op_slots[pre][1] if pre is not None else {}
        slots_post = op_slots[post][1] if post is not None else {}
        passes.append([PASS_CFG, cfg_tiles[0], cfg_tiles[1], block.n_a, block.n_b,
                       slots_pre.get('ga', -1), slots_pre.get('gb', -1), slots_pre.get('t', -1),
                       slots_post.get('t', -1), h0, h1, 0])

    def emit_op(idx):
        passes.append([PASS_OP, op_slots[idx][0]] + [0] * 10)

    def emit_mode(run):
        rot, ad, order_ok = {}, {}, True
        for idx in run:
            op = lowered[idx]
            if isinstance(op, Damp):
                order_ok &= op.bit not in ad
                ad[op.bit] = idx
            else:
                bit = op.xb.bit_length() - 1
                order_ok &= bit not in rot and bit not in ad         # the rotation must precede the damping
                rot[bit] = idx
        if not order_ok:
            for idx in run:
                emit_op(idx)
            return
        bits = sorted(set(rot) | set(ad))
        width = kb if kb is not None else (2 if (block.size >> 4) >= team and len(bits) >= 2 else 1)
        for g in range(0, len(bits), width):
            group = bits[g:g + width]
            mask = sum(1 << b for b in group)
            entries = []
            for a in range(DA):
                if a & mask:
                    continue
                ca = block.cfg_a[a >> nb]
                for b in range(DB):
                    if b & mask:
                        continue
                    cb = block.cfg_b[b >> nb]
                    flags = 0
                    for pos, bit in enumerate(group):
                        site = lowered[rot[bit]].site if bit in rot else -1
                        if site >= 0:
                            flags |= ((ca >> site) & 1) << (2 * pos)
                            flags |= ((cb >> site) & 1) << (2 * pos + 1)
                    entries.append((a * DB + b, flags))
            off, n = tile_table(entries)
            f = [off, n, len(group), group[0], group[1] if len(group) > 1 else 0]
            f += [op_slots[rot[b]][1]['cs'] if b in rot else -1 for b in group] + [-1] * (2 - len(group))
            f += [op_slots[ad[b]][1]['cs'] if b in ad else -1 for b in group] + [-1] * (2 - len(group))
            skip = 0
            for pos, bit in enumerate(group):
                if bit in rot and op_slots[rot[bit]][1].get('skip0'):
                    skip |= 1 << pos
            passes.append([PASS_MODE] + f + [skip, 0])

    i, n = 0, len(lowered)
    while i < n:
        op = lowered[i]
        if isinstance(op, (_Diag, _Hop)) and shape_ok:
            pre = i if isinstance(op, _Diag) else None
            j = i + (1 if pre is not None else 0)
            hop_ops = []
            while j < n and isinstance(lowered[j], _Hop):
                hop_ops.append(j)
                j += 1
            post = None
            if hop_ops and j < n and isinstance(lowered[j], _Diag) and not lowered[j].zterms:
                post = j
                j += 1
            emit_cfg(pre, hop_ops, post)
            i = j
        elif _is_simple_rot(op) or isinstance(op, Damp):
            run = []
            while i < n and (_is_simple_rot(lowered[i]) or isinstance(lowered[i], Damp)):
                run.append(i)
                i += 1
            emit_mode(run)
        else:
            emit_op(i)
            i += 1
    return (np.array(passes, dtype=np.int32).reshape(-1, 12), np.array(tiles, dtype=np.int32),
            np.array(hops, dtype=np.int32))
