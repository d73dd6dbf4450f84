"""Structural support of the states of a response tree (host-side analysis, no arithmetic on amplitudes).

Which elements of a sector block can ever be non-zero, given where the block starts and which operations the step
holds?  The answer is a closure over the coupling graph of the step: exciton hops connect configurations, pseudomode
rotations connect bit patterns (only on configurations where their angle is not identically zero), damping connects
(1, 1) to (0, 0) of one (ket bit, bra bit) pair, diagonal tables connect nothing.  ``engine/response.py`` uses it to
replace MANY chains that enter a delay level by the FEW basis elements of their common support (linearity): after the
first interaction of a rephasing diagram sigma = |g, modes 0><phi| never leaves the 64 elements with empty ket
pseudomodes of its 16 x 64 block, however many t1 values there are.

Follows the operator content of qsextra/qcomo/evolution/qevolve.py:74-348 as lowered by ``engine/program.py``.
"""
from __future__ import annotations

import numpy as np

from .program import Damp, StepProgram, _Diag, _Hop, _Prot
from .sectors import Block, sector_rank

_CACHE: dict = {}


def ground_support(block: Block) -> np.ndarray:
    mask = np.zeros((block.dim_a, block.dim_b), dtype=bool)
    mask[0, 0] = True
    return mask


def dipole_support(mask: np.ndarray, in_block: Block, out_block: Block, side: str) -> np.ndarray:
    """Support of (sum_s mu_s O_s) sigma (side 'k') or sigma (sum_s mu_s O_s) (side 'b'): the configuration of that side
    gains or loses any one site (every site is assumed to carry a dipole), pseudomode bits stay."""
    nb = in_block.n_bits
    m_bits = 1 << nb
    n_sites = in_block.n_sites
    src = mask.reshape(in_block.n_a, m_bits, in_block.n_b, m_bits)
    out = np.zeros((out_block.n_a, m_bits, out_block.n_b, m_bits), dtype=bool)
    ket = side == 'k'
    cfg_in = in_block.cfg_a if ket else in_block.cfg_b
    k_out = out_block.ka if ket else out_block.kb
    rank_out = sector_rank(n_sites, k_out)
    for p, cm in enumerate(cfg_in):
        for s in range(n_sites):
            new = cm ^ (1 << s)
            if bin(new).count('1') != k_out:
                continue
            q = int(rank_out[new])
            if q < 0:
                continue
            if ket:
                out[q] |= src[p]
            else:
                out[:, :, q] |= src[:, :, p]
    return out.reshape(out_block.dim_a, out_block.dim_b)


def step_closure(program: StepProgram, block: Block, mask: np.ndarray) -> np.ndarray:
    """Smallest superset of ``mask`` that every operation of the step maps into itself (so: the support of the block
    after any number of steps).  Unknown operations (resets of literal ancilla bits, ...) give the full block."""
    key = (program.structure_key(), block, mask.tobytes(), _activity(program))
    hit = _CACHE.get(key)
    if hit is not None:
        return hit
    nb = block.n_bits
    m_bits = 1 << nb
    n_sites = block.n_sites
    m = mask.reshape(block.n_a, m_bits, block.n_b, m_bits).copy()
    ranks = (sector_rank(n_sites, block.ka), sector_rank(n_sites, block.kb))
    cfgs = (block.cfg_a, block.cfg_b)
    bits = np.arange(m_bits)
    full = False
    for _ in range(4 * (nb + n_sites) + 8):
        before = int(m.sum())
        for op in program.lowered:
            if isinstance(op, _Diag):
                continue
            if isinstance(op, _Hop):
                flip = (1 << op.i) | (1 << op.j)
                for side in (0, 1):
                    for p, cm in enumerate(cfgs[side]):
                        if (cm >> op.i) & 1 and not (cm >> op.j) & 1:
                            q = int(ranks[side][cm ^ flip])
                            if q < 0:
                                continue
                            if side == 0:
                                both = m[p] | m[q]
                                m[p], m[q] = both, both.copy()
                            else:
                                both = m[:, :, p] | m[:, :, q]
                                m[:, :, p], m[:, :, q] = both, both.copy()
            elif isinstance(op, _Prot):
                act_empty = bool(np.any(op.plain + op.signed != 0)) if op.site >= 0 else bool(np.any(op.plain != 0))
                act_occ = bool(np.any(op.plain - op.signed != 0)) if op.site >= 0 else act_empty
                perm = bits ^ op.xb
                for side in (0, 1):
                    for p, cm in enumerate(cfgs[side]):
                        occupied = op.site >= 0 and (cm >> op.site) & 1
                        if not (act_occ if occupied else act_empty):
                            continue
                        if side == 0:
                            m[p] |= m[p][perm]
                        else:
                            m[:, :, p] |= m[:, :, p][:, :, perm]
            elif isinstance(op, Damp):
                hi = bits[(bits >> op.bit) & 1 == 1]
                lo = hi ^ (1 << op.bit)
                m[:, lo[:, None], :, lo[None, :]] |= m[:, hi[:, None], :, hi[None, :]]
            else:
                full = True
        if full or int(m.sum()) == before:
            break
    out = np.ones_like(mask) if full else m.reshape(block.dim_a, block.dim_b)
    if len(_CACHE) > 256:
        _CACHE.clear()
    _CACHE[key] = out
    return out


def _activity(program: StepProgram) -> tuple:
    """Which conditioned rotations are switched on (their angles are parameters, not structure)."""
    flags = []
    for op in program.lowered:
        if isinstance(op, _Prot):
            flags.append((bool(np.any(op.plain + op.signed != 0)), bool(np.any(op.plain - op.signed != 0)), bool(np.any(op.plain != 0))))
    return tuple(flags)
