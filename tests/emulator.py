"""TEST INFRASTRUCTURE -- NumPy emulation of the engine's ``qsx_op`` semantics (include/qsx.h) on one block.

Lets the CPU-only test-suite check the host-side compiler (``qsextra_b200/engine/program.py``) against the oracle
without a GPU.  Not importable from the product; the product has no CPU path.
"""
import numpy as np

from qsextra_b200.engine.program import OP_DIAG, OP_HOP, OP_PROT, OP_AD, OP_RESET
from qsextra_b200.engine.sectors import sector_rank


def _popc(v):
    v = np.asarray(v, dtype=np.int64)
    out = np.zeros_like(v)
    for q in range(32):
        out += (v >> q) & 1
    return out


def apply_step(sigma, bound, set_index=0):
    """sigma [DA, DB] complex -> after one step of ``bound`` (BoundProgram) with parameter set ``set_index``."""
    blk = bound.block
    nb, DA, DB = blk.n_bits, blk.dim_a, blk.dim_b
    P = bound.params[set_index]
    cfg_a = np.array(blk.cfg_a, dtype=np.int64)[np.arange(DA) >> nb]
    cfg_b = np.array(blk.cfg_b, dtype=np.int64)[np.arange(DB) >> nb]
    rank_a, rank_b = sector_rank(blk.n_sites, blk.ka), sector_rank(blk.n_sites, blk.kb)
    ia, ib = np.arange(DA), np.arange(DB)
    bits = (1 << nb) - 1
    sigma = sigma.copy()
    for op in bound.ops:
        typ, i0, i1, i2, i3, i4, p, _ = (int(v) for v in op)
        if typ == OP_DIAG:
            ga = P[i0:i0 + 2 * DA].reshape(DA, 2) @ [1, 1j]
            gb = P[i1:i1 + 2 * DB].reshape(DB, 2) @ [1, 1j]
            sigma = sigma * ga[:, None] * gb.conj()[None, :]
            if i2 >= 0:
                t = P[i2:i2 + blk.n_a * blk.n_b].reshape(blk.n_a, blk.n_b)
                sigma = sigma * t[ia[:, None] >> nb, ib[None, :] >> nb]
        elif typ == OP_HOP:
            c, s = P[p], P[p + 1]
            flip = (1 << i0) | (1 << i1)
            for side, cfg, rank, idx in ((0, cfg_a, rank_a, ia), (1, cfg_b, rank_b, ib)):
                sel = idx[((cfg >> i0) & 1 == 1) & ((cfg >> i1) & 1 == 0)]
                partner = (rank[cfg[sel] ^ flip].astype(np.int64) << nb) | (sel & bits)
                f = -1j * s if side == 0 else 1j * s
                if side == 0:
                    x, y = sigma[sel].copy(), sigma[partner].copy()
                    sigma[sel], sigma[partner] = c * x + f * y, c * y + f * x
                else:
                    x, y = sigma[:, sel].copy(), sigma[:, partner].copy()
                    sigma[:, sel], sigma[:, partner] = c * x + f * y, c * y + f * x
        elif typ == OP_PROT:
            xm, zm, ny, site, skip0 = i0, i1, i2, i3, i4
            unit = (-1j) * (1j ** ny)
            for side, cfg, idx in ((0, cfg_a, ia), (1, cfg_b, ib)):
                occ = ((cfg >> site) & 1) if site >= 0 else np.zeros_like(cfg)
                c = np.where(occ == 1, P[p + 2], P[p])
                s = np.where(occ == 1, P[p + 3], P[p + 1])
                if skip0:
                    c, s = np.where(occ == 1, c, 1.0), np.where(occ == 1, s, 0.0)
                sign = 1 - 2 * (_popc((idx ^ xm) & zm) & 1)       # (-1)^{|a2 & z|}, a2 = a ^ x
                f = unit * s * sign                               # -i s P[a, a^x]
                if side == 0:
                    sigma = c[:, None] * sigma + f[:, None] * sigma[idx ^ xm, :]
                else:
                    sigma = sigma * c[None, :] + sigma[:, idx ^ xm] * f.conj()[None, :]
        elif typ in (OP_AD, OP_RESET):
            m = 1 << i0
            c, s2 = (0.0, 1.0) if typ == OP_RESET else (P[p], P[p + 1])
            a0, b0 = ia[(ia & m) == 0], ib[(ib & m) == 0]
            new = np.zeros_like(sigma)
            new[np.ix_(a0, b0)] = sigma[np.ix_(a0, b0)] + s2 * sigma[np.ix_(a0 | m, b0 | m)]
            new[np.ix_(a0, b0 | m)] = c * sigma[np.ix_(a0, b0 | m)]
            new[np.ix_(a0 | m, b0)] = c * sigma[np.ix_(a0 | m, b0)]
            new[np.ix_(a0 | m, b0 | m)] = c * c * sigma[np.ix_(a0 | m, b0 | m)]
            sigma = new
        else:
            raise ValueError(typ)
    return sigma


def apply_dipole(sigma, in_blk, out_blk, side, mu):
    """Collective dipole operator between sector blocks (qsx_apply_dipole)."""
    nb = in_blk.n_bits
    bits = (1 << nb) - 1
    out = np.zeros((out_blk.dim_a, out_blk.dim_b), dtype=complex)
    if side == 0:
        rank_in = sector_rank(in_blk.n_sites, in_blk.ka)
        for a in range(out_blk.dim_a):
            cfg = out_blk.cfg_a[a >> nb]
            for s in range(in_blk.n_sites):
                r = rank_in[cfg ^ (1 << s)]
                if r >= 0:
                    out[a, :] += mu[s] * sigma[(r << nb) | (a & bits), :]
    else:
        rank_in = sector_rank(in_blk.n_sites, in_blk.kb)
        for b in range(out_blk.dim_b):
            cfg = out_blk.cfg_b[b >> nb]
            for s in range(in_blk.n_sites):
                r = rank_in[cfg ^ (1 << s)]
                if r >= 0:
                    out[:, b] += mu[s] * sigma[:, (r << nb) | (b & bits)]
    return out


def response_function(program, mu, side_seq, direction_seq, step_counts):
    """CPU emulation of ``qsextra_b200.engine.response.response_function`` (same tree of evolutions)."""
    from qsextra_b200.engine.response import interaction_plan
    from qsextra_b200.engine.sectors import Block
    plan = interaction_plan(side_seq, direction_seq)
    shape = [len(c) for c in step_counts]
    signal = np.zeros(shape, dtype=complex)
    n_sites = program.n_sites

    def descend(sigma, block, level, idx):
        side, kind = plan[level]
        if level == len(plan) - 1:
            ptr, oidx, w = block.emission_observable(np.asarray(mu, dtype=float))
            signal[idx] += np.sum(w * sigma.reshape(-1)[oidx])
            return
        delta = +1 if kind == 'X' or (kind == 'raise') == (side == 'k') else -1
        ka, kb = (block.ka + delta, block.kb) if side == 'k' else (block.ka, block.kb + delta)
        new_block = Block(n_sites, program.n_bits, ka, kb)
        if not new_block.valid():
            return
        work = apply_dipole(sigma, block, new_block, 0 if side == 'k' else 1, mu)
        bound = program.bind(new_block)
        order = np.argsort(step_counts[level], kind='stable')
        done = 0
        for j in order:
            for _ in range(step_counts[level][j] - done):
                work = apply_step(work, bound)
            done = step_counts[level][j]
            descend(work, new_block, level + 1, idx + (int(j),))

    block = program.block(0, 0)
    sigma = np.zeros((block.dim_a, block.dim_b), dtype=complex)
    sigma[0, 0] = 1.0
    descend(sigma, block, 0, ())
    return signal
