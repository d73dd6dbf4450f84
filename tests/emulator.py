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
        typ, i0, i1, i2, i3, i4, p, flags = (int(v) for v in op)
        transposed = bool(flags & 1)                 # QSX_OPF_TRANSPOSE
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
                if transposed and ny & 1:             # P^T = (-1)^nY P
                    s = -s
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
            if transposed:
                new[np.ix_(a0, b0)] = sigma[np.ix_(a0, b0)]
                new[np.ix_(a0 | m, b0 | m)] = c * c * sigma[np.ix_(a0 | m, b0 | m)] + s2 * sigma[np.ix_(a0, b0)]
            else:
                new[np.ix_(a0, b0)] = sigma[np.ix_(a0, b0)] + s2 * sigma[np.ix_(a0 | m, b0 | m)]
                new[np.ix_(a0 | m, b0 | m)] = c * c * sigma[np.ix_(a0 | m, b0 | m)]
            new[np.ix_(a0, b0 | m)] = c * sigma[np.ix_(a0, b0 | m)]
            new[np.ix_(a0 | m, b0)] = c * sigma[np.ix_(a0 | m, b0)]
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


def apply_pass_program(sigma, bound, set_index=0):
    """One step through ``bound.passes`` (QSX_PASS_* semantics of include/qsx.h), tile by tile."""
    from qsextra_b200.engine.passes import PASS_CFG, PASS_MODE, PASS_OP
    blk = bound.block
    nb, DA, DB = blk.n_bits, blk.dim_a, blk.dim_b
    P = bound.params[set_index]
    s = sigma.copy().reshape(-1)
    for ps in bound.passes:
        typ, f = int(ps[0]), [int(v) for v in ps[1:]]
        if typ == PASS_OP:
            single = type(bound)(blk, bound.ops[f[0]:f[0] + 1], bound.params, [])
            s = apply_step(s.reshape(DA, DB), single, set_index).reshape(-1)
        elif typ == PASS_CFG:
            off, n, TA, TB, sga, sgb, stpre, stpost, h0, h1 = f[:10]
            sA, sB = (1 << nb) * DB, 1 << nb
            for ti in range(n):
                base, mm = bound.tiles[off + 2 * ti], bound.tiles[off + 2 * ti + 1]
                ma, mb = mm & 0xffff, mm >> 16
                idx = np.array([[base + i * sA + j * sB for j in range(TB)] for i in range(TA)])
                v = s[idx]
                if sga >= 0:
                    ga = np.array([complex(P[sga + 2 * ((i << nb) | ma)], P[sga + 2 * ((i << nb) | ma) + 1]) for i in range(TA)])
                    gb = np.array([complex(P[sgb + 2 * ((j << nb) | mb)], P[sgb + 2 * ((j << nb) | mb) + 1]) for j in range(TB)])
                    v = v * ga[:, None] * gb.conj()[None, :]
                    if stpre >= 0:
                        v = v * P[stpre:stpre + TA * TB].reshape(TA, TB)
                for h in range(h0, h1):
                    side, p, q, slot = (int(x) for x in bound.hops[4 * h:4 * h + 4])
                    c, sn = P[slot], P[slot + 1]
                    if side == 0:
                        x, y = v[p].copy(), v[q].copy()
                        v[p], v[q] = c * x - 1j * sn * y, c * y - 1j * sn * x
                    else:
                        x, y = v[:, p].copy(), v[:, q].copy()
                        v[:, p], v[:, q] = c * x + 1j * sn * y, c * y + 1j * sn * x
                if stpost >= 0:
                    v = v * P[stpost:stpost + TA * TB].reshape(TA, TB)
                s[idx] = v
        elif typ == PASS_MODE:
            off, n, KB = f[0], f[1], f[2]
            bits, rots, ads, skip = f[3:3 + KB], f[5:5 + KB], f[7:7 + KB], f[9]
            TS = 1 << KB
            offs = [sum(1 << bits[b] for b in range(KB) if (i >> b) & 1) for i in range(TS)]
            for ti in range(n):
                base, flags = int(bound.tiles[off + 2 * ti]), int(bound.tiles[off + 2 * ti + 1])
                idx = np.array([[base + offs[ia] * DB + offs[ib] for ib in range(TS)] for ia in range(TS)])
                v = s[idx]
                for b in range(KB):
                    if rots[b] >= 0:
                        for side in (0, 1):
                            occ = (flags >> (2 * b + side)) & 1
                            if not occ and (skip >> b) & 1:
                                continue
                            c, sn = P[rots[b] + 2 * occ], P[rots[b] + 2 * occ + 1]
                            for i0 in range(TS):
                                if (i0 >> b) & 1:
                                    continue
                                i1 = i0 | (1 << b)
                                if side == 0:
                                    x, y = v[i0].copy(), v[i1].copy()
                                    v[i0], v[i1] = c * x - 1j * sn * y, c * y - 1j * sn * x
                                else:
                                    x, y = v[:, i0].copy(), v[:, i1].copy()
                                    v[:, i0], v[:, i1] = c * x + 1j * sn * y, c * y + 1j * sn * x
                    if ads[b] >= 0:
                        c, s2 = P[ads[b]], P[ads[b] + 1]
                        new = v.copy()
                        for ia in range(TS):
                            for ib in range(TS):
                                ha, hb = (ia >> b) & 1, (ib >> b) & 1
                                if not ha and not hb:
                                    new[ia, ib] = v[ia, ib] + s2 * v[ia | (1 << b), ib | (1 << b)]
                                elif ha != hb:
                                    new[ia, ib] = c * v[ia, ib]
                                else:
                                    new[ia, ib] = c * c * v[ia, ib]
                        v = new
                s[idx] = v
    return s.reshape(DA, DB)


def apply_rr_program(sigma, rr, set_index=0):
    """One step through an RRProgram (QSX_RR_* semantics of include/qsx.h) on the physical layout."""
    from qsextra_b200.engine.rr import RR_AD, RR_DIAG, RR_ROT, RR_ROTT
    E = rr.block.size
    R, NR = rr.reg_bits, 1 << rr.reg_bits
    P = rr.params[set_index]
    v = sigma.reshape(-1)[rr.logical_index].copy()           # physical order
    phys = np.arange(E)
    reg, lane = phys & (NR - 1), phys >> R
    for op in rr.ops:
        kind, rx, lx, rmask, p, sign, aux0, aux1 = (int(x) for x in op)
        rmask &= 0xffffffff
        partner = ((lane ^ lx) << R) | (reg ^ rx)
        if kind == RR_DIAG:
            w = P[2 * aux0 * E: 2 * (aux0 + 1) * E].reshape(E, 2) @ [1, 1j]
            v = v * w
        elif kind == RR_ROT:
            c, s = P[p], P[p + 1] * sign
            active = ((rmask >> reg) & 1) == 1
            v = np.where(active, c * v - 1j * s * v[partner], v)
        elif kind == RR_ROTT:
            t = P[p + 1] * sign
            active = ((rmask >> reg) & 1) == 1
            v = np.where(active, v - 1j * t * v[partner], v)
        elif kind == RR_AD:
            c, s2 = P[p], P[p + 1]
            nset = _popc(reg & rx) + ((lane & aux0) != 0) + ((lane & aux1) != 0)
            v = np.where(nset == 0, v + s2 * v[partner], np.where(nset == 1, c * v, c * c * v))
    out = np.empty(E, dtype=complex)
    out[rr.logical_index] = v
    return out.reshape(sigma.shape)


def apply_rrc_program(sigma, prog, set_index=0):
    """One step through an RRCProgram (engine/rrc.py; kernel csrc/qsx_rrc.cuh) on the logical layout."""
    from qsextra_b200.engine import rrc as K
    blk = prog.block
    nb, nA, nB = blk.n_bits, blk.n_a, blk.n_b
    M = 1 << nb
    P = prog.params[set_index]
    v = sigma.reshape(nA, M, nB, M).transpose(0, 2, 1, 3).copy()            # [i][j][m_a][m_b]
    n_w = 2 * nA * nB
    w_cfg = P[:n_w].reshape(nA, nB, 2) @ [1, 1j]
    gm_a = P[n_w:n_w + 2 * M].reshape(M, 2) @ [1, 1j]
    gm_b = P[n_w + 2 * M:n_w + 4 * M].reshape(M, 2) @ [1, 1j]
    m = np.arange(M)
    for kind, a, b, mask, slot in prog.ops:
        if kind == K.OP_DIAG:
            v = v * w_cfg[:, :, None, None] * gm_a[None, None, :, None] * gm_b.conj()[None, None, None, :]
            continue
        c, t = P[slot], P[slot + 1]
        if kind == K.OP_HOPK:                        # (cos, sin) form
            x, y = v[a].copy(), v[b].copy()
            v[a], v[b] = c * x - 1j * t * y, c * y - 1j * t * x
        elif kind == K.OP_HOPB:
            x, y = v[:, a].copy(), v[:, b].copy()
            v[:, a], v[:, b] = c * x + 1j * t * y, c * y + 1j * t * x
        elif kind == K.OP_ROTK:
            rows = [i for i in range(nA) if (mask >> i) & 1]
            v[rows] = v[rows] - 1j * t * v[rows][:, :, m ^ (1 << a), :]
        elif kind == K.OP_ROTB:
            cols = [j for j in range(nB) if (mask >> j) & 1]
            v[:, cols] = v[:, cols] + 1j * t * v[:, cols][:, :, :, m ^ (1 << a)]
        elif kind == K.OP_AD:
            s2 = t
            ha, hb = ((m >> a) & 1)[:, None], ((m >> a) & 1)[None, :]
            partner = v[:, :, m ^ (1 << a), :][:, :, :, m ^ (1 << a)]
            v = np.where((ha == 0) & (hb == 0), v + s2 * partner, np.where(ha != hb, c * v, c * c * v))
        elif kind == K.OP_ADT:                       # transposed damping: the (0,0) elements feed the (1,1) elements
            s2 = t
            ha, hb = ((m >> a) & 1)[:, None], ((m >> a) & 1)[None, :]
            partner = v[:, :, m ^ (1 << a), :][:, :, :, m ^ (1 << a)]
            v = np.where((ha == 1) & (hb == 1), c * c * v + s2 * partner, np.where(ha != hb, c * v, v))
    return v.transpose(0, 2, 1, 3).reshape(sigma.shape)


def apply_lindblad(bound, sigma, seg_dt, terms=20, theta=1.0):
    """NumPy emulation of qsx_lindblad_evolve on one block: scaled Taylor series of exp(hL) with the ELL generator.
    -> states after every segment [n_seg, E]."""
    acc = np.asarray(sigma, dtype=complex).reshape(-1).copy()
    out = []
    for dt in seg_dt:
        if dt > 0 and bound.norm1 > 0:
            n_sub = max(1, int(np.ceil(bound.norm1 * dt / theta)))
            h = dt / n_sub
            for _ in range(n_sub):
                y = acc
                total = acc.copy()
                for m in range(1, terms + 1):
                    y = (h / m) * (bound.ell_val * y[bound.ell_idx]).sum(axis=0)
                    total = total + y
                acc = total
        out.append(acc.copy())
    return np.array(out)
