"""Map a lowered step program onto the CTA-wide register-resident kernel (csrc/qsx_rrc.cuh) for LARGE blocks of
systems with one 2-level pseudomode per site (C4: 4 sites, blocks up to 96 x 64).

Thread t of the CTA owns the configuration tile v[i][j] = sigma[(i, m_a), (j, m_b)] for ONE pair of pseudomode
states (m_a, m_b): 4**n_bits threads, nA x nB registers each, for the whole evolution.  Diagonal phases, exciton hops
of both sides and the cosine bookkeeping never leave the thread; pseudomode rotations and damping exchange with the
thread whose (m_a, m_b) differs in one bit (pair) -- by warp shuffle when that bit is a lane bit, through a shared
memory buffer when it is a warp bit.  Only compile-time programs exist for this kernel (tools/gen_rrc_programs.py).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .program import BoundProgram, Damp, StepProgram, _Diag, _Hop, _Prot
from .sectors import Block, sector_rank

OP_DIAG, OP_HOPK, OP_HOPB, OP_ROTK, OP_ROTB, OP_AD, OP_ADT = 0, 1, 2, 3, 4, 5, 6   # ROT*: tangent form; HOP*: (cos, sin)


@dataclass
class RRCProgram:
    block: Block
    n_a: int
    n_b: int
    n_bits: int
    ket_pos: list            # thread-bit position of ket pseudomode bit k
    bra_pos: list
    ops: list                # [kind, a, b, mask, slot]  (HOP: a=p, b=q; ROT: a=bit, mask=rows/cols; AD: a=bit)
    emission: list           # [(i, j, site)] pairs of the read-out (threads with m_a == m_b): emission trace or populations
    params: np.ndarray       # [B, stride]: Wcfg (2 nA nB) | gmA (2 * 2^nb) | gmB (2 * 2^nb) | op parameters
    shift: int               # offset of the op parameters

    @property
    def threads(self):
        return 1 << (2 * self.n_bits)

    def executed_flops_per_step(self):
        """(FP64 flops, FP64 instructions) one instance executes per step in ``rrc::evolve_rrc_kernel``, counted from the
        operation list as the kernel applies it: DIAG = complex multiply per element (2 DMUL + 2 DFMA); HOP on a pair of
        configurations = (cos, sin) rotation of both elements (4 DMUL + 4 DFMA per pair and partner index); ROT on a
        participating element = 2 DFMA (tangent form); damping = DMUL + DFMA per component on the quarter of the
        elements that receives, 2 DMUL on the others."""
        nt = self.threads
        flops = instr = 0
        for kind, a, b, mask, slot in self.ops:
            if kind == OP_DIAG:
                n = self.n_a * self.n_b * nt
                flops, instr = flops + 6 * n, instr + 4 * n
            elif kind in (OP_HOPK, OP_HOPB):
                pairs = (self.n_b if kind == OP_HOPK else self.n_a) * nt
                flops, instr = flops + 12 * pairs, instr + 8 * pairs
            elif kind in (OP_ROTK, OP_ROTB):
                n = bin(mask).count('1') * (self.n_b if kind == OP_ROTK else self.n_a) * nt
                flops, instr = flops + 4 * n, instr + 2 * n
            else:
                n = self.n_a * self.n_b * nt
                flops, instr = flops + (6 * n) // 4 + (2 * n * 3) // 4, instr + (4 * n) // 4 + (2 * n * 3) // 4
        return flops, instr

    def source(self, pid: int = 0, comment: str = '') -> str:
        """The ``BigProg<pid>`` specialisation of csrc/qsx_rrc.cuh for this program (the compile-time menu of
        csrc/qsx_rrc_programs.inc and the run-time specialisations of engine/jit.py are both written by this)."""
        def arr(rows):
            return '{' + ', '.join('{' + ', '.join(str(int(v)) for v in r) + '}' for r in rows) + '}'
        sig = self.signature()
        em = self.emission or [(0, 0, 0)]
        lines = ['// ' + comment] if comment else []
        lines.append('template <> struct BigProg<%d> {' % pid)
        lines.append('    static constexpr int NA = %d, NB = %d, NBITS = %d, NOPS = %d, NEM = %d, NSIG = %d;' % (
            self.n_a, self.n_b, self.n_bits, len(self.ops), len(self.emission), len(sig)))
        lines.append('    static constexpr int ket_pos[%d] = {%s}, bra_pos[%d] = {%s};' % (
            self.n_bits, ', '.join(map(str, self.ket_pos)), self.n_bits, ', '.join(map(str, self.bra_pos))))
        lines.append('    static constexpr int ops[%d][5] = %s;' % (len(self.ops), arr(self.ops)))
        lines.append('    static constexpr int em[%d][3] = %s;' % (len(em), arr(em)))
        lines.append('    static constexpr int sig[%d] = {%s};' % (len(sig), ', '.join(str(int(v)) for v in sig)))
        lines.append('};')
        return '\n'.join(lines) + '\n'

    def throughput_form_eligible(self) -> bool:
        """Mirror of ``rrc2::Geo<PID>::eligible()``: forward program, DIAG first, dampings last, tile <= 16 elements."""
        kinds = [op[0] for op in self.ops]
        if any(op[0] == OP_ROTK and self.ket_pos[op[1]] >= 5 for op in self.ops):
            return False
        seen = 0                                        # bra rotations over warp bits share one exchange round
        for op in self.ops:
            if op[0] == OP_ROTB and self.bra_pos[op[1]] >= 5:
                if seen & op[3]:
                    return False
                seen |= op[3]
        if self.n_a * self.n_b > 16 or not kinds or kinds[0] != OP_DIAG or OP_ADT in kinds or OP_DIAG in kinds[1:]:
            return False
        tail = [k == OP_AD for k in kinds[1:]]
        return all(not (a and not b) for a, b in zip(tail, tail[1:]))

    def signature(self) -> np.ndarray:
        """Everything the compile-time copy must agree on, as one int32 vector."""
        sig = [self.n_a, self.n_b, self.n_bits] + list(self.ket_pos) + list(self.bra_pos) + [len(self.ops)]
        for op in self.ops:
            sig += [int(v) for v in op]
        sig += [len(self.emission)]
        for e in self.emission:
            sig += [int(v) for v in e]
        return np.array(sig, dtype=np.int32)


def build(program: StepProgram, bound: BoundProgram) -> RRCProgram | None:
    """``bound`` may be the transposed program (``BoundProgram.transposed()``): the operation list is then reversed and
    the dampings become OP_ADT; hops, tangent-form rotations and the diagonal table are their own transposes, and the
    folded cosines are one scalar per sector, so their position in the step does not matter."""
    transposed = bool(getattr(bound, 'is_transposed', False))
    if transposed:
        bound = bound._forward
    block = bound.block
    nb, n_sites = block.n_bits, block.n_sites
    if nb < 3 or nb > 4 or nb != n_sites or block.n_a * block.n_b > 32:
        return None                               # one 2-level pseudomode per site, 64 or 256 threads
    lowered = program.lowered
    if sum(isinstance(op, _Diag) for op in lowered) != 1 or not isinstance(lowered[0], _Diag) or lowered[0].deph:
        return None
    B = bound.params.shape[0]
    # thread bits: ket pseudomode bits and bra bit 0 on the lanes, the other bra bits select the warp
    order = [('k', k) for k in range(nb)] + [('b', k) for k in range(nb)]
    ket_pos = [order.index(('k', k)) for k in range(nb)]
    bra_pos = [order.index(('b', k)) for k in range(nb)]
    ranks = (sector_rank(n_sites, block.ka), sector_rank(n_sites, block.kb))
    cfgs = (block.cfg_a, block.cfg_b)
    ops, rot_slots = [], []
    for idx, op in enumerate(lowered):
        row = bound.ops[idx]
        if isinstance(op, _Diag):
            ops.append([OP_DIAG, 0, 0, 0, 0])
        elif isinstance(op, _Hop):
            slot = int(row[6])
            flip = (1 << op.i) | (1 << op.j)
            for side, kind in ((0, OP_HOPK), (1, OP_HOPB)):
                for p, mask in enumerate(cfgs[side]):
                    if (mask >> op.i) & 1 and not (mask >> op.j) & 1:
                        q = int(ranks[side][mask ^ flip])
                        ops.append([kind, min(p, q), max(p, q), 0, slot])
            rot_slots.append((slot, 'hop', op))
        elif isinstance(op, _Prot):
            if bin(op.xb).count('1') != 1 or op.zb != 0 or op.site < 0 or np.any(op.plain + op.signed != 0):
                return None
            bit = op.xb.bit_length() - 1
            slot = int(row[6]) + 2
            for side, kind in ((0, OP_ROTK), (1, OP_ROTB)):
                mask = sum(1 << c for c, m in enumerate(cfgs[side]) if (m >> op.site) & 1)
                if mask:
                    ops.append([kind, bit, 0, mask, slot])
            rot_slots.append((slot, 'rot', op))
        elif isinstance(op, Damp):
            ops.append([OP_AD, op.bit, 0, 0, int(row[6])])
        else:
            return None
    # parameters: configuration part and pseudomode part of the diagonal table, then the op parameters
    row = bound.ops[0]
    sa, sb = int(row[1]), int(row[2])
    ga = np.ascontiguousarray(bound.params[:, sa:sa + 2 * block.dim_a]).view(np.complex128)
    gb = np.ascontiguousarray(bound.params[:, sb:sb + 2 * block.dim_b]).view(np.complex128)
    g_cfg_a, g_cfg_b = ga[:, ::1 << nb], gb[:, ::1 << nb]                       # bits = 0
    gm_a = ga[:, :1 << nb] * ga[:, :1].conj()                                    # configuration 0, relative to bits = 0
    gm_b = gb[:, :1 << nb] * gb[:, :1].conj()
    w_cfg = g_cfg_a[:, :, None] * g_cfg_b.conj()[:, None, :]                     # [B, nA, nB]
    shift = 2 * block.n_a * block.n_b + 4 * (1 << nb)
    params = np.empty((B, shift + bound.params.shape[1]))
    params[:, shift:] = bound.params
    # tangent form for the pseudomode rotations: their partners share the configuration pair, so the cosines can be
    # folded into w_cfg -- provided every hop pairs configurations that carry the same pending factor (same number
    # of occupied sites and one coupling constant: true inside a sector; checked).  Hops keep the (cos, sin) form.
    rot_slots_only = sorted({s for s, kind, _ in rot_slots if kind == 'rot'})
    rot_cos = params[:, [shift + s for s in rot_slots_only]]
    if np.any(np.abs(rot_cos) < 0.5) or np.any(rot_cos != rot_cos[:, :1]):
        return None                               # one coupling constant for all sites is what makes the fold exact
    pend_a = np.array([bin(m).count('1') for m in block.cfg_a])
    pend_b = np.array([bin(m).count('1') for m in block.cfg_b])
    if len(set(pend_a)) > 1 or len(set(pend_b)) > 1:
        return None
    for op in ops:
        kind, a, b, mask, slot = op
        c = params[:, shift + slot]
        if kind == OP_ROTK:
            rows = [i for i in range(block.n_a) if (mask >> i) & 1]
            w_cfg[:, rows, :] *= c[:, None, None]
        elif kind == OP_ROTB:
            cols = [j for j in range(block.n_b) if (mask >> j) & 1]
            w_cfg[:, :, cols] *= c[:, None, None]
    for s in rot_slots_only:
        params[:, shift + s + 1] = params[:, shift + s + 1] / params[:, shift + s]
    n_w = 2 * block.n_a * block.n_b
    params[:, :n_w] = np.ascontiguousarray(w_cfg.reshape(B, -1)).view(np.float64)
    params[:, n_w:n_w + 2 * (1 << nb)] = np.ascontiguousarray(gm_a).view(np.float64)
    params[:, n_w + 2 * (1 << nb):shift] = np.ascontiguousarray(gm_b).view(np.float64)
    for op in ops:
        op[4] += shift if op[0] != OP_DIAG else 0
    # the damping-diagonal forms divide by the product of the damping cosines at emissions: keep them away from
    # collisions that empty a pseudomode (almost) completely in one step (the general kernel serves those)
    damp_cos = [params[:, op[4]] for op in ops if op[0] == OP_AD]
    if damp_cos and float(np.min(np.abs(damp_cos))) < 1e-3:
        return None
    if transposed:
        ops = [[OP_ADT if op[0] == OP_AD else op[0]] + op[1:] for op in reversed(ops)]
    emission = []
    if abs(block.ka - block.kb) == 1:
        for j, mask_b in enumerate(block.cfg_b):
            for s in range(n_sites):
                i = ranks[0][mask_b ^ (1 << s)]
                if i >= 0:
                    emission.append((int(i), j, s))
    elif block.ka == block.kb:                        # populations: diagonal pairs (i, i) with site s occupied
        for i, mask in enumerate(block.cfg_a):
            emission += [(i, i, s) for s in range(n_sites) if (mask >> s) & 1]
    return RRCProgram(block, block.n_a, block.n_b, nb, ket_pos, bra_pos, ops, emission, params, shift)
