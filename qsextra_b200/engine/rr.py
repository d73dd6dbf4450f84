"""Map a lowered step program onto the register-resident kernel (include/qsx.h, qsx_rr_args) when the block is
small and all its dimensions are powers of two.

Index bits of an element a*DB + b (logical order, low to high):
    bra pseudomode bits | bra configuration bits | ket pseudomode bits | ket configuration bits
Placement: configuration bits always go to REGISTER bits (so that the conditions "site occupied" are uniform per
register), then pseudomode (ket, bra) pairs while they fit; the remaining pseudomode bits live on the lanes of the team.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .program import BoundProgram, Damp, StepProgram, _Diag, _Hop, _Prot
from .sectors import Block, sector_rank

RR_DIAG, RR_ROT, RR_AD, RR_ROTT = 0, 1, 2, 3
MAX_OPS, MAX_OBS = 24, 8


@dataclass
class RRProgram:
    block: Block
    reg_bits: int
    lane_bits: int
    n_diag: int
    ops: np.ndarray             # [n_ops, 8] int32 (qsx_rr_op; reg_mask stored as uint32 bit pattern)
    logical_index: np.ndarray   # [E] int32: physical -> logical
    labels: list
    # recipe of the parameter sets: n_diag tables (physical order) followed by the op parameters
    bound: BoundProgram = None        # source of op_params: [B, S] float64 (cos/sin pairs, GA/GB tables), evaluated lazily
    diag_specs: list = None           # per diagonal table: (slot GA, slot GB, slot T or -1) into op_params
    a_of: np.ndarray = None           # [E] ket index of every physical element
    b_of: np.ndarray = None
    fold_classes: list = None         # [(slot of the cosine in op_params, exponent per physical element [E])]
    folded_slots: list = None         # op_params slots whose {cos, sin} become {cos, tan}
    _params: np.ndarray = None
    _shared: dict = None              # results that depend on the structure only, shared by programs of equal structure

    @property
    def op_params(self) -> np.ndarray:
        return self.bound.params

    @property
    def stride(self) -> int:
        return 2 * self.n_diag * self.block.size + self.bound.stride

    @property
    def params(self) -> np.ndarray:
        """[B, stride] float64 on the host (tests, table generators); the runtime builds the same array on the device
        with ``device_params`` (the gathers over [B, E] are far cheaper there)."""
        if self._params is None:
            blk, nb, E = self.block, self.block.n_bits, self.block.size
            B = self.op_params.shape[0]
            op_params = self.op_params.copy()
            out = np.empty((B, self.stride))
            for d, (sa, sb, st) in enumerate(self.diag_specs):
                ga = np.ascontiguousarray(op_params[:, sa:sa + 2 * blk.dim_a]).view(np.complex128)
                gb = np.ascontiguousarray(op_params[:, sb:sb + 2 * blk.dim_b]).view(np.complex128)
                w = ga[:, self.a_of] * gb.conj()[:, self.b_of]
                if st >= 0:
                    t = op_params[:, st:st + blk.n_a * blk.n_b].reshape(B, blk.n_a, blk.n_b)
                    w = w * t[:, self.a_of >> nb, self.b_of >> nb]
                if d == 0:
                    for slot, count in self.fold_classes:
                        powers = np.stack([op_params[:, slot] ** p for p in range(int(count.max()) + 1)], axis=1)
                        w = w * powers[:, count]
                out[:, 2 * d * E:2 * (d + 1) * E] = np.ascontiguousarray(w).view(np.float64)
            for slot in self.folded_slots:
                op_params[:, slot + 1] = op_params[:, slot + 1] / op_params[:, slot]
            out[:, 2 * self.n_diag * E:] = op_params
            self._params = out
        return self._params

    def device_params(self, eng):
        """Same array as ``params`` assembled on the device by one ``qsx_rr_tables`` launch from the small
        op-parameter array (index tables are uploaded once per structure)."""
        import torch
        shared = self._shared if self._shared is not None else {}
        key = ('dev', eng.device.index)
        if key not in shared:
            shared[key] = (eng.to_device(self.a_of, torch.int32), eng.to_device(self.b_of, torch.int32))
        a_of, b_of = shared[key]
        counts = None
        if self.fold_classes:
            ckey = ('dev_counts', eng.device.index, tuple(np.asarray(c).tobytes() for _, c in self.fold_classes))
            if ckey not in shared:
                shared[ckey] = eng.to_device(np.stack([c for _, c in self.fold_classes]).astype(np.int8), torch.int8)
            counts = shared[ckey]
        return eng.rr_tables(self, eng.bind_params(self.bound), a_of, b_of, counts)

    @property
    def team(self):
        return 1 << self.lane_bits

    def observables(self, ptr, idx, w):
        """CSR read-outs (logical indices) -> dense weights in physical order + per-read-out register masks."""
        key = ('obs', np.asarray(ptr).tobytes(), np.asarray(idx).tobytes(), np.asarray(w, dtype=complex).tobytes())
        if self._shared is not None and key in self._shared:
            return self._shared[key]
        out = self._observables(ptr, idx, w)
        if self._shared is not None:
            self._shared[key] = out
        return out

    def _observables(self, ptr, idx, w):
        E = self.block.size
        n_obs = len(ptr) - 1
        phys_of = np.empty(E, dtype=np.int64)
        phys_of[self.logical_index] = np.arange(E)
        dense = np.zeros((n_obs, E), dtype=complex)
        masks = np.zeros(n_obs, dtype=np.uint32)
        nr = 1 << self.reg_bits
        for o in range(n_obs):
            for k in range(ptr[o], ptr[o + 1]):
                ph = phys_of[idx[k]]
                dense[o, ph] += w[k]
                masks[o] |= np.uint32(1 << (ph % nr))
        return dense, masks, self._factorise(dense)

    def executed_flops_per_step(self):
        """(FP64 flops, FP64 instructions) one instance executes per step in the register-resident kernels, counted
        from the operation list: DIAG = complex multiply (2 DMUL + 2 DFMA); rotation on a participating element =
        2 DFMA (tangent form) or 2 DMUL + 2 DFMA; damping = 2 DFMA + 6 DMUL per in-register quad, DMUL + DFMA per
        component when the partner comes from another lane."""
        E, nr = self.block.size, 1 << self.reg_bits
        team = E // nr
        flops = instr = 0
        for op in self.ops:
            kind, rx, lx, rmask = int(op[0]), int(op[1]), int(op[2]), int(op[3]) & 0xffffffff
            if kind == RR_DIAG:
                flops, instr = flops + 6 * E, instr + 4 * E
            elif kind in (RR_ROT, RR_ROTT):
                n = bin(rmask).count('1') * team
                flops, instr = flops + (4 if kind == RR_ROTT else 6) * n, instr + (2 if kind == RR_ROTT else 4) * n
            elif kind == RR_AD:
                if lx == 0:
                    flops, instr = flops + 10 * (E // 4), instr + 8 * (E // 4)
                elif rx == 0:
                    flops, instr = flops + 6 * E, instr + 4 * E
                else:
                    flops, instr = flops + 8 * (E // 2), instr + 6 * (E // 2)
        return flops, instr

    def site_pieces(self, kind: int, n_sites: int):
        """(register mask, lane mask) pieces per site of the two known read-outs (1: populations, 2: emission
        trace), in the order the compile-time tables use (tools/gen_rr_programs.py), or None."""
        key = ('pieces', kind, n_sites)
        if self._shared is not None and key in self._shared:
            return self._shared[key]
        out = self._site_pieces(kind, n_sites)
        if self._shared is not None:
            self._shared[key] = out
        return out

    def _site_pieces(self, kind: int, n_sites: int):
        from .sectors import population_observables
        blk = self.block
        if kind == 1 and blk.ka == blk.kb:
            _, _, ent = self.observables(*population_observables(blk))
            return None if ent is None else [sorted((rm, lm) for rm, lm, _ in e) for e in ent]
        if kind == 2 and abs(blk.ka - blk.kb) == 1:
            marks = np.arange(1, n_sites + 1, dtype=float)          # distinct weights identify the sites
            _, _, ent = self.observables(*blk.emission_observable(marks))
            return None if ent is None else [sorted((rm, lm) for rm, lm, w in ent[0] if w == marks[s]) for s in range(n_sites)]
        return None

    def _factorise(self, dense, max_entries: int = 4):
        """Read-outs as sums over (register mask x lane mask) pieces with one real weight each, or None."""
        nr, team = 1 << self.reg_bits, 1 << self.lane_bits
        out = []
        for row in dense:
            if np.any(row.imag != 0):
                return None
            grid = row.real.reshape(team, nr)
            entries = []
            for wgt in np.unique(grid[grid != 0]):
                pattern = grid == wgt
                groups = {}
                for lane in range(team):
                    rm = int(sum(1 << r for r in range(nr) if pattern[lane, r]))
                    if rm:
                        groups[rm] = groups.get(rm, 0) | (1 << lane)
                entries += [(rm, lm, float(wgt)) for rm, lm in groups.items()]
            if not entries or len(entries) > max_entries:
                return None
            out.append(entries)
        return out


def _fold_cosines(ops, phases, E, NR, has_table, shift, shared=None):
    """Tangent form where it is exact (engine/fold.py): the cos factors of the foldable rotations are to be multiplied
    into the first diagonal table, element by element following each rotation's participation mask, and those
    rotations become QSX_RR_ROTT: two fused multiply-adds per element and side instead of two multiplies and two.
    ``phases(slots)`` gives the rotation angles [len(slots), B] behind op-parameter columns (slots relative to the op
    parameters = op slot - shift); cosines are compared and bounded through the angles, without trigonometry.  Returns (fold_classes,
    folded_slots): [(cosine slot, exponent per element)] and the slots whose sine becomes a tangent.  Nothing is folded
    when the step has no diagonal table; rotations with some |angle| > pi/3 (|cos| < 0.5) are never folded."""
    from .fold import foldable, value_classes
    if not has_table or not any(op[0] == RR_ROT for op in ops):
        return [], []
    phys = np.arange(E)
    reg, lane = phys & (NR - 1), phys // NR
    rot_idx = [k for k, op in enumerate(ops) if op[0] == RR_ROT]
    col = lambda op: op[4] - shift
    angles = np.abs(phases([col(ops[k]) for k in rot_idx]))            # cos is even: +-angle give the same cosine
    cls_of = dict(zip(rot_idx, value_classes(list(angles))))
    # |angle| <= pi/3 guarantees |cos| >= 0.5 (larger angles are simply not folded)
    well = {k: bool(angles[pos].max() <= np.pi / 3) for pos, k in enumerate(rot_idx)}
    info = []
    for k, op in enumerate(ops):
        partner = ((lane ^ op[2]) * NR) | (reg ^ op[1])
        if op[0] == RR_ROT:
            ok = well[k]
            info.append(dict(mixes=True, takes_part=((op[3] >> reg) & 1).astype(bool), partner=partner,
                             cls=cls_of[k] if ok else None))
        elif op[0] == RR_AD:
            clear = ((reg & op[1]) == 0) & ((lane & op[6]) == 0) & ((lane & op[7]) == 0)
            info.append(dict(mixes=True, takes_part=clear, partner=partner, cls=None))
        else:
            info.append(dict(mixes=False, takes_part=np.zeros(E, bool), partner=phys, cls=None))
    pattern = ('fold', tuple(op['cls'] for op in info))
    if shared is not None and pattern in shared:
        fold = set(shared[pattern])
    else:
        fold = foldable(E, info)
        if shared is not None:
            shared[pattern] = frozenset(fold)
    unfolded_slots = {ops[k][4] for k in rot_idx if k not in fold}
    fold = {k for k in fold if ops[k][4] not in unfolded_slots}        # a slot is either all tangent or all sine
    if not fold:
        return [], []
    by_class = {}
    for k in fold:
        ops[k][0] = RR_ROTT
        entry = by_class.setdefault(cls_of[k], [col(ops[k]), np.zeros(E, dtype=np.int64)])
        entry[1] += info[k]['takes_part']
    return [(slot, count) for slot, count in by_class.values()], sorted({col(ops[k]) for k in fold})


def _log2(v: int):
    return v.bit_length() - 1 if v > 0 and v & (v - 1) == 0 else None


def eligible(program: StepProgram, block: Block, max_elements: int = 512) -> bool:
    if _log2(block.n_a) is None or _log2(block.n_b) is None or block.size > max_elements:
        return False
    n_diag = 0
    for op in program.lowered:
        if isinstance(op, _Diag):
            n_diag += 1
        elif isinstance(op, (_Hop, Damp)):
            continue
        elif isinstance(op, _Prot):
            if bin(op.xb).count('1') != 1 or op.zb != 0:
                return False
            if op.site >= 0 and np.any(op.plain + op.signed != 0):
                return False                       # a rotation that also acts on an empty site
        else:
            return False
    return n_diag <= 2


def _skeleton(program: StepProgram, bound: BoundProgram, reg_bits: int | None):
    """Everything of the mapping that depends on the structure of the step only (not on parameter values)."""
    block = bound.block
    nb, cA, cB = block.n_bits, _log2(block.n_a), _log2(block.n_b)
    n = 2 * nb + cA + cB
    # logical bit numbers
    bra_mode = list(range(nb))
    bra_cfg = list(range(nb, nb + cB))
    ket_mode = list(range(nb + cB, 2 * nb + cB))
    ket_cfg = list(range(2 * nb + cB, n))
    if reg_bits is None:
        reg_bits = min(4 if max(cA, cB) <= 1 else 3, n)
    reg_bits = min(reg_bits, n)
    if cA + cB > reg_bits or n - reg_bits > 5:
        return {'none': True}
    order = ket_cfg + bra_cfg                       # physical register bits, low to high
    rest = []
    for j in range(nb):
        if len(order) + 2 <= reg_bits:
            order += [ket_mode[j], bra_mode[j]]
        else:
            rest += [ket_mode[j], bra_mode[j]]
    while len(order) < reg_bits and rest:
        order.append(rest.pop(0))
    order += rest                                   # positions >= reg_bits are lane bits
    place = {lb: pos for pos, lb in enumerate(order)}
    if max(cA, cB) >= 2 and reg_bits > 3:
        return {'none': True}
    E, NR = block.size, 1 << reg_bits
    phys = np.arange(E)
    logical = np.zeros(E, dtype=np.int64)
    for pos, lb in enumerate(order):
        logical |= ((phys >> pos) & 1) << lb
    DB = block.dim_b
    a_of, b_of = logical // DB, logical % DB

    def split(mask_logical_bits):
        m = sum(1 << place[lb] for lb in mask_logical_bits)
        return m & (NR - 1), m >> reg_bits

    def reg_cfg_values(side):
        """configuration index of every register (cfg bits are register bits)."""
        bits = ket_cfg if side == 0 else bra_cfg
        out = np.zeros(NR, dtype=np.int64)
        for k, lb in enumerate(bits):
            out |= ((np.arange(NR) >> place[lb]) & 1) << k
        return out

    base = 0                                        # filled below: tables come first
    tables, ops, labels = [], [], []
    ranks = (sector_rank(block.n_sites, block.ka), sector_rank(block.n_sites, block.kb))
    cfgs = (block.cfg_a, block.cfg_b)
    lowered = program.lowered
    slot_of = {}
    cursor = 0
    for idx, op in enumerate(lowered):             # recover parameter slots in the same order bind() assigned them
        row = bound.ops[cursor]
        slot_of[idx] = row
        cursor += 1
    n_diag = sum(isinstance(op, _Diag) for op in lowered)
    shift = 2 * n_diag * E
    for idx, op in enumerate(lowered):
        row = slot_of[idx]
        if isinstance(op, _Diag):
            ops.append([RR_DIAG, 0, 0, 0, 0, 1, len(tables), 0])
            tables.append((int(row[1]), int(row[2]), int(row[3])))
            labels.append('diag')
        elif isinstance(op, _Hop):
            slot = int(row[6])
            flip = (1 << op.i) | (1 << op.j)
            for side in (0, 1):
                cfg_bits = ket_cfg if side == 0 else bra_cfg
                regcfg = reg_cfg_values(side)
                groups = {}
                for p, mask in enumerate(cfgs[side]):
                    if (mask >> op.i) & 1 and not (mask >> op.j) & 1:
                        q = int(ranks[side][mask ^ flip])
                        x = p ^ q
                        groups.setdefault(x, set()).update((p, q))
                for x, members in groups.items():
                    rx, lx = split([cfg_bits[k] for k in range(len(cfg_bits)) if (x >> k) & 1])
                    rmask = sum(1 << r for r in range(NR) if int(regcfg[r]) in members)
                    ops.append([RR_ROT, rx, lx, rmask, shift + slot, 1 if side == 0 else -1, 0, 0])
                    labels.append('hop{}({},{})'.format('kb'[side], op.i, op.j))
        elif isinstance(op, _Prot):
            slot = int(row[6]) + (2 if op.site >= 0 else 0)
            j = op.xb.bit_length() - 1
            for side in (0, 1):
                rx, lx = split([(ket_mode if side == 0 else bra_mode)[j]])
                regcfg = reg_cfg_values(side)
                if op.site >= 0:
                    rmask = sum(1 << r for r in range(NR) if (cfgs[side][int(regcfg[r])] >> op.site) & 1)
                else:
                    rmask = (1 << NR) - 1
                if rmask:
                    ops.append([RR_ROT, rx, lx, rmask, shift + slot, 1 if side == 0 else -1, 0, 0])
                    labels.append('rot{}({})'.format('kb'[side], j))
        elif isinstance(op, Damp):
            slot = int(row[6])
            pa, pb = place[ket_mode[op.bit]], place[bra_mode[op.bit]]
            rx = (1 << pa if pa < reg_bits else 0) | (1 << pb if pb < reg_bits else 0)
            la = 1 << (pa - reg_bits) if pa >= reg_bits else 0
            lb = 1 << (pb - reg_bits) if pb >= reg_bits else 0
            ops.append([RR_AD, rx, la | lb, 0, shift + slot, 1, la, lb])
            labels.append('damp({})'.format(op.bit))
    if len(ops) > MAX_OPS:
        return {'none': True}
    return dict(ops=ops, tables=tables, labels=labels, logical=logical.astype(np.int32), a_of=a_of, b_of=b_of, shift=shift,
                NR=NR, reg_bits=reg_bits, lane_bits=n - reg_bits, n_diag=n_diag, shared={})


_SKELETONS: dict = {}


def build(program: StepProgram, bound: BoundProgram, reg_bits: int | None = None) -> RRProgram | None:
    block = bound.block
    if not eligible(program, block):
        return None
    key = (program.structure_key(), block, reg_bits)
    sk = _SKELETONS.get(key)
    if sk is None:
        if len(_SKELETONS) > 256:
            _SKELETONS.clear()
        sk = _SKELETONS[key] = _skeleton(program, bound, reg_bits)
    if sk.get('none'):
        return None
    ops = [list(op) for op in sk['ops']]
    fold_classes, folded_slots = _fold_cosines(ops, bound.phases, block.size, sk['NR'], bool(sk['tables']), sk['shift'],
                                               sk['shared'])
    return RRProgram(block, sk['reg_bits'], sk['lane_bits'], sk['n_diag'],
                     np.array(ops, dtype=np.int64).astype(np.uint32).view(np.int32).reshape(-1, 8), sk['logical'],
                     sk['labels'], bound=bound, diag_specs=sk['tables'], a_of=sk['a_of'], b_of=sk['b_of'],
                     fold_classes=fold_classes, folded_slots=folded_slots, _shared=sk['shared'])
