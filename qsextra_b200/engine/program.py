"""Host-side compiler: system parameters -> step program of the CUDA engine (``qsx_op`` list + angle tables).

Mirrors what the reference builds with Qiskit objects in ``qevolve.py``:

  :74-105   system block           Z_i (-eps_i/2), then X_iX_j, Y_iY_j (J_ij/2) for i<j, Lie/Suzuki product formula
  :107-195  pseudomode operators   Gray-coded a, a^dagger, a + a^dagger, a^dagger a as Pauli sums
  :197-253  mode evolution, exciton-mode coupling g (a + a^dagger) (I - Z_e)/2
  :255-290  step assembly order    system block; for k: for i: [mode(i,k), coupling(i,k)]
  :292-348  collisions             rzx + reset (dephasing)  |  sqrt(rate/dt)(a sigma+ + a^dagger sigma-) + reset
  :362-428  dt / Trotter_number / step-count logic

but in the engine's vocabulary (include/qsx.h): Pauli strings are (xmask, zmask) pairs, Z-type rotations are
merged into diagonal tables, X_iX_j/Y_iY_j pairs into exciton hops, X_m / X_m Z_e pairs into rotations
conditioned on the site occupation, and the 2-level collision + reset into an amplitude-damping channel.
All parameters carry a leading batch axis so one compile serves many parameter sets of the same structure.
"""
from __future__ import annotations

import functools
import itertools
from dataclasses import dataclass, field

import numpy as np

from .sectors import Block

OP_DIAG, OP_HOP, OP_PROT, OP_AD, OP_RESET = 0, 1, 2, 3, 4
OPF_TRANSPOSE = 1                          # qsx_op.flags: the operation is replaced by its transpose (include/qsx.h)


# ---------------------------------------------------------------------------------------------
# batched system parameters
# ---------------------------------------------------------------------------------------------
@dataclass
class SystemBatch:
    """Parameters of B systems that share one structure (N, pseudomode levels)."""
    energies: np.ndarray                    # [B, N]
    couplings: np.ndarray                   # [B, N, N]
    dipole_moments: np.ndarray              # [B, N]
    omega: np.ndarray | None = None         # [B, W]
    coupl_ep: np.ndarray | None = None      # [B, W]
    levels: tuple[int, ...] | None = None   # W entries

    @property
    def batch(self) -> int:
        return self.energies.shape[0]

    @property
    def n_sites(self) -> int:
        return self.energies.shape[1]

    @property
    def chromophore(self) -> bool:
        return self.levels is not None

    @property
    def n_modes(self) -> int:
        return len(self.levels) if self.levels is not None else 0

    @classmethod
    def from_systems(cls, systems) -> "SystemBatch":
        systems = list(systems)
        first = systems[0]
        md = getattr(first, 'mode_dict', None)
        kw = dict(energies=np.array([np.asarray(s.e_el, dtype=float) for s in systems]),
                  couplings=np.array([np.asarray(s.coupl_el, dtype=float) for s in systems]),
                  dipole_moments=np.array([np.asarray(s.dipole_moments, dtype=float) for s in systems]))
        if md is not None:
            levels = tuple(int(d) for d in md['lvl_mode'])
            for s in systems:
                if s.mode_dict is None or tuple(int(d) for d in s.mode_dict['lvl_mode']) != levels:
                    raise ValueError('all systems of a batch must have the same pseudomode levels')
            kw.update(levels=levels,
                      omega=np.array([np.asarray(s.mode_dict['omega_mode'], dtype=float) for s in systems]),
                      coupl_ep=np.array([np.asarray(s.mode_dict['coupl_ep'], dtype=float) for s in systems]))
        return cls(**kw)


def mode_register_size(d_k: int) -> int:
    """Qubits of a d_k-level pseudomode register (qevolve.py:28-35)."""
    nq = np.log2(d_k)
    return int(np.ceil(nq)) if abs(np.round(nq) - nq) > 1e-10 else int(np.round(nq))


@dataclass
class BitLayout:
    """Positions of pseudomode qubits (and the optional collision ancilla) inside ``bits``; same order as the
    reference's registers mode(0,0), mode(0,1), ..., mode(N-1,W-1) (qevolve.py:25-37)."""
    mode: dict
    ancilla: int | None
    n_bits: int

    @classmethod
    def build(cls, n_sites: int, levels, with_ancilla: bool) -> "BitLayout":
        mode, nxt = {}, 0
        for i in range(n_sites):
            for k, d in enumerate(levels or ()):
                q = mode_register_size(d)
                mode[(i, k)] = list(range(nxt, nxt + q))
                nxt += q
        anc = nxt if with_ancilla else None
        return cls(mode, anc, nxt + (1 if with_ancilla else 0))


# ---------------------------------------------------------------------------------------------
# Pauli sums of the Gray-coded ladder operators  (qevolve.py:107-195)
# ---------------------------------------------------------------------------------------------
def _projector_terms(u: int, v: int, nq: int):
    """Pauli expansion of |u><v| on nq qubits as [(xmask, zmask, coeff)], highest qubit varying slowest and
    (I|X) before (Z|Y) on every qubit -- the enumeration order of the reference's itertools.product."""
    per_qubit = []
    for q in range(nq - 1, -1, -1):
        uq, vq = (u >> q) & 1, (v >> q) & 1
        bit = 1 << q
        if uq == vq:
            per_qubit.append([(0, 0, 0.5), (0, bit, 0.5 if uq == 0 else -0.5)])
        else:
            per_qubit.append([(bit, 0, 0.5), (bit, bit, 0.5j if uq == 0 else -0.5j)])
    out = []
    for combo in itertools.product(*per_qubit):
        x = z = 0
        coeff = 1.0
        first = True
        for cx, cz, cc in combo:
            x |= cx
            z |= cz
            coeff = cc if first else coeff * cc
            first = False
        out.append((x, z, coeff))
    return out


@functools.lru_cache(maxsize=None)
def mode_operator_terms(d_k: int):
    """Ordered {(xmask, zmask): coeff} of a, a^dagger, a + a^dagger and a^dagger a; level l sits on the basis
    state whose bits are Gray(l) = l ^ (l >> 1)."""
    nq = mode_register_size(d_k)
    gray = [l ^ (l >> 1) for l in range(1 << nq)]

    def accumulate(target, u, v, weight):
        for x, z, c in _projector_terms(u, v, nq):
            value = weight * c
            key = (x, z)
            target[key] = target[key] + value if key in target else value

    a, adag, num = {}, {}, {}
    for l in range(1, d_k):
        accumulate(a, gray[l - 1], gray[l], np.sqrt(l))
    for l in range(d_k - 1):
        accumulate(adag, gray[l + 1], gray[l], np.sqrt(l + 1))
    for l in range(d_k):
        accumulate(num, gray[l], gray[l], l)
    a_plus_adag = {k: a[k] + adag[k] for k in a if a[k] != -adag[k]}
    return a, adag, a_plus_adag, num


# ---------------------------------------------------------------------------------------------
# literal term list of one propagation step
# ---------------------------------------------------------------------------------------------
@dataclass
class Rot:
    """exp(-i angle P), P = (Pauli on system qubits xs/zs) x (Pauli on bits xb/zb); angle has shape [B]."""
    xs: int
    zs: int
    xb: int
    zb: int
    angle: np.ndarray


@dataclass
class Dephase:
    site: int
    phi: np.ndarray


@dataclass
class Damp:
    bit: int
    phi: np.ndarray


@dataclass
class Reset:
    bit: int


def product_formula(terms, time, order):
    """[(masks..., coeff[B])] -> [(masks..., angle[B])] in application order.  Order 1 = Lie-Trotter; even orders =
    Suzuki as synthesised by Qiskit 1.0 (order 2: half steps of terms[1:] reversed, full first term, mirrored;
    order 2k by the 1/(4 - 4^(1/(2k-1))) recursion); odd orders > 1 are rejected like Qiskit does."""
    if order == 1:
        return [(m, c * time) for m, c in terms]
    if order % 2 == 1:
        raise ValueError('Suzuki product formulae are symmetric and therefore only defined for even orders.')

    def recurse(o, t):
        if o == 2:
            halves = [(m, c * t / 2) for m, c in reversed(terms[1:])]
            return halves + [(terms[0][0], terms[0][1] * t)] + halves[::-1]
        p = 1 / (4 - 4 ** (1 / (o - 1)))
        outer = recurse(o - 2, p * t)
        return outer + outer + recurse(o - 2, (1 - 4 * p) * t) + outer + outer
    return recurse(order, time)


def _simplify(terms):
    """SparsePauliOp.simplify(): sum duplicates in first-appearance order, drop |c| <= 1e-8."""
    order, sums = [], {}
    for key, c in terms:
        if key not in sums:
            order.append(key)
            sums[key] = 0j
        sums[key] = sums[key] + c
    return [(k, sums[k]) for k in order if not np.isclose(sums[k], 0, atol=1e-8, rtol=1e-5)]


@functools.lru_cache(maxsize=None)
def _coupling_terms(d_k: int):
    """Simplified Pauli sum of (a + a^dagger) (1 - Z_site)/2 for a d_k-level pseudomode: [((x, z, zs), coeff)]."""
    _, _, a_plus_adag, _ = mode_operator_terms(d_k)
    both = [((x, z, zs), c * cz) for (x, z), c in a_plus_adag.items() for zs, cz in ((0, 0.5), (1, -0.5))]
    return tuple(_simplify(both))


def _place(mask: int, positions: list[int]) -> int:
    out = 0
    for j, pos in enumerate(positions):
        if (mask >> j) & 1:
            out |= 1 << pos
    return out


def step_terms(batch: SystemBatch, layout: BitLayout, dt: float, dt_trotter: float, trotter_number: int,
               order: int, coll_rates):
    """One propagation step as a list of Rot / Dephase / Damp / Reset in the reference's order
    (Trotter_number Trotter steps, then one collision round, qevolve.py:419-422).

    ``coll_rates``: None, [B] (excitonic) or [B, W] (chromophore)."""
    n, B = batch.n_sites, batch.batch
    eps, J = batch.energies, batch.couplings

    # system block (qevolve.py:84-101); a term exists if the parameter is non-zero for any member of the batch
    sys_terms = []
    for i in range(n):
        if np.any(eps[:, i] != 0.):
            sys_terms.append(((0, 1 << i, 0, 0), -eps[:, i] / 2))
    for i in range(n):
        for j in range(i + 1, n):
            if np.any(J[:, i, j] != 0.):
                pair = (1 << i) | (1 << j)
                sys_terms.append(((pair, 0, 0, 0), J[:, i, j] / 2))           # X_i X_j
                sys_terms.append(((pair, pair, 0, 0), J[:, i, j] / 2))        # Y_i Y_j
    trotter = [Rot(*m, a) for m, a in product_formula(sys_terms, dt_trotter, order)] if sys_terms else []

    if batch.chromophore:
        for k, d_k in enumerate(batch.levels):                                 # qevolve.py:272-289
            _, _, a_plus_adag, num = mode_operator_terms(d_k)
            omega, g = batch.omega[:, k], batch.coupl_ep[:, k]
            mode_rot, int_rot = [], []
            if np.any(omega != 0):
                terms = [((x, z), omega * np.real(c)) for (x, z), c in num.items()]
                mode_rot = product_formula(terms, dt_trotter, order)
            if np.any(g != 0):
                terms = [(key, g * np.real(c)) for key, c in _coupling_terms(d_k)]
                int_rot = product_formula(terms, dt_trotter, order)
            for i in range(n):
                pos = layout.mode[(i, k)]
                for (x, z), ang in mode_rot:
                    trotter.append(Rot(0, 0, _place(x, pos), _place(z, pos), ang))
                for (x, z, zs), ang in int_rot:
                    trotter.append(Rot(0, (1 << i) if zs else 0, _place(x, pos), _place(z, pos), ang))

    ops = trotter * trotter_number
    if coll_rates is None:
        return ops
    rates = np.asarray(coll_rates, dtype=float)
    if not batch.chromophore:                                                  # qevolve.py:302-304
        phi = np.sqrt(rates.reshape(B)) * np.sqrt(dt)                          # rzx angle / 2
        return ops + [Dephase(i, phi) for i in range(n)]
    rates = rates.reshape(B, batch.n_modes)
    for k, d_k in enumerate(batch.levels):                                     # qevolve.py:328-347
        if not np.any(rates[:, k] != 0):
            continue
        pos0 = layout.mode[(0, k)]
        if len(pos0) == 1:
            # exp(-i phi/2 XX) exp(-i phi/2 YY) on (mode, ancilla |0>) + reset  ==  amplitude damping, phi = sqrt(rate dt)
            phi = dt * np.sqrt(rates[:, k] / dt)
            ops += [Damp(layout.mode[(i, k)][0], phi) for i in range(n)]
            continue
        a, adag, _, _ = mode_operator_terms(d_k)
        # ancilla is the LOW qubit of the reference's gate; here it is its own bit: keys (x_mode, z_mode, x_anc, z_anc)
        both = [((x, z, 1, za), c * ca) for (x, z), c in a.items() for za, ca in ((0, 0.5), (1, -0.5j))]
        both += [((x, z, 1, za), c * ca) for (x, z), c in adag.items() for za, ca in ((0, 0.5), (1, 0.5j))]
        scale = np.sqrt(rates[:, k] / dt)
        terms = [(key, scale * np.real(c)) for key, c in _simplify(both)]
        rots = product_formula(terms, dt, order)
        anc = 1 << layout.ancilla
        for i in range(n):
            pos = layout.mode[(i, k)]
            for (x, z, xa, za), ang in rots:
                ops.append(Rot(0, 0, _place(x, pos) | (anc if xa else 0), _place(z, pos) | (anc if za else 0), ang))
            ops.append(Reset(layout.ancilla))
    return ops


# ---------------------------------------------------------------------------------------------
# lowering to engine operations
# ---------------------------------------------------------------------------------------------
@dataclass
class _Diag:
    zterms: list = field(default_factory=list)      # (zs, zb, angle[B])
    deph: list = field(default_factory=list)        # (site, phi[B])


@dataclass
class _Hop:
    i: int
    j: int
    xx: np.ndarray
    yy: np.ndarray


@dataclass
class _Prot:
    xb: int
    zb: int
    site: int                                        # -1: unconditioned
    plain: np.ndarray                                # summed angle of the terms without Z_site
    signed: np.ndarray                               # summed angle of the terms with Z_site


def _popcount(v: int) -> int:
    return bin(v).count('1')


_POP16 = np.array([bin(v).count('1') for v in range(1 << 16)], dtype=np.int64)


def _popcount_array(values: np.ndarray) -> np.ndarray:
    v = values.astype(np.int64)
    return _POP16[v & 0xffff] + _POP16[(v >> 16) & 0xffff] + _POP16[(v >> 32) & 0xffff]


def _z_commutes(op, zs: int, zb: int) -> bool:
    """Does exp(-i th Z^zs Z^zb) commute with ``op`` (as maps on sigma)?"""
    if isinstance(op, _Hop):
        return _popcount(zs & ((1 << op.i) | (1 << op.j))) % 2 == 0
    if isinstance(op, _Prot):
        return _popcount(zb & op.xb) % 2 == 0
    return True                                      # Damp / Reset commute with Z on their own bit; _Diag trivially


def _deph_commutes(op, site: int) -> bool:
    if isinstance(op, _Hop):
        return site not in (op.i, op.j)
    return True


def lower(terms, n_sites: int):
    """Literal terms -> [_Diag | _Hop | _Prot | Damp | Reset], preserving the result exactly up to rounding:
    Z-type terms move back to the nearest diagonal table they commute with, adjacent XX/YY of one pair fuse into a
    hop, adjacent rotations with the same bit-Pauli fuse into one rotation conditioned on a site's occupation."""
    ops = []

    def add_diagonal(commutes, store):
        j = len(ops)
        while j > 0 and not isinstance(ops[j - 1], _Diag) and commutes(ops[j - 1]):
            j -= 1
        if j > 0 and isinstance(ops[j - 1], _Diag):
            store(ops[j - 1])
        else:
            d = _Diag()
            store(d)
            ops.insert(j, d)

    for t in terms:
        if isinstance(t, Dephase):
            add_diagonal(lambda op: _deph_commutes(op, t.site), lambda d: d.deph.append((t.site, t.phi)))
        elif isinstance(t, (Damp, Reset)):
            ops.append(t)
        elif t.xs == 0 and t.xb == 0:
            if t.zs == 0 and t.zb == 0:
                continue                              # identity term: global phase
            add_diagonal(lambda op: _z_commutes(op, t.zs, t.zb), lambda d: d.zterms.append((t.zs, t.zb, t.angle)))
        elif t.xb == 0:
            if _popcount(t.xs) != 2 or t.zs not in (0, t.xs):
                raise NotImplementedError('system-qubit rotation outside the XX/YY family')
            i, j = [q for q in range(n_sites) if (t.xs >> q) & 1]
            last = ops[-1] if ops else None
            if not (isinstance(last, _Hop) and (last.i, last.j) == (i, j)):
                last = _Hop(i, j, np.zeros_like(t.angle), np.zeros_like(t.angle))
                ops.append(last)
            if t.zs == 0:
                last.xx = last.xx + t.angle
            else:
                last.yy = last.yy + t.angle
        else:
            if t.xs != 0 or _popcount(t.zs) > 1:
                raise NotImplementedError('rotation coupling bits to more than a Z on one site')
            site = [q for q in range(n_sites) if (t.zs >> q) & 1][0] if t.zs else -1
            last = ops[-1] if ops else None
            fusable = (isinstance(last, _Prot) and (last.xb, last.zb) == (t.xb, t.zb)
                       and (site == -1 or last.site in (-1, site)))
            if not fusable:
                last = _Prot(t.xb, t.zb, -1, np.zeros_like(t.angle), np.zeros_like(t.angle))
                ops.append(last)
            if site == -1:
                last.plain = last.plain + t.angle
            else:
                last.site = site
                last.signed = last.signed + t.angle
    for op in ops:
        if isinstance(op, _Hop) and not np.array_equal(op.xx, op.yy):
            raise NotImplementedError('XX and YY angles differ: the step does not conserve the exciton number')
    return ops


# ---------------------------------------------------------------------------------------------
# step program bound to a sector block
# ---------------------------------------------------------------------------------------------
_SIGN_TABLES: dict = {}
_PASS_TABLES: dict = {}


class ParamRecipe:
    """The parameter row of a bound program as functions of the raw angles of the step:

        column value = kind( sum_j coef_j * raw_j )        COS | SIN | NSIN (-sin) | SIN2 (sin squared)
                     = prod_j (cos^2 - sin^2)(raw_j)       PRODCOS (dephasing factors)
                     = 0                                   ZERO (padding that keeps slots 16-byte aligned)

    Columns come in blocks: ``pairs`` adds n entries of (cos, second) column pairs that share a phase -- the (cos, -sin)
    phase tables of the merged Z rotations, the (cos, sin) of a rotation, the (cos, sin^2) of a damping.  Evaluated on
    the host (``evaluate``, NumPy) or on the device (``qsx_bind_params``) from the raw angles [B, R]; both follow the
    same order of operations."""
    ZERO, COS, SIN, NSIN, SIN2, PRODCOS = 0, 1, 2, 3, 4, 5

    def __init__(self, n_sets: int):
        self.n_sets = n_sets
        self.raws, self._raw_id, self._angles = [], {}, None
        self.blocks = []                 # dict(slot, second, idx [n, depth], coef [n, depth], share) | dict(slot, products)
        self.width = 0
        self._shared = {}

    def raw(self, values: np.ndarray) -> int:
        key = id(values)
        if key not in self._raw_id:
            self._raw_id[key] = len(self.raws)
            self.raws.append(np.ascontiguousarray(values, dtype=np.float64).reshape(self.n_sets))
            self._angles = None
        return self._raw_id[key]

    def pairs(self, second: int, idx, coef, share=None) -> int:
        """n entries -> 2n columns (COS, ``second``) of phase_e = sum_t coef[e, t] * raw[idx[e, t]]; returns the slot.
        Blocks with the same ``share`` key have identical content and are evaluated once."""
        idx = np.asarray(idx, dtype=np.int32).reshape(len(coef), -1)
        coef = np.asarray(coef, dtype=np.float64).reshape(idx.shape)
        slot = self.width
        self.blocks.append(dict(slot=slot, second=second, idx=idx, coef=coef, share=share))
        self.width += 2 * idx.shape[0]
        return slot

    def products(self, term_lists) -> int:
        """One PRODCOS column per entry of ``term_lists`` (raw indices), padded to an even width."""
        slot = self.width
        self.blocks.append(dict(slot=slot, products=[list(t) for t in term_lists]))
        self.width += len(term_lists) + len(term_lists) % 2
        return slot

    def angles(self) -> np.ndarray:
        """[B, R] raw angles (R >= 1)."""
        if self._angles is None:
            self._angles = np.stack(self.raws, axis=1) if self.raws else np.zeros((self.n_sets, 1))
        return self._angles

    def tables(self):
        """(kind [S] int32, ptr [S+1] int32, idx [nnz] int32, coef [nnz] float64) for the device evaluation."""
        kinds, counts, idxs, coefs = [], [], [], []
        for blk in self.blocks:
            if 'products' in blk:
                n = len(blk['products'])
                kinds.append(np.array([self.PRODCOS] * n + [self.ZERO] * (n % 2), dtype=np.int32))
                counts.append(np.array([len(t) for t in blk['products']] + [0] * (n % 2), dtype=np.int32))
                idxs.append(np.array([j for t in blk['products'] for j in t], dtype=np.int32))
                coefs.append(np.full(idxs[-1].size, 2.0))
            else:
                n, depth = blk['idx'].shape
                kinds.append(np.tile(np.array([self.COS, blk['second']], dtype=np.int32), n))
                counts.append(np.full(2 * n, depth, dtype=np.int32))
                idxs.append(np.repeat(blk['idx'], 2, axis=0).reshape(-1))
                coefs.append(np.repeat(blk['coef'], 2, axis=0).reshape(-1))
        ptr = np.zeros(self.width + 1, dtype=np.int32)
        if kinds:
            ptr[1:] = np.cumsum(np.concatenate(counts))
            return (np.concatenate(kinds), ptr, np.concatenate(idxs).astype(np.int32), np.concatenate(coefs))
        return np.zeros(0, np.int32), ptr, np.zeros(0, np.int32), np.zeros(0)

    def _block_phase(self, blk) -> np.ndarray:
        """[n, B] phases of a pair block, accumulated term by term (identical term lists give identical bits)."""
        idx, coef = blk['idx'], blk['coef']
        raws = self.angles().T if self.n_sets == 1 else None
        phase = np.zeros((idx.shape[0], self.n_sets))
        for t in range(idx.shape[1]):
            if raws is not None:
                phase += raws[idx[:, t]] * coef[:, t, None]
            else:
                same = idx[0, t]
                if np.all(idx[:, t] == same):                    # one raw angle for the whole block (phase tables)
                    phase += self.raws[same][None, :] * coef[:, t, None]
                else:
                    phase += np.stack([self.raws[j] for j in idx[:, t]]) * coef[:, t, None]
        return phase

    def phases(self, columns) -> np.ndarray:
        """[len(columns), B] phases of linear columns (no trigonometry)."""
        out = np.zeros((len(columns), self.n_sets))
        for r, c in enumerate(columns):
            for blk in self.blocks:
                n_cols = self.blocks_width(blk)
                if blk['slot'] <= c < blk['slot'] + n_cols and 'products' not in blk:
                    e = (c - blk['slot']) // 2
                    for t in range(blk['idx'].shape[1]):
                        out[r] += self.raws[blk['idx'][e, t]] * blk['coef'][e, t]
        return out

    @staticmethod
    def blocks_width(blk) -> int:
        if 'products' in blk:
            return len(blk['products']) + len(blk['products']) % 2
        return 2 * blk['idx'].shape[0]

    def evaluate(self) -> np.ndarray:
        """[B, S] on the host."""
        out = np.zeros((self.n_sets, max(self.width, 0)))
        done = {}
        for blk in self.blocks:
            slot = blk['slot']
            if 'products' in blk:
                for k, term in enumerate(blk['products']):
                    value = np.ones(self.n_sets)
                    for j in term:
                        c, s_ = np.cos(self.raws[j]), np.sin(self.raws[j])
                        value = value * (c * c - s_ * s_)
                    out[:, slot + k] = value
                continue
            n = blk['idx'].shape[0]
            key = blk['share']
            if key is not None and key in done:
                out[:, slot:slot + 2 * n] = out[:, done[key]:done[key] + 2 * n]
                continue
            phase = self._block_phase(blk)
            pair = out[:, slot:slot + 2 * n].reshape(self.n_sets, n, 2)
            pair[:, :, 0] = np.cos(phase).T
            sines = np.sin(phase).T
            if blk['second'] == self.NSIN:
                np.negative(sines, out=sines)
            elif blk['second'] == self.SIN2:
                sines = sines ** 2
            pair[:, :, 1] = sines
            if key is not None:
                done[key] = slot
        return out


class BoundProgram:
    """A step program bound to one (KA, KB) block: the qsx_op array, the parameter rows (evaluated from ``recipe`` on
    first use, or handed in) and, when fused, the pass tables."""

    def __init__(self, block: Block, ops: np.ndarray, params: np.ndarray = None, labels: list = None, passes=None,
                 tiles=None, hops=None, team: int = 0, recipe: ParamRecipe = None):
        self.block = block
        self.ops = ops                   # [n_ops, 8] int32 (qsx_op)
        self._params = params            # [B, stride] float64
        self.labels = labels if labels is not None else []
        self.passes = passes             # [n_passes, 12] int32 (qsx_pass) or None: run the plain operation list
        self.tiles = tiles
        self.hops = hops
        self.team = team                 # threads per instance the pass tables were built for (0: library default)
        self.recipe = recipe
        self.is_transposed = False

    @property
    def params(self) -> np.ndarray:
        if self._params is None:
            self._params = self.recipe.evaluate() if self.recipe.width else np.zeros((self.recipe.n_sets, 2))
        return self._params

    @property
    def n_sets(self) -> int:
        return self._params.shape[0] if self._params is not None else self.recipe.n_sets

    @property
    def stride(self) -> int:
        return self._params.shape[1] if self._params is not None else max(self.recipe.width, 2)

    def transposed(self) -> "BoundProgram":
        """The step that propagates a read-out W instead of the state: <W, Phi sigma> = <Phi^T W, sigma>.  Operations
        in reverse order, each flagged as its transpose (include/qsx.h QSX_OPF_TRANSPOSE); same parameter rows.  The
        pass tables regroup the forward order, so the transposed program runs as a plain operation list."""
        ops = self.ops[::-1].copy()
        ops[:, 7] |= OPF_TRANSPOSE
        out = BoundProgram(self.block, np.ascontiguousarray(ops), self._params, self.labels[::-1], recipe=self.recipe)
        out.is_transposed = True
        out._forward = self
        return out

    def phases(self, slots) -> np.ndarray:
        """[len(slots), B] rotation angles behind the cos/sin columns ``slots`` (from the recipe, no trigonometry)."""
        return self.recipe.phases(list(slots))



class StepProgram:
    """One propagation step of ``qevolve`` for a batch of parameter sets, independent of the sector."""

    def __init__(self, batch: SystemBatch, dt, Trotter_number=1, Trotter_order=1, coll_rates=None,
                 dt_trotter=None):
        self.batch = batch
        self.n_sites = batch.n_sites
        rates = None if coll_rates is None else np.asarray(coll_rates, dtype=float)
        needs_ancilla = (rates is not None and batch.chromophore
                         and any(mode_register_size(d) > 1 and np.any(rates.reshape(batch.batch, -1)[:, k] != 0)
                                 for k, d in enumerate(batch.levels)))
        self.layout = BitLayout.build(batch.n_sites, batch.levels, needs_ancilla)
        self.n_bits = self.layout.n_bits
        if dt_trotter is None:
            dt_trotter = dt / Trotter_number
        self.terms = step_terms(batch, self.layout, dt, dt_trotter, Trotter_number, Trotter_order, rates)
        self.lowered = lower(self.terms, batch.n_sites)
        self._structure_key = None

    def block(self, ka: int, kb: int) -> Block:
        return Block(self.n_sites, self.n_bits, ka, kb)

    def structure_key(self) -> tuple:
        """The lowered step without its parameter values: programs with equal keys share every table that depends on
        the structure only (pass tables, register-resident mappings)."""
        if self._structure_key is None:
            sig = []
            for op in self.lowered:
                if isinstance(op, _Diag):
                    sig.append(('D', tuple((zs, zb) for zs, zb, _ in op.zterms), tuple(site for site, _ in op.deph)))
                elif isinstance(op, _Hop):
                    sig.append(('H', op.i, op.j))
                elif isinstance(op, _Prot):
                    sig.append(('P', op.xb, op.zb, op.site))
                elif isinstance(op, Damp):
                    sig.append(('A', op.bit))
                else:
                    sig.append((type(op).__name__, getattr(op, 'bit', None)))
            modes = tuple((key, tuple(bits)) for key, bits in sorted(self.layout.mode.items()))
            self._structure_key = (self.n_sites, self.n_bits, modes, self.layout.ancilla, tuple(sig))
        return self._structure_key

    def bind(self, block: Block, fused: bool = True, team: int | None = None, kb: int | None = None) -> BoundProgram:
        """Angle tables, qsx_op array and (``fused``) the equivalent pass program for one (KA, KB) block."""
        B = self.batch.batch
        nb = block.n_bits
        ops, labels = [], []
        op_slots = []                                # per lowered op: (index in ops, {name: parameter slot})
        recipe = ParamRecipe(B)
        R = ParamRecipe
        bits_of = np.arange(1 << nb, dtype=np.int64)

        def side_table(cfgs, zterms, key):
            """exp(-i phase) of the merged Z rotations for every (configuration, mode bits) of one side: 2*dim
            columns (cos, -sin); returns the slot."""
            sign_key = (tuple(cfgs), nb, tuple((zs, zb) for zs, zb, _ in zterms))
            coef = _SIGN_TABLES.get(sign_key)                  # +-1 per (entry, Z term): structure only
            if coef is None:
                masks = np.array(cfgs, dtype=np.int64)
                coef = np.empty((len(cfgs) << nb, len(zterms)))
                for t, (zs, zb, _) in enumerate(zterms):
                    par = (_popcount_array(masks & zs)[:, None] + _popcount_array(bits_of & zb)[None, :]) & 1
                    coef[:, t] = (1.0 - 2.0 * par).reshape(-1)
                if len(_SIGN_TABLES) > 512:
                    _SIGN_TABLES.clear()
                _SIGN_TABLES[sign_key] = coef
            idx = np.tile(np.array([recipe.raw(ang) for _, _, ang in zterms], dtype=np.int32), (coef.shape[0], 1))
            return recipe.pairs(R.NSIN, idx, coef, share=(key, tuple(cfgs)))       # ket and bra sides share when KA == KB

        for op in self.lowered:
            if isinstance(op, _Diag):
                slot_a = side_table(block.cfg_a, op.zterms, id(op))
                slot_b = side_table(block.cfg_b, op.zterms, id(op))
                slot_t = -1
                if op.deph:
                    slot_t = recipe.products([[recipe.raw(phi) for site, phi in op.deph if (ca >> site) & 1 != (cb >> site) & 1]
                                              for ca in block.cfg_a for cb in block.cfg_b])
                op_slots.append((len(ops), dict(ga=slot_a, gb=slot_b, t=slot_t)))
                ops.append([OP_DIAG, slot_a, slot_b, slot_t, 0, 0, 0, 0])
                labels.append('diag')
            elif isinstance(op, _Hop):
                slot = recipe.pairs(R.SIN, [[recipe.raw(op.xx)]], [[2.0]])
                op_slots.append((len(ops), dict(cs=slot)))
                ops.append([OP_HOP, op.i, op.j, 0, 0, 0, slot, 0])
                labels.append('hop({},{})'.format(op.i, op.j))
            elif isinstance(op, _Prot):
                skip0 = int(op.site >= 0 and not np.any(op.plain + op.signed != 0))   # Z_site = +1 (empty) / -1 (occupied)
                jp, js = recipe.raw(op.plain), recipe.raw(op.signed)
                slot = recipe.pairs(R.SIN, [[jp, js], [jp, js]], [[1.0, 1.0], [1.0, -1.0]])
                ny = _popcount(op.xb & op.zb) & 3
                op_slots.append((len(ops), dict(cs=slot, skip0=skip0)))
                ops.append([OP_PROT, op.xb, op.zb, ny, op.site, skip0, slot, 0])
                labels.append('prot(x={:b},z={:b},site={})'.format(op.xb, op.zb, op.site))
            elif isinstance(op, Damp):
                slot = recipe.pairs(R.SIN2, [[recipe.raw(op.phi)]], [[1.0]])
                op_slots.append((len(ops), dict(cs=slot)))
                ops.append([OP_AD, op.bit, 0, 0, 0, 0, slot, 0])
                labels.append('damp({})'.format(op.bit))
            elif isinstance(op, Reset):
                op_slots.append((len(ops), {}))
                ops.append([OP_RESET, op.bit, 0, 0, 0, 0, 0, 0])
                labels.append('reset({})'.format(op.bit))
        bound = BoundProgram(block, np.array(ops, dtype=np.int32).reshape(-1, 8), None, labels, recipe=recipe)
        if fused and self.lowered and block.size * 16 <= 220 * 1024:      # larger blocks stream through global memory
            from .passes import build_passes, default_team
            bound.team = team or default_team(block)
            key = (self.structure_key(), block, bound.team, kb, tuple(slots.get('skip0') for _, slots in op_slots))
            tables = _PASS_TABLES.get(key)                     # structure only: shared by programs of equal structure
            if tables is None:
                bit_site = {bit: site for (site, _), bits in self.layout.mode.items() for bit in bits}
                if len(_PASS_TABLES) > 256:
                    _PASS_TABLES.clear()
                tables = _PASS_TABLES[key] = build_passes(self.lowered, op_slots, block, bit_site, bound.team, kb)
            bound.passes, bound.tiles, bound.hops = tables
        return bound


# ---------------------------------------------------------------------------------------------
# time-grid logic  (qevolve.py:373-391)
# ---------------------------------------------------------------------------------------------
def resolve_time_grid(time, dt, Trotter_number):
    """-> (dt, dt_trotter, Trotter_number, integer step count of every entry of ``time``).

    ``dt is None``: dt = (time[1] - time[0]) / Trotter_number (or time / Trotter_number for a scalar) and the
    collisions then fire every dt (Trotter_number becomes 1).  Every time must be a multiple of dt within 1e-10,
    else ``ValueError``."""
    if dt is None:
        dt_trotter = (time if np.isscalar(time) else time[1] - time[0]) / Trotter_number
        dt, Trotter_number = dt_trotter, 1
    else:
        dt_trotter = dt / Trotter_number
    times = np.atleast_1d(np.asarray(time, dtype=float))
    counts = np.round(times / dt)
    if not np.all(np.abs(times - counts * dt) < 1e-10):
        raise ValueError('Entries in time are not multiples of dt_collisions.')
    return dt, dt_trotter, Trotter_number, counts.astype(int)
