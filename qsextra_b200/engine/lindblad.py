"""Lindblad limit of the collision model (the reference's comparison path: ``clevolve.py:16-108`` builds H, the
collapse operators and the population operators and hands them to QuTiP ``mesolve``/``sesolve``).

Here the same master equation

    d sigma/dt = -i [H, sigma] + sum_c ( C sigma C^+ - 1/2 {C^+ C, sigma} )

is solved on the GPU in the exciton-number sector layout of the engine (``sectors.py``): H and every collapse operator of
the reference conserve the exciton number, so a (KA, KB) block of sigma evolves by itself under the sparse generator

    L = -i (H_A (x) 1 - 1 (x) H_B^T) + sum_c [ C_A (x) conj(C_B) - 1/2 (C^+C)_A (x) 1 - 1/2 1 (x) (C^+C)_B^T ]

(row-major vec, index a * DB + b).  ``qsx_lindblad_evolve`` propagates with a scaled Taylor series of exp(h L) -- a
batched sparse matrix-vector product per term (ELL storage, coalesced) -- which is exact to rounding for the step sizes
chosen here, so the result does not depend on an ODE tolerance.

Operators (reference file:line):
    H_e   = sum_i -e_i/2 Z_i + sum_{i<j} J_ij (s-_j s+_i + h.c.)                        system.py:260-276
    H_p   = sum_{i,k} w_k a+_{ik} a_{ik};  H_ep = sum_{i,k} g_k/2 (1 - Z_i)(a_{ik} + a+_{ik})   system.py:475-527
    C     = sqrt(rate) Z_i (ExcitonicSystem)  |  sqrt(rate_k) a_{ik} (ChromophoreSystem)   clevolve.py:42-44, :62-77
Pseudomode (i, k) with d_k levels occupies ceil(log2 d_k) bits (binary-coded level) of the ``bits`` part of an index;
codes >= d_k are never populated (their rows and columns of L are empty).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from .sectors import Block, sector_configs, sector_rank

TAYLOR_TERMS = 20          # 1 / 20! = 4e-19: the series of exp(hL) with ||hL||_1 <= THETA is exact to rounding
THETA = 1.0


@dataclass
class BoundLindblad:
    """ELL form of the generator of one block, ready for upload."""
    block: Block
    width: int
    ell_idx: np.ndarray      # [width, E] int32 (padding: own row, value 0)
    ell_val: np.ndarray      # [width, E] complex128
    norm1: float


class LindbladProgram:
    """Generator of the master equation of one system, per sector block.  Duck-types the part of ``StepProgram`` that
    ``engine.response.response_function`` uses (``n_sites``, ``n_bits``, ``block``, ``bind``) with ``continuous`` set."""
    continuous = True

    def __init__(self, energies, couplings, rates=None, omega=None, coupl_ep=None, levels=()):
        self.energies = np.asarray(energies, dtype=float)
        self.n_sites = n = len(self.energies)
        self.couplings = np.asarray(couplings, dtype=float).reshape(n, n)
        self.levels = tuple(int(d) for d in levels)
        self.omega = None if omega is None else np.asarray(omega, dtype=float)
        self.coupl_ep = None if coupl_ep is None else np.asarray(coupl_ep, dtype=float)
        self.chromophore = len(self.levels) > 0
        if rates is None:
            self.rates = None
        elif self.chromophore:
            self.rates = np.atleast_1d(np.asarray(rates, dtype=float))
        else:
            self.rates = float(rates)
        self.reg_bits = tuple(max(1, int(np.ceil(np.log2(d)))) for d in self.levels)
        per_site = sum(self.reg_bits)
        self.n_bits = n * per_site
        if self.n_bits > 24:
            raise ValueError('pseudomode registers need {} bits; at most 24 are supported'.format(self.n_bits))
        self.offset = {}                                     # (site, mode) -> (bit offset, width)
        cursor = 0
        for i in range(n):
            for k, q in enumerate(self.reg_bits):
                self.offset[i, k] = (cursor, q)
                cursor += q
        self._sector = {}

    @classmethod
    def from_system(cls, system, rates=None):
        md = getattr(system, 'mode_dict', None)
        if md is None:
            return cls(system.e_el, system.coupl_el, rates)
        return cls(system.e_el, system.coupl_el, rates, md['omega_mode'], md['coupl_ep'], md['lvl_mode'])

    def block(self, ka: int, kb: int) -> Block:
        return Block(self.n_sites, self.n_bits, ka, kb)

    # ---- basis bookkeeping ------------------------------------------------------------------------
    def level_table(self):
        """levels[(site, mode)] of every ``bits`` value and the mask of valid ones."""
        bits = np.arange(1 << self.n_bits, dtype=np.int64)
        lv = {key: (bits >> off) & ((1 << q) - 1) for key, (off, q) in self.offset.items()}
        valid = np.ones(bits.size, dtype=bool)
        for (i, k), v in lv.items():
            valid &= v < self.levels[k]
        return lv, valid

    def full_space_index(self, k_exc: int) -> np.ndarray:
        """Position of every basis state of sector ``k_exc`` in the reference's full tensor-product space
        [e_{N-1}..e_0, modes(N-1)..modes(0)] (clevolve.py:58-60, :83-86); -1 for padded level codes."""
        lv, valid = self.level_table()
        mode_index = np.zeros(1 << self.n_bits, dtype=np.int64)
        for i in range(self.n_sites - 1, -1, -1):
            for k, d in enumerate(self.levels):
                mode_index = mode_index * d + np.where(valid, lv[i, k], 0)
        m_total = int(np.prod([d for d in self.levels], dtype=np.int64)) ** self.n_sites if self.levels else 1
        cfgs = np.array(sector_configs(self.n_sites, k_exc), dtype=np.int64)
        full = cfgs[:, None] * m_total + mode_index[None, :]
        full[:, ~valid] = -1
        return full.reshape(-1)

    @property
    def full_dimension(self) -> int:
        return (1 << self.n_sites) * (int(np.prod(self.levels, dtype=np.int64)) ** self.n_sites if self.levels else 1)

    # ---- operators of one exciton-number sector ---------------------------------------------------
    def sector_operators(self, k_exc: int):
        """-> (H, [C ...]) as CSR matrices on the basis ``cfg_index << n_bits | bits`` of sector ``k_exc``."""
        if k_exc in self._sector:
            return self._sector[k_exc]
        n, nb = self.n_sites, self.n_bits
        cfgs = np.array(sector_configs(n, k_exc), dtype=np.int64)
        rank = sector_rank(n, k_exc)
        dim = len(cfgs) << nb
        idx = np.arange(dim, dtype=np.int64)
        p, bits = idx >> nb, idx & ((1 << nb) - 1)
        mask = cfgs[p]
        lv, valid_bits = self.level_table()
        valid = valid_bits[bits]
        occ = [(mask >> i) & 1 for i in range(n)]

        diag = np.zeros(dim)
        for i in range(n):
            diag += -self.energies[i] / 2 * (1 - 2 * occ[i])
        rows, cols, vals = [idx[valid]], [idx[valid]], []
        if self.chromophore:
            for i in range(n):
                for k, w in enumerate(self.omega):
                    diag += w * lv[i, k][bits]
        vals.append(diag[valid])
        for i in range(n):
            for j in range(i + 1, n):
                if self.couplings[i, j] == 0:
                    continue
                move = valid & (occ[i] != occ[j])
                src = idx[move]
                dst = (rank[mask[move] ^ ((1 << i) | (1 << j))].astype(np.int64) << nb) | bits[move]
                rows.append(dst), cols.append(src), vals.append(np.full(src.size, self.couplings[i, j]))
        lowering = {}                                      # (site, mode) -> CSR of a_{ik} on this sector
        if self.chromophore:
            for i in range(n):
                for k, g in enumerate(self.coupl_ep):
                    off, _ = self.offset[i, k]
                    level = lv[i, k][bits]
                    up = valid & (level + 1 < self.levels[k])
                    src, dst = idx[up], idx[up] + (1 << off)
                    amp = np.sqrt(level[up] + 1.0)
                    lowering[i, k] = sp.csr_matrix((amp, (src, dst)), shape=(dim, dim))      # a |l+1> = sqrt(l+1) |l>
                    coupled = occ[i][up] == 1
                    for r, c in ((dst, src), (src, dst)):
                        rows.append(r[coupled]), cols.append(c[coupled]), vals.append(g * amp[coupled])
        ham = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(dim, dim))
        collapse = []
        if self.rates is not None:
            if self.chromophore:
                for i in range(n):
                    for k, rate in enumerate(self.rates):
                        collapse.append(np.sqrt(rate) * lowering[i, k])
            else:
                for i in range(n):
                    collapse.append(sp.diags(np.sqrt(self.rates) * (1.0 - 2 * occ[i]) * valid, format='csr'))
        self._sector[k_exc] = (ham, collapse)
        return ham, collapse

    def generator(self, block: Block) -> sp.csr_matrix:
        """Sparse L of one (KA, KB) block acting on row-major vec(sigma)."""
        ham_a, col_a = self.sector_operators(block.ka)
        ham_b, col_b = self.sector_operators(block.kb)
        eye_a, eye_b = sp.identity(block.dim_a, format='csr'), sp.identity(block.dim_b, format='csr')
        gen = -1j * (sp.kron(ham_a, eye_b, format='csr') - sp.kron(eye_a, ham_b.T, format='csr'))
        for ca, cb in zip(col_a, col_b):
            na, nb_ = (ca.conj().T @ ca), (cb.conj().T @ cb)
            gen = gen + sp.kron(ca, cb.conj(), format='csr') - 0.5 * sp.kron(na, eye_b, format='csr') \
                - 0.5 * sp.kron(eye_a, nb_.T, format='csr')
        gen = sp.csr_matrix(gen, dtype=complex)
        gen.sum_duplicates()
        gen.eliminate_zeros()
        return gen

    def bind(self, block: Block, transposed: bool = False, **_ignored) -> BoundLindblad:
        """ELL tables of the generator; ``transposed``: of L^T, which propagates a read-out W instead of the state
        (<W, e^{Lt} sigma> = <e^{L^T t} W, sigma>, pairing without conjugation)."""
        gen = self.generator(block)
        return _to_ell(gen.T if transposed else gen, block)

    def bind_ket(self, k_exc: int) -> BoundLindblad:
        """Schroedinger generator -iH of one exciton-number sector (a ket evolves as a one-column block)."""
        ham, _ = self.sector_operators(k_exc)
        return _to_ell(sp.csr_matrix(-1j * ham), _KetBlock(self.n_sites, self.n_bits, k_exc, ham.shape[0]))

    # ---- states ------------------------------------------------------------------------------------
    def mode_amplitudes(self, mode_kets=None) -> np.ndarray:
        """Amplitude of every ``bits`` value for the product state of the pseudomodes (vacuum by default)."""
        lv, valid = self.level_table()
        amp = valid.astype(complex)
        for (i, k), level in lv.items():
            ket = np.zeros(self.levels[k], dtype=complex)
            if mode_kets is None:
                ket[0] = 1.0
            else:
                ket[:] = np.asarray(mode_kets[k], dtype=complex).reshape(-1)
            amp = amp * np.where(valid, ket[np.minimum(level, self.levels[k] - 1)], 0)
        return amp


@dataclass(frozen=True)
class _KetBlock:
    """Stand-in for ``Block`` when a ket (not an operator block) is propagated: only ``size`` matters."""
    n_sites: int
    n_bits: int
    ka: int
    size: int


def _to_ell(gen: sp.csr_matrix, block) -> BoundLindblad:
    gen = sp.csr_matrix(gen, dtype=complex)
    gen.sum_duplicates()
    gen.eliminate_zeros()
    gen.sort_indices()
    E = gen.shape[0]
    counts = np.diff(gen.indptr)
    width = max(1, int(counts.max()) if E else 1)
    ell_idx = np.tile(np.arange(E, dtype=np.int32), (width, 1))
    ell_val = np.zeros((width, E), dtype=complex)
    row_of = np.repeat(np.arange(E), counts)
    slot = np.arange(gen.nnz) - gen.indptr[row_of]
    ell_idx[slot, row_of] = gen.indices
    ell_val[slot, row_of] = gen.data
    norm1 = float(abs(gen).sum(axis=0).max()) if gen.nnz else 0.0
    return BoundLindblad(block, width, np.ascontiguousarray(ell_idx), np.ascontiguousarray(ell_val), norm1)


def segments_from_times(times) -> np.ndarray:
    """Durations between consecutive emissions for emission times measured from the start of the propagation."""
    t = np.asarray(times, dtype=float)
    return np.diff(np.concatenate([[0.0], t]))
