"""TEST INFRASTRUCTURE (oracle) -- the reference's gate list evolved inside ONE exciton-number sector.

BASELINE configs[4] (8 sites, one 2-level pseudomode each: 16 qubits + ancilla) cannot be simulated in the full space
on a CPU (the density matrix has 2**32 entries), and neither can the reference itself.  This module keeps the density
matrix on the states with exactly ``k`` excited chromophores (8 x 256 = 2048 states for k = 1) and applies the
LITERAL gate list of ``oracle/circuits.py`` (qevolve.py:74-348) to it:

  * consecutive gates are grouped until the product of the group commutes with the exciton-number operator
    (checked numerically on the small matrix of the qubits the group touches): every Z-type rotation and every
    pseudomode rotation is its own group; X_iX_j and Y_iY_j only conserve the number together (qevolve.py:89-94);
  * a group acts on the sector as the sparse matrix S built from its small unitary; rho -> S rho S^dagger;
  * collisions enter as the channels of ``oracle.fast.kraus_form`` (ancilla traced out), amplitude damping /
    dephasing of one qubit.

Nothing here shares code with the product (which works with fused hops, conditioned rotations and diagonal tables).
The restriction itself is checked against the full-space simulator on small systems in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import circuits as CIR
from .densesim import label_to_masks
from .fast import kraus_form

_PAULI = {'I': np.eye(2, dtype=complex), 'X': np.array([[0, 1], [1, 0]], dtype=complex),
          'Y': np.array([[0, -1j], [1j, 0]], dtype=complex), 'Z': np.array([[1, 0], [0, -1]], dtype=complex)}


def _small_rotation(label_on_qubits: str, theta: float) -> np.ndarray:
    """exp(-i theta P) on the qubits of the label (leftmost character = highest of them)."""
    p = np.ones((1, 1), dtype=complex)
    for ch in label_on_qubits:
        p = np.kron(p, _PAULI[ch])
    return np.cos(theta) * np.eye(p.shape[0]) - 1j * np.sin(theta) * p


class SectorSimulator:
    """Density matrix on the ``k``-exciton states of ``layout`` (no ancilla qubit)."""

    def __init__(self, system, coll_rates, k: int):
        self.layout = CIR.Layout(system, coll_rates is not None)
        lay = self.layout
        self.n_qubits = lay.n_qubits - (1 if lay.anc is not None else 0)
        sys_mask = sum(1 << q for q in lay.sys)
        states = [v for v in range(1 << self.n_qubits) if bin(v & sys_mask).count('1') == k]
        self.states = np.array(states, dtype=np.int64)
        self.index = {v: i for i, v in enumerate(states)}
        self.sys_mask = sys_mask
        self.rho = np.zeros((len(states), len(states)), dtype=complex)
        self._dense = {}

    # ---- step construction ------------------------------------------------------------------------------------
    def compile_step(self, system, dt, dt_trotter, trotter_number, order, coll_rates):
        """-> list of ('unitary', csr S) | ('damp', qubit, phi) | ('dephase', qubit, phi)."""
        lay = self.layout
        ops = kraus_form(CIR.step_ops(system, lay, dt, dt_trotter, trotter_number, order, coll_rates), lay)
        if ops is None:
            raise NotImplementedError('collisions outside the two-gate patterns')
        out, pending = [], []                          # pending: [(qubits (ascending), small unitary on them)]

        def flush_if_conserving():
            qubits = sorted({q for qs, _ in pending for q in qs})
            total = np.eye(1 << len(qubits), dtype=complex)
            for qs, u in pending:                      # embed every pending gate on the union of qubits, in order
                total = self._embed(u, qs, qubits) @ total
            number = np.array([bin(sum(((v >> j) & 1) << q for j, q in enumerate(qubits)) & self.sys_mask).count('1')
                               for v in range(1 << len(qubits))])
            off = total[number[:, None] != number[None, :]]
            mixing = np.abs(off).max() if off.size else 0.0
            if mixing < 1e-14:
                out.append(('unitary', self._sector_matrix(total, qubits)))
                pending.clear()
                return True
            return False

        for op in ops:
            if op[0] == 'rot':
                x, z, _ = label_to_masks(op[1])
                qubits = [q for q in range(self.n_qubits) if ((x | z) >> q) & 1]
                if not qubits:
                    continue                           # identity term: global phase
                label = ''.join(op[1][self.n_qubits - 1 - q] for q in reversed(qubits))
                pending.append((qubits, _small_rotation(label, op[2])))
                flush_if_conserving()
            else:
                if pending:
                    raise ValueError('a channel follows gates that do not conserve the exciton number')
                out.append(op)
        if pending:
            raise ValueError('the step ends inside a group that does not conserve the exciton number')
        # consecutive unitary groups multiply into one sparse matrix (built once per step program, applied every step)
        merged = []
        for op in out:
            if op[0] == 'unitary' and merged and merged[-1][0] == 'unitary':
                merged[-1] = ('unitary', (op[1] @ merged[-1][1]).tocsr())
            else:
                merged.append(op)
        return merged

    @staticmethod
    def _embed(u, qs, qubits):
        """Small unitary on qubits ``qs`` -> on the sorted superset ``qubits`` (little-endian bit order)."""
        pos = [qubits.index(q) for q in qs]
        dim = 1 << len(qubits)
        out = np.zeros((dim, dim), dtype=complex)
        for v in range(dim):
            sub_v = sum(((v >> p) & 1) << j for j, p in enumerate(pos))
            rest = v & ~sum(1 << p for p in pos)
            for sub_w in range(1 << len(qs)):
                w = rest | sum(((sub_w >> j) & 1) << p for j, p in enumerate(pos))
                out[w, v] = u[sub_w, sub_v]
        return out

    def _sector_matrix(self, total, qubits):
        rows, cols, vals = [], [], []
        qmask = sum(1 << q for q in qubits)
        for i, v in enumerate(self.states):
            sub_v = sum(((int(v) >> q) & 1) << j for j, q in enumerate(qubits))
            for sub_w in range(1 << len(qubits)):
                amp = total[sub_w, sub_v]
                if amp == 0:
                    continue
                w = (int(v) & ~qmask) | sum(((sub_w >> j) & 1) << q for j, q in enumerate(qubits))
                j = self.index.get(w)
                if j is None:
                    assert abs(amp) < 1e-14
                    continue
                rows.append(j), cols.append(i), vals.append(amp)
        n = len(self.states)
        return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))

    # ---- evolution ----------------------------------------------------------------------------------------------
    def set_pure(self, amplitudes: dict):
        """{full basis state: amplitude} -> rho = |psi><psi|."""
        psi = np.zeros(len(self.states), dtype=complex)
        for v, a in amplitudes.items():
            psi[self.index[v]] = a
        self.rho = np.outer(psi, psi.conj())

    def apply_step(self, compiled):
        for op in compiled:
            if op[0] == 'unitary':
                s = op[1]
                if s.nnz > 32 * s.shape[0]:             # the merged step is fairly dense: two BLAS products
                    u = self._dense.setdefault(id(s), s.toarray())
                    self.rho = u @ self.rho @ u.conj().T
                else:
                    half = s @ self.rho                 # S rho, then (S^* (S rho)^T)^T = S rho S^dagger
                    self.rho = np.ascontiguousarray((s.conj() @ np.ascontiguousarray(half.T)).T)
            else:
                bit = (self.states >> op[1]) & 1
                c, sn = np.cos(op[2]), np.sin(op[2])
                if op[0] == 'dephase':
                    self.rho *= np.where(bit[:, None] != bit[None, :], c * c - sn * sn, 1.0)
                else:                                   # amplitude damping: K0 = diag(1, cos), K1 = -i sin |0><1|
                    ones = np.nonzero(bit)[0]
                    lower = np.array([self.index[int(v) & ~(1 << op[1])] for v in self.states[ones]])
                    moved = self.rho[np.ix_(ones, ones)]                  # (a copy) the |1><1| part of the qubit
                    factor = np.where(bit == 1, c, 1.0)
                    self.rho *= factor[:, None]
                    self.rho *= factor[None, :]
                    self.rho[np.ix_(lower, lower)] += sn * sn * moved     # `lower` has no duplicates

    def populations(self):
        diag = np.real(np.diag(self.rho))
        return np.array([diag[((self.states >> q) & 1) == 1].sum() for q in self.layout.sys])


def populations(system, steps, dt, coll_rates=None, Trotter_number=1, Trotter_order=1, k=1):
    """Site populations after each step count in ``steps`` (ascending) for a system prepared in a state of the
    ``k``-exciton sector ('localized excitation' / 'delocalized excitation')."""
    rates = CIR.normalise_coll_rates(system, coll_rates)
    dt_, dt_trotter, tn, _ = CIR.resolve_time_grid(0.0, dt, Trotter_number)
    sim = SectorSimulator(system, rates, k)
    compiled = sim.compile_step(system, dt_, dt_trotter, tn, Trotter_order, rates)
    lay = sim.layout
    if system.state_type == 'localized excitation':
        sim.set_pure({1 << lay.sys[system.state]: 1.0})
    elif system.state_type == 'delocalized excitation':
        sim.set_pure({1 << lay.sys[i]: c for i, c in enumerate(system.state) if c != 0})
    else:
        raise ValueError('sector oracle needs a one-exciton initial state')
    out, done = [], 0
    for target in steps:
        for _ in range(int(target) - done):
            sim.apply_step(compiled)
        done = int(target)
        out.append(sim.populations())
    return np.array(out)
