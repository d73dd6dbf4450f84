"""TEST INFRASTRUCTURE (oracle) -- dense density-matrix simulator of small qubit circuits.

Exact-expectation semantics of the gate set the reference emits through Qiskit
(qevolve.py:51-64, :100-105, :215-218, :244-253, :303-304, :334-348; qspectroscopy.py:73-79, :87):
Pauli rotations exp(-i theta P), rzx, h/x/cx/cy, initialize, reset (trace out + |0>), and the two
read-outs the reference uses -- measurement probabilities of a register and <P> of a Pauli string.

Conventions (Qiskit): qubit q is bit q of the basis-state index (little-endian); a Pauli label is
written with the highest qubit leftmost.  P = i^{#Y} X^x Z^z, i.e. P|v> = i^{#Y} (-1)^{|v&z|} |v^x>.

Nothing in the product imports this file; only tests/, bench.py's cpu_baseline/--impl reference
legs, __graft_entry__.smoke() and tests/golden/make_golden.py do.
"""
from __future__ import annotations

import numpy as np

_PARITY_CACHE: dict[int, np.ndarray] = {}


def _parity_table(n: int) -> np.ndarray:
    """popcount parity of every index < 2**n."""
    tab = _PARITY_CACHE.get(n)
    if tab is None:
        idx = np.arange(1 << n, dtype=np.int64)
        par = np.zeros_like(idx)
        for q in range(n):
            par ^= (idx >> q) & 1
        _PARITY_CACHE[n] = tab = par
    return tab


def label_to_masks(label: str) -> tuple[int, int, int]:
    """'XIZY' -> (xmask, zmask, number of Y); leftmost character = highest qubit."""
    n = len(label)
    x = z = ny = 0
    for pos, ch in enumerate(label):
        q = n - 1 - pos
        if ch == 'X':
            x |= 1 << q
        elif ch == 'Z':
            z |= 1 << q
        elif ch == 'Y':
            x |= 1 << q
            z |= 1 << q
            ny += 1
        elif ch != 'I':
            raise ValueError('bad Pauli label {!r}'.format(label))
    return x, z, ny


def place_label(label: str, qubits: list[int], n: int) -> str:
    """Embed a label acting on ``qubits`` (label's qubit j -> qubits[j]) into an n-qubit label."""
    chars = ['I'] * n
    k = len(label)
    for j in range(k):
        chars[n - 1 - qubits[j]] = label[k - 1 - j]
    return ''.join(chars)


class DensityMatrix:
    """n-qubit density operator (or any operator: linear maps only), all qubits start in |0>."""

    def __init__(self, n: int):
        self.n = n
        dim = 1 << n
        self.rho = np.zeros((dim, dim), dtype=complex)
        self.rho[0, 0] = 1.0
        self._idx = np.arange(dim, dtype=np.int64)

    def copy(self) -> "DensityMatrix":
        out = DensityMatrix.__new__(DensityMatrix)
        out.n, out.rho, out._idx = self.n, self.rho.copy(), self._idx
        return out

    # ---- Pauli machinery -----------------------------------------------------------
    def _left_right_phases(self, x: int, z: int, ny: int):
        par = _parity_table(self.n)
        unit = 1j ** (ny % 4)
        left = unit * (1 - 2 * par[(self._idx ^ x) & z])      # (P rho)[a,:] = left[a] rho[a^x,:]
        right = unit * (1 - 2 * par[self._idx & z])            # (rho P)[:,b] = rho[:,b^x] right[b]
        return left, right

    def pauli_rotation(self, label: str, theta: float):
        """rho -> G rho G^dagger with G = exp(-i theta P) = cos(theta) I - i sin(theta) P."""
        x, z, ny = label_to_masks(label)
        if x == 0 and z == 0:
            return                                             # global phase
        c, s = np.cos(theta), np.sin(theta)
        left, right = self._left_right_phases(x, z, ny)
        perm = self._idx ^ x
        rho = self.rho
        p_rho = left[:, None] * rho[perm, :]
        rho_p = rho[:, perm] * right[None, :]
        p_rho_p = p_rho[:, perm] * right[None, :]
        self.rho = c * c * rho + s * s * p_rho_p - 1j * s * c * (p_rho - rho_p)

    def expectation_pauli(self, label: str) -> complex:
        """Tr[P rho]."""
        x, z, ny = label_to_masks(label)
        left, _ = self._left_right_phases(x, z, ny)
        return complex(np.sum(left * self.rho[self._idx ^ x, self._idx]))

    # ---- generic small unitaries ---------------------------------------------------
    def apply_matrix(self, mat: np.ndarray, qubits: list[int]):
        """rho -> U rho U^dagger for a 2^k x 2^k matrix on ``qubits`` (matrix bit j <-> qubits[j])."""
        n, k = self.n, len(qubits)
        mat = np.asarray(mat, dtype=complex).reshape([2] * (2 * k))
        # matrix axes: (out bits k-1..0, in bits k-1..0)
        t = self.rho.reshape([2] * (2 * n))
        row_axes = [n - 1 - q for q in reversed(qubits)]             # highest matrix bit first
        col_axes = [2 * n - 1 - q for q in reversed(qubits)]
        t = np.tensordot(mat, t, axes=(list(range(k, 2 * k)), row_axes))
        t = np.moveaxis(t, list(range(k)), row_axes)
        t = np.tensordot(mat.conj(), t, axes=(list(range(k, 2 * k)), col_axes))
        t = np.moveaxis(t, list(range(k)), col_axes)
        self.rho = np.ascontiguousarray(t).reshape(1 << n, 1 << n)

    def h(self, q):
        self.apply_matrix(np.array([[1, 1], [1, -1]]) / np.sqrt(2), [q])

    def x(self, q):
        self.apply_matrix(np.array([[0, 1], [1, 0]]), [q])

    def cx(self, control, target):
        m = np.eye(4, dtype=complex)        # bit0 = control, bit1 = target
        m[[1, 3]] = m[[3, 1]]
        self.apply_matrix(m, [control, target])

    def cy(self, control, target):
        m = np.zeros((4, 4), dtype=complex)
        m[0, 0] = m[2, 2] = 1
        m[3, 1] = 1j                          # |c=1,t=0> -> i |c=1,t=1>
        m[1, 3] = -1j                         # |c=1,t=1> -> -i |c=1,t=0>
        self.apply_matrix(m, [control, target])

    def rzx(self, theta, q0, q1):
        """Qiskit RZXGate: exp(-i theta/2 Z_{q0} X_{q1})."""
        chars = ['I'] * self.n
        chars[self.n - 1 - q0] = 'Z'
        chars[self.n - 1 - q1] = 'X'
        self.pauli_rotation(''.join(chars), theta / 2)

    def reset(self, q):
        """Trace out qubit q and put it back in |0>."""
        n = self.n
        t = self.rho.reshape([2] * (2 * n))
        ra, ca = n - 1 - q, 2 * n - 1 - q
        reduced = np.trace(t, axis1=ra, axis2=ca)                   # drops both axes
        out = np.zeros_like(t)
        sl = [slice(None)] * (2 * n)
        sl[ra] = 0
        sl[ca] = 0
        out[tuple(sl)] = reduced
        self.rho = out.reshape(1 << n, 1 << n)

    def initialize(self, amplitudes, qubits: list[int]):
        """Qiskit ``initialize``: reset the qubits, then prepare ``amplitudes`` on them
        (amplitude index bit j <-> qubits[j])."""
        for q in qubits:
            self.reset(q)
        amp = np.asarray(amplitudes, dtype=complex).reshape(-1)
        dim = amp.size
        basis = np.eye(dim, dtype=complex)
        basis[:, 0] = amp
        u, r = np.linalg.qr(basis)
        u[:, 0] *= r[0, 0] / abs(r[0, 0])                           # first column == amp exactly up to rounding
        self.apply_matrix(u, qubits)

    # ---- read-outs -------------------------------------------------------------------
    def probabilities(self, qubits: list[int]) -> np.ndarray:
        """Distribution of a computational-basis measurement of ``qubits``;
        outcome index bit j <-> qubits[j]."""
        diag = np.real(np.diag(self.rho))
        out = np.zeros(1 << len(qubits))
        key = np.zeros_like(self._idx)
        for j, q in enumerate(qubits):
            key |= ((self._idx >> q) & 1) << j
        np.add.at(out, key, diag)
        return out

    def trace(self) -> complex:
        return complex(np.trace(self.rho))


def pauli_matrix(label: str) -> np.ndarray:
    """Dense matrix of a Pauli label by Kronecker products (independent of the mask code above;
    used by tests to cross-check ``pauli_rotation``)."""
    single = {'I': np.eye(2), 'X': np.array([[0, 1], [1, 0]]),
              'Y': np.array([[0, -1j], [1j, 0]]), 'Z': np.diag([1, -1])}
    out = np.ones((1, 1), dtype=complex)
    for ch in label:
        out = np.kron(out, single[ch])
    return out
