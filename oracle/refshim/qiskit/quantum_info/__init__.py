"""SparsePauliOp stand-in: ordered list of (label, coeff).  ``a ^ b`` = tensor with ``b`` on the LOW
qubits and a-major term order; ``simplify`` sums duplicates in first-appearance order and drops
|c| <= 1e-8 (qiskit 1.0 `SparsePauliOp.simplify`, atol=1e-8)."""
from __future__ import annotations

import numpy as np


class SparsePauliOp:
    __array_ufunc__ = None      # make numpy scalars defer to __rmul__

    def __init__(self, data, coeffs=None):
        if isinstance(data, SparsePauliOp):
            labels, cs = list(data.labels), np.array(data.coeffs, dtype=complex)
        else:
            labels = [data] if isinstance(data, str) else list(data)
            cs = np.ones(len(labels), dtype=complex) if coeffs is None else np.array(coeffs, dtype=complex).reshape(-1)
        if len(labels) != len(cs) or not labels:
            raise ValueError('bad SparsePauliOp')
        self.labels, self.coeffs = labels, cs

    @property
    def num_qubits(self):
        return len(self.labels[0])

    def to_list(self):
        return [(l, complex(c)) for l, c in zip(self.labels, self.coeffs)]

    def tensor(self, other):
        labels = [a + b for a in self.labels for b in other.labels]
        return SparsePauliOp(labels, np.kron(self.coeffs, other.coeffs))

    __xor__ = tensor

    def __add__(self, other):
        return SparsePauliOp(self.labels + other.labels, np.concatenate([self.coeffs, other.coeffs]))

    def __neg__(self):
        return SparsePauliOp(self.labels, -self.coeffs)

    def __sub__(self, other):
        return self + (-other)

    def __mul__(self, k):
        return SparsePauliOp(self.labels, self.coeffs * complex(k))

    __rmul__ = __mul__

    def simplify(self, atol=1e-8, rtol=1e-5):
        order, sums = [], {}
        for l, c in zip(self.labels, self.coeffs):
            if l not in sums:
                order.append(l)
                sums[l] = 0j
            sums[l] += c
        kept = [l for l in order if not np.isclose(sums[l], 0, atol=atol, rtol=rtol)]
        if not kept:
            return SparsePauliOp(['I' * self.num_qubits], [0])
        return SparsePauliOp(kept, [sums[l] for l in kept])
