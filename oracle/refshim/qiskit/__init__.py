"""Minimal stand-in for the parts of qiskit==1.0.0 the reference uses (see ../README.md)."""
from __future__ import annotations

import numpy as np


class _Bit:
    __slots__ = ('register', 'index')

    def __init__(self, register, index):
        self.register, self.index = register, index


class _Register:
    _counter = 0

    def __init__(self, size, name=None):
        self.size = int(size)
        if name is None:
            _Register._counter += 1
            name = 'r{}'.format(_Register._counter)
        self.name = name
        self._bits = [_Bit(self, i) for i in range(self.size)]

    def __len__(self):
        return self.size

    def __getitem__(self, i):
        return self._bits[i]

    def __iter__(self):
        return iter(self._bits)


class QuantumRegister(_Register):
    pass


class AncillaRegister(QuantumRegister):
    pass


class ClassicalRegister(_Register):
    pass


class QuantumCircuit:
    """Records instructions as (name, params, qubit indices[, clbit indices])."""

    def __init__(self, *regs):
        self.qubits: list = []
        self.clbits: list = []
        self._qindex: dict = {}
        self._cindex: dict = {}
        self.data: list = []
        for r in regs:
            if isinstance(r, (int, np.integer)):
                r = QuantumRegister(int(r))
            self.add_register(r)

    @property
    def num_qubits(self):
        return len(self.qubits)

    def add_register(self, reg):
        if isinstance(reg, ClassicalRegister):
            for b in reg:
                self._cindex[b] = len(self.clbits)
                self.clbits.append(b)
        else:
            for b in reg:
                self._qindex[b] = len(self.qubits)
                self.qubits.append(b)

    def _q(self, spec) -> list[int]:
        if isinstance(spec, (int, np.integer)):
            return [int(spec)]
        if isinstance(spec, _Bit):
            return [self._qindex[spec]]
        out = []
        for item in spec:
            out += self._q(item)
        return out

    def copy(self):
        new = QuantumCircuit()
        new.qubits, new.clbits = list(self.qubits), list(self.clbits)
        new._qindex, new._cindex = dict(self._qindex), dict(self._cindex)
        new.data = list(self.data)
        return new

    def decompose(self, reps=1):
        return self

    def _add(self, name, params, qubits, clbits=()):
        self.data.append((name, params, tuple(qubits), tuple(clbits)))

    def x(self, q):
        self._add('x', (), self._q(q))

    def h(self, q):
        self._add('h', (), self._q(q))

    def cx(self, c, t):
        self._add('cx', (), self._q(c) + self._q(t))

    def cy(self, c, t):
        self._add('cy', (), self._q(c) + self._q(t))

    def rzx(self, theta, q0, q1):
        self._add('rzx', (float(theta),), self._q(q0) + self._q(q1))

    def reset(self, q):
        for qi in self._q(q):
            self._add('reset', (), [qi])

    def initialize(self, state, qubits=None):
        qs = list(range(self.num_qubits)) if qubits is None else self._q(qubits)
        self._add('initialize', (tuple(np.asarray(state, dtype=complex).tolist()),), qs)

    def measure(self, qubits, clbits):
        qs = self._q(qubits)
        cs = [self._cindex[b] for b in clbits] if not isinstance(clbits, _Bit) else [self._cindex[clbits]]
        self._add('measure', (), qs, cs)

    def append(self, gate, qargs):
        qs = self._q(qargs)
        for name, params, local in gate.definition_ops():
            self._add(name, params, [qs[j] for j in local])

    def compose(self, other, qubits=None, inplace=False):
        target = self if inplace else self.copy()
        mapping = list(range(other.num_qubits)) if qubits is None else target._q(qubits)
        if len(mapping) != other.num_qubits:
            raise ValueError('compose: qubit count mismatch')
        for name, params, qs, cs in other.data:
            target.data.append((name, params, tuple(mapping[j] for j in qs), cs))
        return None if inplace else target
