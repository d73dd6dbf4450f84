"""Minimal stand-in for the parts of QuTiP the reference uses (see ../README.md).

Dense ``numpy`` operators with QuTiP's tensor ordering; ``mesolve``/``sesolve`` are EXACT
(matrix exponential of the Lindblad generator between requested times) instead of an adaptive ODE.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm


class Qobj:
    __array_ufunc__ = None

    def __init__(self, data, dims=None, type=None):
        arr = np.array(data.data if isinstance(data, Qobj) else data, dtype=complex)
        if arr.ndim == 1:
            arr = arr.reshape(-1, 1)
        self.data = arr
        self.dims = dims
        if type is None:
            type = 'ket' if arr.shape[1] == 1 and arr.shape[0] > 1 else ('bra' if arr.shape[0] == 1 and arr.shape[1] > 1 else 'oper')
        self.type = type

    @property
    def shape(self):
        return self.data.shape

    def full(self):
        return self.data

    def dag(self):
        t = {'ket': 'bra', 'bra': 'ket'}.get(self.type, 'oper')
        return Qobj(self.data.conj().T, type=t)

    def proj(self):
        return Qobj(self.data @ self.data.conj().T, type='oper')

    def _binary(self, other, fn):
        if isinstance(other, Qobj):
            return Qobj(fn(self.data, other.data))
        return Qobj(fn(self.data, other))

    def __add__(self, other):
        if not isinstance(other, Qobj) and other == 0:
            return self
        return self._binary(other, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, other):
        return self._binary(other, lambda a, b: a - b)

    def __neg__(self):
        return Qobj(-self.data, type=self.type)

    def __mul__(self, other):
        if isinstance(other, Qobj):
            return Qobj(self.data @ other.data)
        return Qobj(self.data * other, type=self.type)

    def __rmul__(self, other):
        return Qobj(self.data * other, type=self.type)

    __matmul__ = __mul__

    def __truediv__(self, k):
        return Qobj(self.data / k, type=self.type)


def sigmaz():
    return Qobj(np.diag([1.0, -1.0]))


def sigmax():
    return Qobj(np.array([[0, 1.0], [1.0, 0]]))


def identity(d):
    return Qobj(np.eye(int(np.prod(d))))


def destroy(d):
    return Qobj(np.diag(np.sqrt(np.arange(1, d)), 1))


def qzero(dimensions):
    dim = int(np.prod(dimensions))
    return Qobj(np.zeros((dim, dim)))


def zero_ket(dimensions):
    return Qobj(np.zeros((int(np.prod(dimensions)), 1)), type='ket')


def basis(dimensions, n=0):
    if np.isscalar(dimensions):
        v = np.zeros((dimensions, 1))
        v[n, 0] = 1
        return Qobj(v, type='ket')
    out = np.ones((1, 1))
    for d, k in zip(dimensions, n):
        v = np.zeros((d, 1))
        v[k, 0] = 1
        out = np.kron(out, v)
    return Qobj(out, type='ket')


def tensor(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = args[0]
    out = np.ones((1, 1))
    kinds = set()
    for a in args:
        out = np.kron(out, a.data)
        kinds.add(a.type)
    return Qobj(out, type='ket' if kinds == {'ket'} else 'oper')


def ket2dm(ket):
    return ket.proj()


def expect(ops, state):
    rho = state.data if state.type == 'oper' else state.data @ state.data.conj().T

    def one(op):
        val = np.trace(op.data @ rho)
        return float(np.real(val)) if abs(np.imag(val)) < 1e-14 * max(1.0, abs(val)) and np.allclose(op.data, op.data.conj().T) else complex(val)
    if isinstance(ops, (list, tuple)):
        return np.array([one(o) for o in ops])
    return one(ops)


def _liouvillian(H, c_ops):
    """Row-major vectorisation: vec(A rho B) = (A kron B^T) vec(rho)."""
    d = H.shape[0]
    eye = np.eye(d)
    L = -1j * (np.kron(H, eye) - np.kron(eye, H.T))
    for c in c_ops:
        cd_c = c.conj().T @ c
        L += np.kron(c, c.conj()) - 0.5 * np.kron(cd_c, eye) - 0.5 * np.kron(eye, cd_c.T)
    return L


class _SolverResult:
    def __init__(self, times):
        self.times = list(times)
        self.states = []
        self.expect = []


def mesolve(H, rho0, tlist, c_ops=None, e_ops=None, **kw):
    c_ops = [c.data for c in (c_ops or [])]
    rho = rho0.data if rho0.type == 'oper' else rho0.data @ rho0.data.conj().T
    tlist = [float(t) for t in tlist]
    res = _SolverResult(tlist)
    L = _liouvillian(H.data, c_ops)
    d = rho.shape[0]
    vec0 = rho.reshape(-1)
    props = {}
    expectations = [[] for _ in (e_ops or [])]
    for t in tlist:
        dt = t - tlist[0]
        key = round(dt, 12)
        if key not in props:
            props[key] = expm(L * dt)
        r = (props[key] @ vec0).reshape(d, d)
        if e_ops:
            for k, op in enumerate(e_ops):
                expectations[k].append(np.real(np.trace(op.data @ r)))
        else:
            res.states.append(Qobj(r, type='oper'))
    res.expect = [np.array(e) for e in expectations]
    return res


def sesolve(H, psi0, tlist, e_ops=None, **kw):
    tlist = [float(t) for t in tlist]
    res = _SolverResult(tlist)
    expectations = [[] for _ in (e_ops or [])]
    for t in tlist:
        psi = expm(-1j * H.data * (t - tlist[0])) @ psi0.data
        if e_ops:
            for k, op in enumerate(e_ops):
                expectations[k].append(np.real((psi.conj().T @ op.data @ psi)[0, 0]))
        else:
            res.states.append(Qobj(psi, type='ket'))
    res.expect = [np.array(e) for e in expectations]
    return res
