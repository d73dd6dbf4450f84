"""TEST INFRASTRUCTURE -- CPU restatement of the reference's master-equation comparison path; never used by the product.

Dense NumPy/SciPy restatement of ``qsextra/qcomo/evolution/clevolve.py:16-108`` (operators handed to QuTiP) and of
``qsextra/spectroscopy/simulation/clspectroscopy.py:18-132`` (pathway loops), with the solver replaced by the exact
propagator exp(L t) of the dense Liouvillian on the FULL tensor-product space (no sectors, no sparsity: independent of
``qsextra_b200/engine/lindblad.py``).  QuTiP is absent from this image, so the solver itself is unpinned ("parity
unpinned" for the ODE integration, whose default tolerances are 1e-6 relative / 1e-8 absolute anyway); what the
goldens ``tests/golden/clevolve_*.npz`` and the ``signal_cl`` arrays of ``tests/golden/qspectroscopy_*.npz`` pin is
the reference's own operator construction and pathway logic, run unmodified over ``oracle/refshim``.
"""
from __future__ import annotations

from itertools import product

import numpy as np
from scipy.linalg import expm

I2 = np.eye(2, dtype=complex)
SZ = np.diag([1., -1.]).astype(complex)
SX = np.array([[0., 1.], [1., 0.]], dtype=complex)
SM = np.array([[0., 1.], [0., 0.]], dtype=complex)          # destroy(2) = |0><1|
P1 = np.diag([0., 1.]).astype(complex)                      # basis(2, 1).proj()


def kron(*ops):
    out = np.ones((1, 1), dtype=complex)
    for op in ops:
        out = np.kron(out, op)
    return out


def destroy(d):
    return np.diag(np.sqrt(np.arange(1, d)), 1).astype(complex)


class Space:
    """Operators of the reference on [e_{N-1}..e_0, modes(N-1)..modes(0)] (clevolve.py:22-27, :56-77)."""

    def __init__(self, system):
        self.system = system
        self.n = system.system_size
        md = getattr(system, 'mode_dict', None)
        self.levels = list(md['lvl_mode']) if md is not None else []
        self.eyes = [np.eye(d, dtype=complex) for d in self.levels]
        self.hamiltonian = np.asarray(system.get_global_Hamiltonian() if md is not None else system.get_e_Hamiltonian(),
                                      dtype=complex)
        self.dim = self.hamiltonian.shape[0]

    def site(self, i, op):
        return kron(*[I2] * (self.n - i - 1), op, *[I2] * i, *self.eyes * self.n)

    def mode(self, i, k, op):
        return kron(*[I2] * self.n, *self.eyes * (self.n - i - 1), *self.eyes[:k], op, *self.eyes[k + 1:], *self.eyes * i)

    def collapse(self, rates):
        if rates is None:
            return []
        if not self.levels:
            return [np.sqrt(rates) * self.site(i, SZ) for i in range(self.n)]
        rates = np.atleast_1d(rates)
        return [np.sqrt(rates[k]) * self.mode(i, k, destroy(self.levels[k])) for i in range(self.n) for k in range(len(rates))]

    def liouvillian(self, rates):
        """Row-major vec: (A sigma B) <-> kron(A, B^T)."""
        eye = np.eye(self.dim, dtype=complex)
        gen = -1j * (np.kron(self.hamiltonian, eye) - np.kron(eye, self.hamiltonian.T))
        for c in self.collapse(rates):
            cdc = c.conj().T @ c
            gen += np.kron(c, c.conj()) - 0.5 * np.kron(cdc, eye) - 0.5 * np.kron(eye, cdc.T)
        return gen

    def ground(self):
        psi = np.zeros(self.dim, dtype=complex)
        psi[0] = 1.0
        return psi

    def with_modes(self, e_state):
        modes = np.ones(1, dtype=complex)
        if self.levels:
            kets = self.system.get_state_mode()
            for _ in range(self.n):
                for k in range(len(self.levels)):
                    modes = np.kron(modes, np.asarray(kets[k], dtype=complex).reshape(-1))
        return np.kron(np.asarray(e_state, dtype=complex).reshape(-1), modes)


def clevolve_exact(system, time, rates=None, state=None):
    """-> dict(populations [T, N], states [T, D, D]) of clevolve.py:110-192 with an exact propagator."""
    sp = Space(system)
    psi = sp.with_modes(system.get_e_state()) if state is None else np.asarray(state, dtype=complex)
    rho0 = np.outer(psi, psi.conj()) if psi.ndim == 1 else psi
    gen = sp.liouvillian(rates)
    times = np.atleast_1d(np.asarray(time, dtype=float))
    states = np.array([(expm(gen * (t - times[0])) @ rho0.reshape(-1)).reshape(sp.dim, sp.dim) for t in times])
    pops = np.array([[np.trace(sp.site(i, P1) @ rho).real for i in range(sp.n)] for rho in states])
    return dict(populations=pops, states=states)


def clspectroscopy_exact(system, spectroscopy, rates=None):
    """Literal pathway / grid loops of clspectroscopy.py:82-131 with exact propagators (cached per distinct delay)."""
    sp = Space(system)
    n = sp.n
    gen = sp.liouvillian(rates)
    cache = {}

    def propagate(rho, t):
        if t not in cache:
            cache[t] = expm(gen * t)
        return (cache[t] @ rho.reshape(-1)).reshape(sp.dim, sp.dim)

    mu = np.asarray(system.dipole_moments, dtype=float)
    emission = sum(mu[i] * sp.site(i, SX) for i in range(n))
    rho_ground = np.outer(sp.ground(), sp.ground().conj())
    sizes = [len(t) for t in spectroscopy.delay_time]
    signal = np.zeros(sizes, dtype=complex)
    for sites in product(range(n), repeat=len(sizes)):
        for idx in product(*[range(s) for s in sizes]):
            first_k = first_b = True
            rho = rho_ground
            for level, site in enumerate(sites):
                side, direction = spectroscopy.side_seq[level], spectroscopy.direction_seq[level]
                if side == 'b' and first_b:
                    op, first_b = SX, False
                elif side == 'k' and first_k:
                    op, first_k = SX, False
                elif (side == 'k') ^ (direction == 'i'):
                    op = SM
                else:
                    op = SM.conj().T
                full = sp.site(site, op)
                rho = rho @ full if side == 'b' else full @ rho
                rho = propagate(rho, float(spectroscopy.delay_time[level][idx[level]]))
            signal[idx] += np.trace(emission @ rho) * np.prod(mu[list(sites)])
    return signal
