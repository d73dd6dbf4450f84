"""``clevolve`` -- continuous-time comparison path of the reference (``qsextra/qcomo/evolution/clevolve.py:110-192``),
which hands H, the collapse operators and the population operators to QuTiP ``sesolve``/``mesolve``.

Here the master equation is solved on the GPU (``engine/lindblad.py`` -> ``qsx_lindblad_evolve``): sigma is split into
exciton-number sector blocks, each propagated by a scaled Taylor series of exp(hL) with the sparse generator of the
block; there is no ODE tolerance, the result is exact to rounding.  Same arguments, checks and messages as the reference;
the return value is a small ``Result`` with QuTiP's attribute names (``times``, ``expect``, ``states``), states being
NumPy arrays (``State``: ``.full()``, ``.ptrace(sel)``, ``.dims``) because QuTiP is not a dependency.
"""
from __future__ import annotations

import warnings

import numpy as np

from ...engine.lindblad import LindbladProgram
from ...engine.sectors import population_observables
from ...system import ChromophoreSystem, ExcitonicSystem
from ...tools.oracle import ending_sentence
from ...tools.tools import if_scalar_to_list

MAX_STATE_BYTES = 2 << 30


class State(np.ndarray):
    """Dense ket ([D, 1]) or operator ([D, D]) on the reference's tensor-product space, with the few ``Qobj`` methods
    the reference's notebooks use on solver output."""

    def __new__(cls, data, dims):
        obj = np.asarray(data, dtype=complex).view(cls)
        obj.dims = list(dims)
        return obj

    def __array_finalize__(self, obj):
        self.dims = getattr(obj, 'dims', None)

    @property
    def isket(self):
        return self.ndim == 2 and self.shape[1] == 1 and self.shape[0] != 1

    def full(self):
        return np.array(self, dtype=complex)

    def dag(self):
        return State(np.asarray(self).conj().T, self.dims)

    def tr(self):
        return complex(np.trace(np.asarray(self)))

    def ptrace(self, sel):
        """Reduced operator of the subsystems ``sel`` (positions in ``dims``, kept in increasing order)."""
        keep = sorted(if_scalar_to_list(sel))
        dims = list(self.dims)
        rho = np.asarray(self)
        if self.isket:
            rho = rho @ rho.conj().T
        n = len(dims)
        rho = rho.reshape(dims + dims)
        letters_k = [chr(ord('a') + i) for i in range(n)]
        letters_b = [chr(ord('A') + i) if i in keep else letters_k[i] for i in range(n)]
        out = [letters_k[i] for i in keep] + [letters_b[i] for i in keep]
        red = np.einsum(''.join(letters_k) + ''.join(letters_b) + '->' + ''.join(out), rho)
        d = int(np.prod([dims[i] for i in keep]))
        return State(red.reshape(d, d), [dims[i] for i in keep])


class Result:
    """Solver output with QuTiP's attribute names (qutip.solver.Result: ``times``, ``expect``, ``states``)."""

    def __init__(self, times, expect, states):
        self.times = list(times)
        self.expect = expect
        self.states = states
        self.num_expect = len(expect)
        self.solver = 'qsx_lindblad_evolve'

    def __repr__(self):
        return 'Result(times={}, expect={}, states={})'.format(len(self.times), len(self.expect), len(self.states))


def _full_state(system, program: LindbladProgram, state):
    """(kind, array) on the full space: a ket [D] or an operator [D, D]; an electronic-only state is completed with the
    pseudomode states of the system (clevolve.py:80-86)."""
    arr = np.asarray(state.full() if hasattr(state, 'full') else state, dtype=complex)
    n = program.n_sites
    ket = arr.ndim == 1 or (arr.ndim == 2 and arr.shape[1] == 1 and arr.shape[0] != 1)
    arr = arr.reshape(-1) if ket else arr
    D = program.full_dimension
    if arr.shape[0] == D:
        return ket, arr
    if arr.shape[0] != 2 ** n:
        raise ValueError('state has dimension {}; expected {} (electronic) or {} (with pseudomodes)'.format(arr.shape[0], 2 ** n, D))
    modes = np.ones(1, dtype=complex)
    kets = system.get_state_mode()
    for _ in range(n):
        for k in range(len(program.levels)):
            modes = np.kron(modes, np.asarray(kets[k], dtype=complex).reshape(-1))
    if ket:
        return True, np.kron(arr, modes)
    return False, np.kron(arr, np.outer(modes, modes.conj()))


def clevolve(system: ChromophoreSystem | ExcitonicSystem,
             time,
             rates=None,
             measure_populations: bool = True,
             state_overwrite=None,
             verbose: bool = True,
             ) -> Result:
    """Master-equation dynamics of ``system`` (same signature as the reference).

    Parameters
    ----------
    system : ChromophoreSystem | ExcitonicSystem
    time : float | list[float] | numpy.ndarray
        Output times; the propagation starts at ``time[0]`` (as QuTiP's solvers do).
    rates : float | list[float] | None
        ``None``: Schroedinger dynamics.  ExcitonicSystem: collapse operators sqrt(rates) Z_i.  ChromophoreSystem:
        sqrt(rates[k]) a_{ik} for pseudomode k of every chromophore.
    measure_populations : bool
        ``True``: ``result.expect[i]`` is the population of chromophore i at every time; ``False``: ``result.states``
        holds the state at every time (kets for a ket input with ``rates=None``, operators otherwise).
    state_overwrite : array | None
        Ket or operator to propagate instead of ``system.get_e_state()`` (electronic space or full space).
    verbose : bool
    """
    if not type(system) is ChromophoreSystem and not type(system) is ExcitonicSystem:
        raise TypeError('system must be an instance of class ChromophoreSystem or ExcitonicSystem.')
    if not system._validity:
        raise Exception('Make sure all the system parameters are specified and consistent.')
    if type(system) is ChromophoreSystem and system.mode_dict is None:
        warnings.warn("Pseudomodes not specified. Executing the dynamics with system as an ExcitonicSystem.")
        # like the reference (clevolve.py:153-157) the recursion does not forward state_overwrite / verbose
        return clevolve(system.extract_ExcitonicSystem(), time, rates, measure_populations)
    time = if_scalar_to_list(time)
    if type(system) is ExcitonicSystem:
        if not np.isscalar(rates) and rates is not None:
            raise TypeError('For an ExcitonicSystem, rates must be a scalar.')
    elif rates is not None:
        rates = if_scalar_to_list(rates)
        if len(rates) != len(system.mode_dict['omega_mode']):
            raise ValueError('The number of dephasing rates is different from the number of pseudomodes per chromophore.')

    from ...engine.runtime import get_engine
    eng = get_engine()
    program = LindbladProgram.from_system(system, rates)
    n = program.n_sites
    is_ket, full = _full_state(system, program, system.get_e_state() if state_overwrite is None else state_overwrite)
    times = np.asarray(time, dtype=float)
    seg = np.diff(np.concatenate([times[:1], times]))
    if np.any(seg < 0):
        raise ValueError('time must be non-decreasing')
    index = {k: program.full_space_index(k) for k in range(n + 1)}
    D = program.full_dimension
    dims = [2] * n + list(program.levels) * n

    def block_of(ka, kb):
        ia, ib = index[ka], index[kb]
        sel = np.zeros((ia.size, ib.size), dtype=complex)
        va, vb = ia >= 0, ib >= 0
        if is_ket:
            sel[np.ix_(va, vb)] = np.outer(full[ia[va]], full[ib[vb]].conj())
        else:
            sel[np.ix_(va, vb)] = full[np.ix_(ia[va], ib[vb])]
        return sel

    pure_unitary = is_ket and rates is None
    expect = [np.zeros(len(times)) for _ in range(n)] if measure_populations else []
    states = []
    if measure_populations:
        for k in range(n + 1):
            sigma0 = block_of(k, k)
            if not np.any(sigma0):
                continue
            block = program.block(k, k)
            dev = eng.upload_lindblad(program.bind(block))
            obs = eng.upload_observables(*population_observables(block), kind='populations')
            out, _ = eng.evolve_lindblad(dev, eng.to_device(sigma0.reshape(1, -1)), 1, seg, obs=obs, obs_real=True)
            pops = out[0].cpu().numpy()
            for i in range(n):
                expect[i] += pops[:, i]
    else:
        if len(times) * D * (1 if pure_unitary else D) * 16 > MAX_STATE_BYTES:
            raise MemoryError('states of dimension {} at {} times do not fit the host budget; use '
                              'measure_populations=True'.format(D, len(times)))
        dense = np.zeros((len(times), D, D), dtype=complex) if not pure_unitary else None
        kets = np.zeros((len(times), D), dtype=complex) if pure_unitary else None
        for ka in range(n + 1):
            if pure_unitary:
                ia = index[ka]
                va = ia >= 0
                psi0 = np.zeros(ia.size, dtype=complex)
                psi0[va] = full[ia[va]]
                if not np.any(psi0):
                    continue
                dev = eng.upload_lindblad(program.bind_ket(ka))
                _, snap = eng.evolve_lindblad(dev, eng.to_device(psi0.reshape(1, -1)), 1, seg, snapshots=True)
                kets[:, ia[va]] = snap[0].cpu().numpy()[:, va]
                continue
            for kb in range(n + 1):
                sigma0 = block_of(ka, kb)
                if not np.any(sigma0):
                    continue
                block = program.block(ka, kb)
                dev = eng.upload_lindblad(program.bind(block))
                _, snap = eng.evolve_lindblad(dev, eng.to_device(sigma0.reshape(1, -1)), 1, seg, snapshots=True)
                ia, ib = index[ka], index[kb]
                va, vb = ia >= 0, ib >= 0
                got = snap[0].cpu().numpy().reshape(len(times), ia.size, ib.size)
                dense[:, ia[va][:, None], ib[vb][None, :]] = got[:, va][:, :, vb]
        if pure_unitary:
            states = [State(kets[t].reshape(-1, 1), dims) for t in range(len(times))]
        else:
            states = [State(dense[t], dims) for t in range(len(times))]
    ending_sentence(verbose)
    return Result(times.tolist(), expect, states)
