"""``qevolve`` -- collision-model dynamics on the B200 engine (drop-in for the reference's
``qsextra/qcomo/evolution/qevolve.py:430-564``), plus the batched entry point ``qevolve_batch``.

The reference builds one Qiskit circuit per requested time and hands them to ``AerSimulator`` (stochastic
state-vector trajectories because of the mid-circuit ``reset``).  Here the same gate sequence is applied to the
density operator of each exciton-number sector by ``libqsx.so`` (include/qsx.h), incrementally in time, so the
read-outs are the exact expectation values the reference's samples converge to.

Return value (the reference returns Qiskit objects, which cannot exist here):
  * ``shots=None``  -> list of ``CollisionCircuit`` (one per distinct step count, ascending, like :417-427)
  * ``shots=0``     -> ``EvolutionResult`` with exact ``populations`` [T, N] and ``distribution`` [T, 2**N]
                       (the reference's Result carries no data in this case -- Appendix A of SURVEY.md)
  * ``shots>0``     -> ``EvolutionResult`` whose ``get_counts()`` draws ``shots`` outcomes of measuring ``sys_e``
                       from the exact distribution (bitstring with chromophore N-1 leftmost, as Aer formats it)
"""
from __future__ import annotations

import os
import warnings
from dataclasses import dataclass

import numpy as np

from ...system import ChromophoreSystem, ExcitonicSystem
from ...tools.tools import if_scalar_to_list
from ...tools.oracle import ending_sentence
from ...engine import dist as qdist
from ...engine.cabi import RR_MAX_OBS
from ...engine.program import StepProgram, SystemBatch, resolve_time_grid
from ...engine.rr import eligible as rr_eligible
from ...engine.sectors import population_observables, sector_configs


@dataclass
class CollisionCircuit:
    """What ``shots=None`` returns in place of a ``QuantumCircuit``: ``n_steps`` repetitions of ``program``
    (Trotter step(s) + collision round) after an optional state preparation."""
    program: StepProgram
    n_steps: int
    initialize: bool
    system: object

    @property
    def operations(self):
        """Names of the engine operations of one step, in application order."""
        return self.program.bind(self.program.block(0, 0)).labels

    @property
    def num_qubits(self) -> int:
        """Qubits of the circuit the reference builds: sys_e, the pseudomode registers, one collision ancilla when
        there are collisions (qevolve.py:18-43)."""
        return self.program.n_sites + self._mode_bits + (1 if self._collisions else 0)

    @property
    def _mode_bits(self) -> int:
        return sum(len(bits) for bits in self.program.layout.mode.values())

    @property
    def _collisions(self) -> bool:
        from ...engine.program import Damp, Dephase, Reset
        return any(isinstance(t, (Damp, Dephase, Reset)) for t in self.program.terms)

    def qasm(self, member: int = 0, measure: bool = False) -> str:
        """OpenQASM 2.0 text of this circuit at gate level: state preparation (basis states only), then ``n_steps``
        repetitions of the step -- every Pauli rotation exp(-i angle P) of the product formula as basis changes, a CX
        ladder and one ``rz``; every collision as its two-qubit rotations on the ancilla followed by ``reset``
        (qevolve.py:74-348).  Qubit numbering is the reference's: sys_e[i] -> i, pseudomode registers next, ancilla
        last.  ``member`` selects the parameter set of a batched program."""
        from ...engine.program import Damp, Dephase, Reset, Rot
        prog = self.program
        n = prog.n_sites
        anc = n + self._mode_bits
        lines = ['OPENQASM 2.0;', 'include "qelib1.inc";', 'qreg q[{}];'.format(self.num_qubits)]
        if measure:
            lines.append('creg c[{}];'.format(n))

        def qubit_of_bit(bit):
            return anc if bit == prog.layout.ancilla else n + bit

        def rotation(paulis, angle):
            """exp(-i angle P) for P = product of (qubit, 'X'|'Y'|'Z')."""
            if not paulis or angle == 0.0:
                return
            for q, p in paulis:
                if p == 'X':
                    lines.append('h q[{}];'.format(q))
                elif p == 'Y':
                    lines.extend(['sdg q[{}];'.format(q), 'h q[{}];'.format(q)])
            qs = [q for q, _ in paulis]
            for a, b in zip(qs[:-1], qs[1:]):
                lines.append('cx q[{}],q[{}];'.format(a, b))
            lines.append('rz({!r}) q[{}];'.format(2.0 * float(angle), qs[-1]))
            for a, b in reversed(list(zip(qs[:-1], qs[1:]))):
                lines.append('cx q[{}],q[{}];'.format(a, b))
            for q, p in paulis:
                if p == 'X':
                    lines.append('h q[{}];'.format(q))
                elif p == 'Y':
                    lines.extend(['h q[{}];'.format(q), 's q[{}];'.format(q)])

        def letters(x, z, to_qubit, width):
            out = []
            for j in range(width):
                xb, zb = (x >> j) & 1, (z >> j) & 1
                if xb or zb:
                    out.append((to_qubit(j), 'Y' if xb and zb else 'X' if xb else 'Z'))
            return out

        if self.initialize:
            kind = getattr(self.system, 'state_type', None)
            if kind == 'localized excitation':
                lines.append('x q[{}];'.format(int(self.system.state)))
            elif kind not in (None, 'ground'):
                raise ValueError('OpenQASM 2.0 has no gate for an arbitrary state preparation; build the circuit with '
                                 'initialize_circuit=False and prepare the {!r} state separately'.format(kind))
        step = []
        mark = len(lines)
        for term in prog.terms:
            if isinstance(term, Rot):
                rotation(letters(term.xs, term.zs, lambda j: j, n) +
                         letters(term.xb, term.zb, qubit_of_bit, prog.n_bits), term.angle[member])
            elif isinstance(term, Dephase):                       # rzx(2 phi) on (site, ancilla), reset  (:292-305)
                rotation([(term.site, 'Z'), (anc, 'X')], term.phi[member])
                lines.append('reset q[{}];'.format(anc))
            elif isinstance(term, Damp):                          # exp(-i phi/2 XX) exp(-i phi/2 YY), reset  (:307-348)
                m = qubit_of_bit(term.bit)
                rotation([(m, 'X'), (anc, 'X')], term.phi[member] / 2)
                rotation([(m, 'Y'), (anc, 'Y')], term.phi[member] / 2)
                lines.append('reset q[{}];'.format(anc))
            elif isinstance(term, Reset):
                lines.append('reset q[{}];'.format(qubit_of_bit(term.bit)))
        step = lines[mark:]
        lines.extend(step * (self.n_steps - 1) if self.n_steps > 0 else [])
        if self.n_steps == 0:
            del lines[mark:]
        if measure:
            lines.extend('measure q[{0}] -> c[{0}];'.format(i) for i in range(n))
        return '\n'.join(lines) + '\n'


class EvolutionResult:
    """Exact read-outs of the evolution (and sampled counts when ``shots`` > 0)."""

    def __init__(self, steps, dt, populations, distribution, shots=None, seed=None):
        self.steps = np.asarray(steps)
        self.times = self.steps * dt
        self.populations = populations            # [T, N]  P_i(t) = Prob(chromophore i excited)
        self.distribution = distribution          # [T, 2**N] outcome probabilities of measuring sys_e
        self.shots = shots
        self._rng = np.random.default_rng(seed)

    def get_counts(self):
        """List (one dict per emitted time; a bare dict for a single time) of {bitstring: count}."""
        if not self.shots:
            raise ValueError('get_counts() needs shots > 0')
        n = int(np.log2(self.distribution.shape[1]))
        out = []
        for p in self.distribution:
            p = np.clip(p, 0.0, None)
            draws = self._rng.multinomial(self.shots, p / p.sum())
            out.append({format(k, '0{}b'.format(n)): int(c) for k, c in enumerate(draws) if c})
        return out[0] if len(out) == 1 else out


def _validate_system(system):
    """Type / validity checks of qevolve.py:484-487 (same exception types and messages)."""
    if type(system) is not ChromophoreSystem and type(system) is not ExcitonicSystem:
        raise TypeError('system must be an instance of class ChromophoreSystem or ExcitonicSystem.')
    if not system._validity:
        raise Exception('Make sure all the system parameters are specified and consistent.')


def _validate_rates(system, coll_rates):
    """coll_rates checks of qevolve.py:505-512."""
    if type(system) is ExcitonicSystem:
        if not np.isscalar(coll_rates) and coll_rates is not None:
            raise TypeError('For an ExcitonicSystem, coll_rates must be a scalar.')
    elif coll_rates is not None:
        coll_rates = if_scalar_to_list(coll_rates)
        if len(coll_rates) != len(system.mode_dict['omega_mode']):
            raise ValueError('The number of dephasing rates is different from the number of pseudomodes per chromophore.')
    return coll_rates


def initial_amplitudes(system, initialize: bool = True) -> np.ndarray:
    """2**N amplitudes on ``sys_e`` prepared by qevolve.py:45-65 (pseudomodes and ancilla start in |0>)."""
    n = system.system_size
    psi = np.zeros(2 ** n, dtype=complex)
    st = system.state_type if initialize else None
    if st == 'state':
        psi[:] = np.asarray(system.state, dtype=complex)
    elif st == 'delocalized excitation':
        for site, c in enumerate(system.state):
            psi[1 << site] += c
    elif st == 'localized excitation':
        psi[1 << system.state] = 1.0
    else:
        psi[0] = 1.0
    return psi


def run_populations(batch: SystemBatch, psi: np.ndarray, program: StepProgram, step_counts, readout='populations',
                    device_output: bool = False):
    """Evolve every member of ``batch`` from the pure state ``psi`` [B, 2**N] and read out after ``step_counts``
    (sorted, distinct).  -> [B, T, N] populations or [B, T, 2**N] distributions (numpy, or a CUDA tensor)."""
    import torch
    from ...engine.runtime import get_engine
    eng = get_engine()
    n, B = batch.n_sites, batch.batch
    steps = np.asarray(step_counts, dtype=np.int32)
    width = n if readout == 'populations' else 2 ** n
    total = None
    emit = eng.to_device(steps, torch.int32)
    inst_set = torch.arange(B, dtype=torch.int32, device=eng.device) if B > 1 else None
    if B == 0:                                        # a rank without parameter sets (more ranks than sets)
        empty = torch.zeros((0, len(steps), width), dtype=torch.float64, device=eng.device)
        return empty if device_output else empty.cpu().numpy()
    sectors = [k for k in range(n + 1) if np.any(psi[:, np.array(sector_configs(n, k), dtype=np.int64)] != 0)]
    same_state = B > 1 and not np.any(psi != psi[0])                   # one initial block shared by every instance
    inst_src = torch.zeros(B, dtype=torch.int32, device=eng.device) if same_state else None
    for k in sectors:
        if readout == 'populations' and k == 0:
            continue                                  # nothing is excited in the ground-state block: no population
        block = program.block(k, k)
        # the pass tables are only needed when the register-resident kernel cannot serve the block or the read-outs
        try_rr = rr_eligible(program, block) and width <= RR_MAX_OBS and not os.environ.get('QSX_DISABLE_RR')
        dev = eng.upload_program(program.bind(block, fused=not try_rr), program)
        if try_rr and 'rr' not in dev:
            dev = eng.upload_program(program.bind(block), program)
        cfgs = np.array(sector_configs(n, k), dtype=np.int64)
        amp = psi[:1, cfgs] if same_state else psi[:, cfgs]            # [B or 1, nA]
        sigma0 = np.zeros((amp.shape[0], block.dim_a, block.dim_b), dtype=complex)
        at = np.arange(len(cfgs)) << block.n_bits
        sigma0[:, at[:, None], at[None, :]] = amp[:, :, None] * amp.conj()[:, None, :]
        if readout == 'populations':
            obs = eng.upload_observables(*population_observables(block), kind='populations')
        else:
            obs = eng.upload_observables(*block.config_probability_observables())
        out, _ = eng.evolve(dev, eng.to_device(sigma0.reshape(sigma0.shape[0], -1)), B, int(steps.max()), emit, obs=obs,
                            obs_real=True, inst_set=inst_set, inst_src=inst_src)
        if readout == 'populations':
            total = out if total is None else total + out
        else:
            if total is None:
                total = torch.zeros((B, len(steps), width), dtype=torch.float64, device=eng.device)
            total[:, :, torch.as_tensor(cfgs, device=eng.device)] += out
    if total is None:                                 # only the ground state is populated
        total = torch.zeros((B, len(steps), width), dtype=torch.float64, device=eng.device)
    return total if device_output else eng.to_host(total)


def _program_for(system, dt, dt_trotter, Trotter_number, Trotter_order, coll_rates) -> StepProgram:
    batch = SystemBatch.from_systems([system])
    rates = None if coll_rates is None else np.atleast_1d(np.asarray(coll_rates, dtype=float))[None]
    return StepProgram(batch, dt, Trotter_number, Trotter_order, rates, dt_trotter=dt_trotter)


def qevolve(system: ChromophoreSystem | ExcitonicSystem,
            time,
            shots: int | None = None,
            initialize_circuit: bool = True,
            Trotter_number: int = 1,
            Trotter_order: int = 1,
            dt: float = None,
            coll_rates=None,
            GPU: bool = False,
            verbose: bool = True,
            ):
    """Collision-model simulation of the dynamics of ``system`` (same signature as the reference).

    Parameters
    ----------
    system : ChromophoreSystem | ExcitonicSystem
    time : float | list[float] | numpy.ndarray
        Target time(s); each must be a multiple of ``dt`` (tolerance 1e-10) or ``ValueError`` is raised.
    shots : int | None
        ``None``: return the step programs; ``0``: exact read-outs; ``> 0``: exact read-outs plus sampled counts.
    initialize_circuit : bool
        Prepare the system state first.
    Trotter_number, Trotter_order : int
        Trotter steps per ``dt`` and product-formula order (1 = Lie-Trotter, even = Suzuki).
    dt : float
        Collision interval; default ``(time[1] - time[0]) / Trotter_number``.
    coll_rates : float | list[float] | None
        ``None``: unitary dynamics.  ExcitonicSystem: dephasing rate.  ChromophoreSystem: pseudomode damping rate(s).
    GPU : bool
        Accepted for compatibility; execution is always on the GPU.
    verbose : bool
        Print progress lines.
    """
    say = print if verbose else (lambda *a, **k: None)
    _validate_system(system)
    if type(system) is ChromophoreSystem and system.mode_dict is None:
        warnings.warn("Pseudomodes not specified. Executing the dynamics with system as an ExcitonicSystem.")
        return qevolve(system.extract_ExcitonicSystem(), time, shots, initialize_circuit, Trotter_number,
                       Trotter_order, dt, coll_rates, GPU, verbose)
    coll_rates = _validate_rates(system, coll_rates)

    say('Start creating the circuits...')
    dt, dt_trotter, Trotter_number, counts = resolve_time_grid(time, dt, Trotter_number)
    program = _program_for(system, dt, dt_trotter, Trotter_number, Trotter_order, coll_rates)
    steps = sorted(set(int(c) for c in counts))
    say('Circuits created...')
    if shots is None:
        ending_sentence(verbose)
        return [CollisionCircuit(program, s, initialize_circuit, system) for s in steps]
    if shots:
        say('Start measuring the circuits...')
    psi = initial_amplitudes(system, initialize_circuit)[None]
    dist = run_populations(program.batch, psi, program, steps, readout='distribution')[0]
    n = system.system_size
    occupied = np.array([[(v >> i) & 1 for v in range(2 ** n)] for i in range(n)], dtype=float)
    result = EvolutionResult(steps, dt, dist @ occupied.T, dist, shots=shots)
    ending_sentence(verbose)
    return result


def qevolve_batch(systems, time, Trotter_number: int = 1, Trotter_order: int = 1, dt: float = None, coll_rates=None,
                  initialize_circuit: bool = True, device_output: bool = False, shard: bool = True):
    """Population dynamics of MANY parameter sets in one launch (new API; the reference loops ``qevolve``).

    ``systems``: list of ExcitonicSystem / ChromophoreSystem with the same structure and state type, or a
    ``SystemBatch`` (then ``initial_state`` is site 0 excited unless ``systems.psi`` is set).
    ``coll_rates``: one value (list) for all, or an array with a leading batch axis.
    Returns populations [B, T, N] for the distinct requested times in ascending order.  Under
    ``torch.distributed`` (and ``shard=True``) the parameter sets are sharded over the ranks and gathered with one
    collective; ``shard=False`` evolves the whole batch on the calling rank.
    """
    if isinstance(systems, SystemBatch):
        batch = systems
        psi = getattr(systems, 'psi', None)
        if psi is None:
            psi = np.zeros((batch.batch, 2 ** batch.n_sites), dtype=complex)
            psi[:, 1] = 1.0
    else:
        systems = list(systems)
        for s in systems:
            _validate_system(s)
        batch = SystemBatch.from_systems(systems)
        psi = np.array([initial_amplitudes(s, initialize_circuit) for s in systems])
    B = batch.batch
    rates = None
    if coll_rates is not None:
        rates = np.asarray(coll_rates, dtype=float)
        width = batch.n_modes if batch.chromophore else 1
        rates = np.broadcast_to(rates.reshape(-1, width) if rates.size != B * width else rates.reshape(B, width),
                                (B, width)).copy()
        if not batch.chromophore:
            rates = rates.reshape(B)
    dt, dt_trotter, Trotter_number, counts = resolve_time_grid(time, dt, Trotter_number)
    steps = counts.tolist() if np.all(np.diff(counts) > 0) else np.unique(counts).tolist()

    rank, size = qdist.world() if shard else (0, 1)
    lo, hi = qdist.shard_bounds(B, rank, size)
    sub = SystemBatch(batch.energies[lo:hi], batch.couplings[lo:hi], batch.dipole_moments[lo:hi],
                      None if batch.omega is None else batch.omega[lo:hi],
                      None if batch.coupl_ep is None else batch.coupl_ep[lo:hi], batch.levels)
    program = StepProgram(sub, dt, Trotter_number, Trotter_order, None if rates is None else rates[lo:hi],
                          dt_trotter=dt_trotter)
    local = run_populations(sub, psi[lo:hi], program, steps, readout='populations', device_output=True)
    full = qdist.all_gather_rows(local, B) if size > 1 else local
    if device_output:
        return full
    from ...engine.runtime import get_engine
    return get_engine().to_host(full)
