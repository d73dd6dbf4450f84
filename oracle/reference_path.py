"""TEST INFRASTRUCTURE (oracle) -- exact-expectation evaluation of the reference's hot path on the CPU.

Three evaluators, all FP64 NumPy on ``oracle/densesim.py``:

* ``qevolve_exact``            populations / register distribution after every requested time of the circuits
                               ``qevolve`` builds (qevolve.py:430-564), evolved incrementally.
* ``qspectroscopy_literal``    the reference's own algorithm, circuit by circuit: every grid point, every site
                               pathway, every X/Y variant, each circuit simulated from |0..0> with the ketbra
                               ancilla, <X> + i<Y> weighted by the pathway coefficient (qspectroscopy.py:117-183).
                               Cost grows like the reference's; small grids only.
* ``qspectroscopy_operator``   same numbers from the operator form (collective dipole operators on the ket x bra
                               block, shared prefixes along t1 -> t2 -> t3); checked against the literal form in
                               tests/ and used where the literal form is too slow.

PARITY UNPINNED against real Qiskit/Aer (not installable here); pinned against the unmodified reference code
run over ``oracle/refshim`` (tests/golden/) and against the notebook known-answers listed in SURVEY.md section 4.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import circuits as C
from .densesim import DensityMatrix


def run_ops(dm: DensityMatrix, ops) -> DensityMatrix:
    for op in ops:
        kind = op[0]
        if kind == 'rot':
            dm.pauli_rotation(op[1], op[2])
        elif kind == 'reset':
            dm.reset(op[1])
        elif kind == 'x':
            dm.x(op[1])
        elif kind == 'h':
            dm.h(op[1])
        elif kind == 'cx':
            dm.cx(op[1], op[2])
        elif kind == 'cy':
            dm.cy(op[1], op[2])
        elif kind == 'init':
            dm.initialize(op[1], op[2])
        else:
            raise ValueError(kind)
    return dm


# ---------------------------------------------------------------------------------------------
# qevolve
# ---------------------------------------------------------------------------------------------
def qevolve_exact(system, time, Trotter_number=1, Trotter_order=1, dt=None, coll_rates=None,
                  initialize_circuit=True):
    """Exact read-outs of the circuits qevolve emits.

    Returns dict: ``steps`` (distinct step counts, ascending -- one circuit each, qevolve.py:417-427),
    ``populations`` [T, N] (P_i = Prob(e_i = 1)), ``distribution`` [T, 2**N] (measurement of ``sys_e``;
    index bit i = chromophore i), ``trace`` [T].
    """
    coll_rates = C.normalise_coll_rates(system, coll_rates)
    dt, dt_trotter, trotter_number, counts = C.resolve_time_grid(time, dt, Trotter_number)
    layout = C.Layout(system, coll_rates is not None)
    step = C.step_ops(system, layout, dt, dt_trotter, trotter_number, Trotter_order, coll_rates)
    dm = DensityMatrix(layout.n_qubits)
    if initialize_circuit:
        run_ops(dm, C.initial_state_ops(system, layout))
    wanted = sorted(set(int(c) for c in counts))
    out = {'steps': np.array(wanted), 'populations': [], 'distribution': [], 'trace': []}
    n = system.system_size
    for ndt in range(max(wanted) + 1):
        if ndt:
            run_ops(dm, step)
        if ndt in wanted:
            dist = dm.probabilities(layout.sys)
            out['distribution'].append(dist)
            out['populations'].append([dist[[(v >> i) & 1 == 1 for v in range(1 << n)]].sum() for i in range(n)])
            out['trace'].append(dm.trace().real)
    for key in ('populations', 'distribution', 'trace'):
        out[key] = np.array(out[key])
    return out


# ---------------------------------------------------------------------------------------------
# qspectroscopy, literal
# ---------------------------------------------------------------------------------------------
def _evolution_ops(system, layout, delay, qevolve_kwds):
    """ops of ``qevolve(system, delay, shots=None, initialize_circuit=False, **kwds)[0]`` (qspectroscopy.py:133-139)."""
    coll_rates = C.normalise_coll_rates(system, qevolve_kwds.get('coll_rates'))
    dt, dt_trotter, trotter_number, counts = C.resolve_time_grid(
        delay, qevolve_kwds.get('dt'), qevolve_kwds.get('Trotter_number', 1))
    step = C.step_ops(system, layout, dt, dt_trotter, trotter_number,
                      qevolve_kwds.get('Trotter_order', 1), coll_rates)
    return step * int(counts[0]), step, int(counts[0])


def qspectroscopy_literal(system, spectroscopy, **qevolve_kwds):
    """signal[idx] = sum over pathway circuits of coeff * (<X_kb> + i <Y_kb>)  (qspectroscopy.py:180-183)."""
    n = system.system_size
    delays = spectroscopy.delay_time
    layout = C.Layout(system, qevolve_kwds.get('coll_rates') is not None, offset=1)
    nq = layout.n_qubits
    x_label = 'I' * (nq - 1) + 'X'
    y_label = 'I' * (nq - 1) + 'Y'
    signal = np.zeros([len(t) for t in delays], dtype=complex)
    pathways = list(itertools.product(range(n), repeat=len(delays) + 1))
    for idx in itertools.product(*[range(len(t)) for t in delays]):
        evo = [_evolution_ops(system, layout, delays[k][idx[k]], qevolve_kwds)[0] for k in range(len(delays))]
        total = 0j
        for path in pathways:
            for ops, coeff in C.pathway_circuits(system, spectroscopy.side_seq, spectroscopy.direction_seq, evo, path):
                dm = run_ops(DensityMatrix(nq), ops)
                total += coeff * (dm.expectation_pauli(x_label).real + 1j * dm.expectation_pauli(y_label).real)
        signal[idx] += total
    return signal


# ---------------------------------------------------------------------------------------------
# qspectroscopy, operator form
# ---------------------------------------------------------------------------------------------
def _site_operator(nq: int, qubit: int, mat: np.ndarray) -> np.ndarray:
    out = np.ones((1, 1), dtype=complex)
    for q in range(nq - 1, -1, -1):
        out = np.kron(out, mat if q == qubit else np.eye(2))
    return out


_SX = np.array([[0, 1], [1, 0]], dtype=complex)
_LOWER = np.array([[0, 1], [0, 0]], dtype=complex)      # |0><1| = (X + iY)/2
_RAISE = _LOWER.T.copy()                                 # |1><0| = (X - iY)/2


def interaction_operators(system, layout, side_seq, direction_seq):
    """Collective operator of every interaction and how it acts on sigma = |ket><bra|.

    From the coefficient rules of qspectroscopy.py:144-174: the first interaction on a side applies X; later
    ones apply (X + iY)/2 = |0><1| when (side == 'k') xor (direction == 'i'), else (X - iY)/2 = |1><0|.
    A ket-side operator A acts as sigma -> A sigma; a bra-side one as sigma -> sigma A (the circuit applies A to
    the bra branch and the coefficient is NOT conjugated).  The last entry is the emission: Tr[(sum mu X) sigma].
    """
    nq, mu = layout.n_qubits, system.dipole_moments
    first = {'k': True, 'b': True}
    out = []
    for n, (side, direction) in enumerate(zip(side_seq, direction_seq)):
        last = n == len(side_seq) - 1
        if first[side]:
            # (a first interaction that is also the emission still takes the first_* branch: X)
            mat = _SX
            first[side] = False
        elif last:
            mat = _SX
        elif (side == 'k') ^ (direction == 'i'):
            mat = _LOWER
        else:
            mat = _RAISE
        op = sum(mu[s] * _site_operator(nq, layout.sys[s], mat) for s in range(system.system_size))
        out.append((side, op))
    return out


def qspectroscopy_operator(system, spectroscopy, **qevolve_kwds):
    """Response function by evolving the ket x bra block (no ketbra ancilla), sharing prefixes."""
    delays = spectroscopy.delay_time
    layout = C.Layout(system, qevolve_kwds.get('coll_rates') is not None, offset=0)
    inter = interaction_operators(system, layout, spectroscopy.side_seq, spectroscopy.direction_seq)
    step = None
    counts, evolutions = [], []
    shared_step = qevolve_kwds.get('dt') is not None    # dt=None: every delay is its own step size (qevolve.py:373-379)
    for k in range(len(delays)):
        row, ops_row = [], []
        for t in delays[k]:
            ops, step, c = _evolution_ops(system, layout, t, qevolve_kwds)
            row.append(c)
            ops_row.append(ops)
        counts.append(row)
        evolutions.append(ops_row)
    signal = np.zeros([len(t) for t in delays], dtype=complex)

    def apply(dm, side, op):
        dm.rho = op @ dm.rho if side == 'k' else dm.rho @ op

    def descend(dm, level, idx):
        """dm = block just before interaction ``level``."""
        side, op = inter[level]
        work = dm.copy()
        apply(work, side, op)
        if level == len(delays):                       # emission: trace after the last operator
            signal[idx] += np.trace(work.rho)
            return
        if not shared_step:                            # no common prefix: evolve every delay from the interaction
            for j in range(len(counts[level])):
                branch = work.copy()
                run_ops(branch, evolutions[level][j])
                descend(branch, level + 1, idx + (j,))
            return
        order = np.argsort(counts[level], kind='stable')
        done = 0
        for j in order:
            target = counts[level][j]
            for _ in range(target - done):
                run_ops(work, step)
            done = target
            descend(work, level + 1, idx + (int(j),))

    descend(DensityMatrix(layout.n_qubits), 0, ())
    return signal
