"""``qevolve(shots=None)`` circuits exported as OpenQASM 2.0: the text is interpreted gate by gate with the oracle's
dense simulator and must reproduce the outcome distribution the unmodified reference produced (tests/golden)."""
import re

import numpy as np
import pytest

from helpers import build_system, golden_names, load_golden, time_arg
from oracle.densesim import DensityMatrix
from qsextra_b200.qcomo import qevolve

S = np.diag([1, 1j])
GATE = re.compile(r'^(\w+)(?:\(([^)]*)\))?\s+(.*);$')


def run_qasm(text: str):
    lines = [ln.strip() for ln in text.strip().splitlines()]
    assert lines[0] == 'OPENQASM 2.0;' and lines[1] == 'include "qelib1.inc";'
    n = int(re.match(r'qreg q\[(\d+)\];', lines[2]).group(1))
    dm = DensityMatrix(n)
    for ln in lines[3:]:
        if ln.startswith('creg') or ln.startswith('measure'):
            continue
        name, arg, operands = GATE.match(ln).groups()
        qs = [int(q) for q in re.findall(r'q\[(\d+)\]', operands)]
        if name == 'h':
            dm.h(qs[0])
        elif name == 'x':
            dm.x(qs[0])
        elif name == 'cx':
            dm.cx(qs[0], qs[1])
        elif name == 's':
            dm.apply_matrix(S, qs)
        elif name == 'sdg':
            dm.apply_matrix(S.conj(), qs)
        elif name == 'rz':
            lam = float(arg)
            dm.apply_matrix(np.diag([np.exp(-0.5j * lam), np.exp(0.5j * lam)]), qs)
        elif name == 'reset':
            dm.reset(qs[0])
        else:
            raise AssertionError('unexpected gate ' + ln)
    return dm, n


def exportable(name):
    meta, _ = load_golden(name)
    state = meta['system'].get('state')
    return state is None or state[0] in ('localized excitation', 'ground')


@pytest.mark.parametrize('name', [n for n in golden_names('qevolve_') if exportable(n)])
def test_exported_circuits_reproduce_reference_distributions(name):
    meta, arrays = load_golden(name)
    system = build_system(meta['system'])
    kw = dict(meta['kwargs'])
    t = kw.pop('time')
    circuits = qevolve(system, time_arg(t), shots=None, verbose=False, **kw)
    assert len(circuits) == len(arrays['distribution'])
    n = system.system_size
    for k in sorted({0, len(circuits) // 2, len(circuits) - 1}):
        text = circuits[k].qasm(measure=True)
        dm, nq = run_qasm(text)
        assert nq == circuits[k].num_qubits
        got = dm.probabilities(list(range(n)))
        assert np.abs(got - arrays['distribution'][k]).max() < 1e-10, (name, k)
        assert text.count('measure') == n


def test_general_initial_states_are_refused_and_can_be_skipped():
    system = build_system(dict(energies=[1., 2.], couplings=0.3, dipole_moments=[1., 1.],
                               state=['delocalized excitation', [0.6, 0.8]]))
    circuit = qevolve(system, [0.2], shots=None, verbose=False, dt=0.1, coll_rates=0.1)[0]
    with pytest.raises(ValueError, match='state preparation'):
        circuit.qasm()
    bare = qevolve(system, [0.2], shots=None, initialize_circuit=False, verbose=False, dt=0.1, coll_rates=0.1)[0]
    text = bare.qasm()
    assert text.count('reset') == 2 * 2 and 'measure' not in text and bare.num_qubits == 3
    assert qevolve(system, [0.0], shots=None, initialize_circuit=False, verbose=False, dt=0.1)[0].qasm().count('\n') == 3
