"""AerSimulator stand-in: exact density-matrix execution of the recorded circuits (oracle/densesim.py)."""
import numpy as np

from oracle.densesim import DensityMatrix, place_label
from qiskit.result.result import Result


class AerError(Exception):
    pass


def execute_circuit(qc):
    """Run one recorded circuit from |0...0>; returns (DensityMatrix, measured qubits by clbit)."""
    n = qc.num_qubits
    dm = DensityMatrix(n)
    measured = {}
    for name, params, qs, cs in qc.data:
        if name == 'pauli_rot':
            dm.pauli_rotation(place_label(params[0], list(qs), n), params[1])
        elif name == 'rzx':
            dm.rzx(params[0], qs[0], qs[1])
        elif name == 'x':
            dm.x(qs[0])
        elif name == 'h':
            dm.h(qs[0])
        elif name == 'cx':
            dm.cx(qs[0], qs[1])
        elif name == 'cy':
            dm.cy(qs[0], qs[1])
        elif name == 'reset':
            dm.reset(qs[0])
        elif name == 'initialize':
            dm.initialize(np.array(params[0]), list(qs))
        elif name == 'measure':
            for q, c in zip(qs, cs):
                measured[c] = q
        else:
            raise AerError('unsupported instruction ' + name)
    return dm, measured


class _Job:
    def __init__(self, result):
        self._result = result

    def result(self):
        return self._result


class AerSimulator:
    def __init__(self, method='automatic', **options):
        self.method, self.options = method, dict(options)

    def set_option(self, **kw):
        if kw.get('device') == 'GPU':
            raise AerError('GPU device is not available in the refshim')
        self.options.update(kw)

    def run(self, circuits, shots=None, seed_simulator=None, **kw):
        circuits = circuits if isinstance(circuits, (list, tuple)) else [circuits]
        probs, ncl, states = [], [], []
        for qc in circuits:
            dm, measured = execute_circuit(qc)
            states.append(dm)
            if measured:
                order = [measured[c] for c in sorted(measured)]
                probs.append(dm.probabilities(order))
                ncl.append(len(order))
            else:
                probs.append(None)
                ncl.append(0)
        return _Job(Result(probs, ncl, shots, seed=seed_simulator, states=states))
