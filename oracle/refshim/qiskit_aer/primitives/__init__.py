"""Estimator stand-in: exact <observable> of each circuit's final state (shots ignored = shots -> inf)."""
import numpy as np

from qiskit_aer import execute_circuit, _Job


class _EstimatorResult:
    def __init__(self, values):
        self.values = np.array(values)


class Estimator:
    def __init__(self, run_options=None, **kw):
        self.run_options = run_options or {}
        self._cache = {}            # id(circuit) -> (circuit kept alive, final state)

    def run(self, circuits, observables, **kw):
        if not (circuits and id(circuits[0]) in self._cache):
            self._cache = {}
        values = []
        for qc, obs in zip(circuits, observables):
            hit = self._cache.get(id(qc))
            if hit is None:
                hit = self._cache[id(qc)] = (qc, execute_circuit(qc)[0])
            dm = hit[1]
            val = 0.0
            for label, coeff in obs.to_list():
                val += np.real(coeff * dm.expectation_pauli(label))
            values.append(val)
        return _Job(_EstimatorResult(values))
