"""PauliEvolutionGate stand-in: exp(-i * time * H), expanded by its synthesis into Pauli rotations."""


class PauliEvolutionGate:
    def __init__(self, operator, time=1.0, label=None, synthesis=None):
        from qiskit.synthesis import LieTrotter
        self.operator, self.time = operator, time
        self.synthesis = synthesis if synthesis is not None else LieTrotter()
        self.num_qubits = operator.num_qubits

    def definition_ops(self):
        n = self.num_qubits
        ops = []
        for label, angle in self.synthesis.expand(self.operator, self.time):
            ops.append(('pauli_rot', (label, float(angle)), list(range(n))))
        return ops
