"""Result stand-in: exact probabilities of the measured classical register per circuit, plus counts
sampled from them when shots > 0 (keys: clbit N-1 leftmost, like Qiskit)."""
import numpy as np


class Result:
    def __init__(self, probabilities, n_clbits, shots, seed=None, states=None):
        self.exact_probabilities = probabilities      # list of arrays (index bit j = clbit j) or None
        self.states = states                          # list of DensityMatrix (unmeasured circuits)
        self.n_clbits, self.shots = n_clbits, shots
        self._rng = np.random.default_rng(seed)

    def get_counts(self):
        out = []
        for p, ncl in zip(self.exact_probabilities, self.n_clbits):
            p = np.clip(p, 0, None)
            draws = self._rng.multinomial(self.shots, p / p.sum())
            out.append({format(k, '0{}b'.format(ncl)): int(c) for k, c in enumerate(draws) if c})
        return out[0] if len(out) == 1 else out
