"""LieTrotter / SuzukiTrotter stand-ins (qiskit 1.0 `qiskit.synthesis.evolution`), reps = 1.

LieTrotter: terms in operator order, each exp(-i * time * Re(coeff) * P).
SuzukiTrotter(order): odd orders rejected; order 2 = half steps of terms[1:] in REVERSED order, the
full step of terms[0], then the half steps mirrored; order 2k by the p_k = 1/(4 - 4^(1/(2k-1)))
recursion (outer, outer, inner, outer, outer).
"""
import numpy as np


class LieTrotter:
    def __init__(self, reps=1):
        self.reps = reps

    def expand(self, operator, time):
        terms = [(l, float(np.real(c))) for l, c in operator.to_list()]
        out = []
        for _ in range(self.reps):
            out += [(l, c * time / self.reps) for l, c in terms]
        return out


class SuzukiTrotter:
    def __init__(self, order=2, reps=1):
        if order % 2 == 1:
            raise ValueError('Suzuki product formulae are symmetric and therefore only defined for even orders.')
        self.order, self.reps = order, reps

    @staticmethod
    def _recurse(order, time, terms):
        if order == 1:
            return terms
        if order == 2:
            halves = [(l, c * time / 2) for l, c in reversed(terms[1:])]
            full = [(terms[0][0], time * terms[0][1])]
            return halves + full + list(reversed(halves))
        reduction = 1 / (4 - 4 ** (1 / (order - 1)))
        outer = 2 * SuzukiTrotter._recurse(order - 2, reduction * time, terms)
        inner = SuzukiTrotter._recurse(order - 2, (1 - 4 * reduction) * time, terms)
        return outer + inner + outer

    def expand(self, operator, time):
        terms = [(l, float(np.real(c))) for l, c in operator.to_list()]
        return self.reps * self._recurse(self.order, time / self.reps, terms)
