"""Which rotations of a step may run in tangent form (G = cos(th) (1 - i tan(th) P)) with their cosines multiplied
into the step's diagonal table beforehand.

Pre-scaling element e by the cosines it will meet later is exact only if every operation in between that MIXES two
elements (a rotation, folded or not, or an amplitude damping) pairs elements carrying the same pending factor.
Cosines are compared by value class (rotations with bit-identical cos arrays share a class), so e.g. the two
conditioned pseudomode rotations of a dimer -- different elements, same coupling constant -- still fold.
"""
from __future__ import annotations

import numpy as np


def foldable(n_elem: int, ops: list) -> set:
    """ops: list of dict(mixes=bool, takes_part=bool[n_elem], partner=int[n_elem], cls=int | None) in application order;
    ``cls`` is the cosine's value class for rotations that are candidates for folding, None otherwise.
    Returns the indices of the rotations that may be folded."""
    classes = sorted({op['cls'] for op in ops if op['cls'] is not None})
    candidates = {k for k, op in enumerate(ops) if op['cls'] is not None}
    while True:
        pending = np.zeros((n_elem, max(1, len(classes))), dtype=np.int64)
        bad = set()
        for k in range(len(ops) - 1, -1, -1):
            op = ops[k]
            if op['mixes']:
                tp = op['takes_part']
                differ = np.any(pending[tp] != pending[op['partner'][tp]], axis=0)
                for ci in np.nonzero(differ)[0]:
                    bad.add(classes[ci])
            if k in candidates:
                pending[op['takes_part'], classes.index(op['cls'])] += 1
        drop = {k for k in candidates if ops[k]['cls'] in bad}
        if not drop:
            return candidates
        candidates -= drop


def value_classes(columns: list) -> list:
    """Group parameter columns (arrays over the batch) that are bit-identical; returns one class id per column."""
    reps, out = [], []
    for col in columns:
        for ci, rep in enumerate(reps):
            if np.array_equal(rep, col):
                out.append(ci)
                break
        else:
            reps.append(col)
            out.append(len(reps) - 1)
    return out
