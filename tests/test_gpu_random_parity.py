"""Seeded random sweep of the public API against the CPU oracle: system size, couplings, states, pseudomode levels and
counts, Trotter number/order, collision rates and sequences are drawn at random, so that kernel-selection corners
(register-resident menu hit / miss, pass program, operation list, several sectors at once) are all crossed.
Tolerance 1e-10 relative (FP64)."""
import zlib

import numpy as np
import pytest

from oracle import reference_path as R
from qsextra_b200 import ChromophoreSystem, ExcitonicSystem
from qsextra_b200.qcomo import qevolve
from qsextra_b200.spectroscopy import FeynmanDiagram, Spectroscopy, qspectroscopy

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def random_system(rng, n, modes, levels=None, ring=False):
    energies = rng.uniform(1.3, 1.7, n).tolist()
    coup = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            if j == i + 1 or ring or rng.random() < 0.5:
                coup[i, j] = coup[j, i] = rng.uniform(-0.2, 0.2)
    mu = rng.uniform(-1.2, 1.2, n).tolist()
    if modes == 0:
        return ExcitonicSystem(energies=energies, couplings=coup.tolist() if n > 2 else float(coup[0, 1]), dipole_moments=mu)
    levels = levels or [2] * modes
    return ChromophoreSystem(energies=energies, couplings=coup.tolist() if n > 2 else float(coup[0, 1]), dipole_moments=mu,
                             frequencies_pseudomode=rng.uniform(0.05, 0.4, modes).tolist(), levels_pseudomode=levels,
                             couplings_ep=rng.uniform(0.05, 0.3, modes).tolist())


def set_random_state(rng, system, kind):
    n = system.system_size
    if kind == 'localized':
        system.set_state('localized excitation', int(rng.integers(n)))
    elif kind == 'delocalized':
        c = rng.normal(size=n) + 1j * rng.normal(size=n)
        system.set_state('delocalized excitation', (c / np.linalg.norm(c)).tolist())
    else:
        c = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
        system.set_state('state', (c / np.linalg.norm(c)).tolist())


EVOLVE_CASES = [  # (n, modes, levels, state kind, Trotter number, Trotter order, with rates)
    (2, 0, None, 'general', 1, 1, True), (2, 0, None, 'delocalized', 2, 2, True), (3, 0, None, 'general', 1, 1, True),
    (4, 0, None, 'general', 2, 1, True), (5, 0, None, 'localized', 1, 2, True), (3, 0, None, 'general', 1, 4, False),
    (2, 1, None, 'general', 1, 1, True), (2, 1, None, 'localized', 3, 1, True), (2, 2, None, 'general', 1, 1, True),
    (3, 1, None, 'general', 1, 1, True), (2, 1, [3], 'localized', 1, 1, True), (2, 1, [4], 'delocalized', 1, 2, True),
    (2, 2, [2, 3], 'localized', 1, 1, True), (3, 1, None, 'delocalized', 2, 2, False), (4, 1, None, 'localized', 1, 1, True),
    (1, 1, [4], 'localized', 1, 1, True), (8, 0, None, 'delocalized', 1, 1, True), (7, 0, None, 'general', 1, 1, True),
]


@pytest.mark.parametrize('case', EVOLVE_CASES, ids=[str(i) for i in range(len(EVOLVE_CASES))])
def test_qevolve_random(engine, case):
    n, modes, levels, kind, tn, order, with_rates = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()))
    system = random_system(rng, n, modes, levels) if n > 1 else ChromophoreSystem(
        energies=[1.4], couplings=[[0.]], dipole_moments=[1.], frequencies_pseudomode=[0.2], levels_pseudomode=levels,
        couplings_ep=[0.25])
    set_random_state(rng, system, kind)
    dt = float(rng.uniform(0.05, 0.4))
    steps = sorted(set(int(s) for s in rng.integers(0, 9, size=5)))
    times = [s * dt for s in steps][::-1] + [steps[0] * dt]       # unsorted with a duplicate, as the reference allows
    rates = None
    if with_rates:
        rates = float(rng.uniform(0.02, 0.5)) if modes == 0 else rng.uniform(0.05, 1.0, modes).tolist()
    kw = dict(dt=dt, Trotter_number=tn, Trotter_order=order, coll_rates=rates)
    got = qevolve(system, times, shots=0, verbose=False, **kw)
    want = R.qevolve_exact(system, times, **kw)
    assert list(got.steps) == list(want['steps'])
    assert np.abs(got.distribution - want['distribution']).max() < RTOL
    assert np.abs(got.populations - want['populations']).max() < RTOL


SPECTRO_CASES = [  # (n, modes, label or (side, direction), Trotter number, with rates)
    (2, 0, 'a', 1, True), (3, 0, 'gsb', 1, True), (3, 0, 'se', 2, True), (3, 0, 'esa', 1, True), (4, 0, 'esa', 1, False),
    (2, 1, 'esa', 1, True), (2, 1, ('kbkk', 'iiio'), 1, True), (2, 2, 'se', 1, True), (3, 1, 'gsb', 1, True),
    (2, 0, ('kk', 'io'), 1, True), (2, 0, ('bkbkk', 'iioio'), 1, True), (3, 0, ('kkkk', 'iooi'), 1, True),
    (2, [3], 'esa', 1, True), (2, [4], 'se', 1, True),          # multi-qubit pseudomode registers: Y rotations, resets
    (6, 0, 'esa', 1, True), (5, 0, ('kbkk', 'iiio'), 1, True),  # larger aggregates: 15 x 6 and 10 x 5 configuration blocks
]


@pytest.mark.parametrize('picture', ['planned', 'heisenberg'])
@pytest.mark.parametrize('case', SPECTRO_CASES, ids=[str(i) for i in range(len(SPECTRO_CASES))])
def test_qspectroscopy_random(engine, case, picture, monkeypatch):
    if picture == 'heisenberg':                     # read-out propagated with the transposed step, whatever the size
        monkeypatch.setenv('QSX_HEISENBERG_MIN', '1')
        monkeypatch.setenv('QSX_HEISENBERG_ALWAYS', '1')
        monkeypatch.setenv('QSX_DISABLE_RR', '1')
    n, modes, label, tn, with_rates = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()))
    levels = modes if isinstance(modes, list) else None
    modes = len(levels) if levels else modes
    system = random_system(rng, n, modes, levels, ring=True)
    dt = float(rng.uniform(0.2, 1.0))
    n_delays = (len(label[0]) - 1) if isinstance(label, tuple) else (1 if label == 'a' else 3)
    delays = [[int(s) * dt for s in rng.integers(0, 5, size=rng.integers(1, 4))] for _ in range(n_delays)]
    if isinstance(label, tuple):
        spec = Spectroscopy(delays)
        spec.side_seq, spec.direction_seq = label
    else:
        spec = FeynmanDiagram(label, delays)
    rates = None
    if with_rates:
        rates = float(rng.uniform(0.01, 0.2)) if modes == 0 else rng.uniform(0.05, 0.5, modes).tolist()
    kw = dict(dt=dt, Trotter_number=tn, coll_rates=rates)
    got = qspectroscopy(system, spec, shots=0, verbose=False, **kw)
    want = R.qspectroscopy_operator(system, spec, **kw)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < RTOL * max(1.0, np.abs(want).max())
