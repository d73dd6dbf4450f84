"""Seeded random systems through every host compiler (pass tables, register-resident mappings, CTA-wide programs,
transposed programs): each compiled form, emulated in NumPy, must act on random blocks exactly like the plain operation
list, which the other CPU tests tie to the goldens of the unmodified reference."""
import zlib

import numpy as np
import pytest

import emulator
from qsextra_b200 import ChromophoreSystem, ExcitonicSystem
from qsextra_b200.engine import rr as RR, rrc as RRC
from qsextra_b200.qcomo.evolution.qevolve import _program_for

CASES = []
for n in (1, 2, 3, 4, 5):
    for modes in ([], [2], [3], [2, 2], [4], [2, 3]):
        if n * sum(int(np.ceil(np.log2(d))) for d in modes) > 6:       # keep the blocks small enough for NumPy
            continue
        for order in (1, 2):
            CASES.append((n, tuple(modes), order))


def make_system(rng, n, modes):
    coup = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            if j == i + 1 or rng.random() < 0.4:
                coup[i, j] = coup[j, i] = rng.uniform(-0.3, 0.3)
    kw = dict(energies=rng.uniform(1.0, 2.0, n).tolist(), couplings=coup.tolist() if n != 2 else float(coup[0, 1]),
              dipole_moments=[1.] * n)
    if not modes:
        return ExcitonicSystem(**kw)
    return ChromophoreSystem(frequencies_pseudomode=rng.uniform(0.1, 0.5, len(modes)).tolist(), levels_pseudomode=list(modes),
                             couplings_ep=rng.uniform(0.1, 0.4, len(modes)).tolist(), **kw)


@pytest.mark.parametrize('case', CASES, ids=['n{}_modes{}_order{}'.format(n, 'x'.join(map(str, m)) or '0', o) for n, m, o in CASES])
def test_every_compiled_form_equals_the_operation_list(case):
    n, modes, order = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()))
    system = make_system(rng, n, modes)
    rates = float(rng.uniform(0.05, 0.5)) if not modes else rng.uniform(0.1, 1.0, len(modes)).tolist()
    tn = int(rng.integers(1, 3))
    dt = float(rng.uniform(0.05, 0.3))
    prog = _program_for(system, dt, dt / tn, tn, order, rates)
    sectors = sorted({0, 1, min(2, n), n})
    pairs = [(ka, kb) for ka in sectors for kb in sectors if abs(ka - kb) <= 1][:6]
    checked = set()
    for ka, kb in pairs:
        blk = prog.block(ka, kb)
        if blk.size > 1 << 14:
            continue
        plain = prog.bind(blk, fused=False)
        sigma = rng.normal(size=(blk.dim_a, blk.dim_b)) + 1j * rng.normal(size=(blk.dim_a, blk.dim_b))
        want = emulator.apply_step(sigma, plain)
        scale = max(1.0, np.abs(want).max())
        fused = prog.bind(blk)
        if fused.passes is not None:
            assert np.abs(emulator.apply_pass_program(sigma, fused) - want).max() < 1e-12 * scale, ('passes', ka, kb)
            checked.add('passes')
        rrp = RR.build(prog, plain)
        if rrp is not None:
            assert np.abs(emulator.apply_rr_program(sigma, rrp) - want).max() < 1e-12 * scale, ('rr', ka, kb)
            checked.add('rr')
        big = RRC.build(prog, plain)
        if big is not None:
            assert np.abs(emulator.apply_rrc_program(sigma, big) - want).max() < 1e-12 * scale, ('rrc', ka, kb)
            checked.add('rrc')
        # transposed program: <W, Phi sigma> = <Phi^T W, sigma>
        back = plain.transposed()
        w = rng.normal(size=sigma.shape) + 1j * rng.normal(size=sigma.shape)
        lhs, rhs = np.sum(w * want), np.sum(emulator.apply_step(w, back) * sigma)
        assert abs(lhs - rhs) < 1e-11 * max(1.0, abs(lhs)), ('transposed', ka, kb)
        big_t = RRC.build(prog, back)
        if big_t is not None:
            assert np.abs(emulator.apply_rrc_program(w, big_t) - emulator.apply_step(w, back)).max() < 1e-12 * max(1.0, np.abs(w).max())
