"""Host side of the master-equation comparison path: the dense oracle against the goldens the
unmodified reference produced, and the sector generators / ELL tables of ``engine/lindblad.py`` against that oracle."""
import glob
import os

import numpy as np
import pytest
from scipy.linalg import expm

import emulator
from helpers import build_system, load_golden
from oracle import lindblad as oracle_lindblad
from qsextra_b200.engine.lindblad import LindbladProgram, segments_from_times
from qsextra_b200.spectroscopy import FeynmanDiagram, Spectroscopy

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def spectroscopy_of(meta):
    label = meta['label']
    if isinstance(label, (list, tuple)):
        spec = Spectroscopy(meta['delay_time'])
        spec.side_seq, spec.direction_seq = label
        return spec
    return FeynmanDiagram(label, meta['delay_time'])


@pytest.mark.parametrize('name', ['ex1_dephasing', 'ex1_unitary', 'ex2_pseudomode'])
def test_oracle_reproduces_reference_clevolve(name):
    meta, g = load_golden('clevolve_' + name)
    got = oracle_lindblad.clevolve_exact(build_system(meta['system']), meta['time'], meta['rates'])['populations']
    assert np.abs(got - g['populations']).max() < 1e-13


def test_oracle_reproduces_reference_clspectroscopy():
    seen = 0
    for path in sorted(glob.glob(os.path.join(GOLDEN, 'qspectroscopy_*.npz'))):
        meta, g = load_golden(os.path.basename(path)[:-4])
        if 'signal_cl' not in g:
            continue
        got = oracle_lindblad.clspectroscopy_exact(build_system(meta['system']), spectroscopy_of(meta),
                                                   meta['kwargs'].get('coll_rates'))
        assert np.abs(got - g['signal_cl']).max() < 1e-13, path
        seen += 1
    assert seen >= 10


SYSTEMS = {
    'tetramer_dephasing': (dict(energies=[1., 2., 3., 4.], couplings=[[0, 1, 0, 1], [1, 0, 1, 0], [0, 1, 0, 1], [1, 0, 1, 0]],
                                dipole_moments=[1.] * 4), 0.1),
    'dimer_mode_d3': (dict(energies=[1., 2.], couplings=0.5, dipole_moments=[1., 1.], frequencies_pseudomode=[0.3],
                           levels_pseudomode=[3], couplings_ep=[0.2]), [0.4]),
    'monomer_modes_d3_d2': (dict(energies=[1.3], couplings=[[0.]], dipole_moments=[1.], frequencies_pseudomode=[0.3, 0.7],
                                 levels_pseudomode=[3, 2], couplings_ep=[0.1, 0.2]), [0.4, 0.9]),
    'dimer_unitary_modes': (dict(energies=[1.5, 1.4], couplings=0.1, dipole_moments=[1., 1.], frequencies_pseudomode=[0.2],
                                 levels_pseudomode=[3], couplings_ep=[0.3]), None),
}


@pytest.mark.parametrize('name', sorted(SYSTEMS))
def test_sector_generators_reproduce_the_full_space_master_equation(name):
    sysdef, rates = SYSTEMS[name]
    system = build_system(sysdef)
    program = LindbladProgram.from_system(system, rates)
    space = oracle_lindblad.Space(system)
    assert program.full_dimension == space.dim
    dense = expm(space.liouvillian(rates) * 0.7)
    rng = np.random.default_rng(5)
    n = program.n_sites
    blocks = [b for b in [(0, 0), (1, 0), (1, 1), (2, 1), (n, n)] if max(b) <= n]
    for ka, kb in blocks:
        block = program.block(ka, kb)
        bound = program.bind(block)
        ia, ib = program.full_space_index(ka), program.full_space_index(kb)
        va, vb = ia >= 0, ib >= 0
        sigma = np.zeros((block.dim_a, block.dim_b), dtype=complex)
        sigma[np.ix_(va, vb)] = rng.normal(size=(va.sum(), vb.sum())) + 1j * rng.normal(size=(va.sum(), vb.sum()))
        got = emulator.apply_lindblad(bound, sigma, [0.0, 0.3, 0.4]).reshape(3, block.dim_a, block.dim_b)
        full = np.zeros((space.dim, space.dim), dtype=complex)
        full[np.ix_(ia[va], ib[vb])] = sigma[np.ix_(va, vb)]
        want = (dense @ full.reshape(-1)).reshape(space.dim, space.dim)
        assert np.array_equal(got[0], sigma)
        assert np.abs(got[2][np.ix_(va, vb)] - want[np.ix_(ia[va], ib[vb])]).max() < 1e-12, (ka, kb)
        assert np.abs(got[2][~va]).max(initial=0) == 0 and np.abs(got[2][:, ~vb]).max(initial=0) == 0
        # the block is closed: nothing of the propagated state lies outside it
        outside = want.copy()
        outside[np.ix_(ia[va], ib[vb])] = 0
        assert np.abs(outside).max() < 1e-13
        # ELL tables: padding points at the own row with value 0, every row fits the width
        assert bound.ell_idx.shape == bound.ell_val.shape == (bound.width, block.size)
        assert bound.ell_idx.min() >= 0 and bound.ell_idx.max() < block.size
        assert np.abs(program.generator(block)).sum(axis=0).max() == pytest.approx(bound.norm1)


def test_ket_generator_and_segments():
    sysdef, _ = SYSTEMS['dimer_unitary_modes']
    system = build_system(sysdef)
    program = LindbladProgram.from_system(system, None)
    space = oracle_lindblad.Space(system)
    bound = program.bind_ket(1)
    fi = program.full_space_index(1)
    ok = fi >= 0
    psi = np.zeros(fi.size, dtype=complex)
    psi[ok] = np.random.default_rng(2).normal(size=ok.sum())
    got = emulator.apply_lindblad(bound, psi, [0.9])[0]
    full = np.zeros(space.dim, dtype=complex)
    full[fi[ok]] = psi[ok]
    want = expm(-1j * space.hamiltonian * 0.9) @ full
    assert np.abs(got[ok] - want[fi[ok]]).max() < 1e-12
    assert np.array_equal(segments_from_times([0., 0.5, 2.0]), [0., 0.5, 1.5])
    with pytest.raises(ValueError):
        LindbladProgram([1.] * 2, np.zeros((2, 2)), None, [0.1] * 2, [0.1] * 2, (64, 128))


def test_generator_reproduces_the_heom_curve_of_example_3():
    """K-HEOM on the host: the sparse generator ``engine/lindblad.py`` builds for the 4 x 2-level pseudomode model of
    the notebook 'Example 3 - More is better', propagated with SciPy, against the HEOM curve the reference ships
    (tests/golden/heom_example3.npz <- Benchmark Example 3/data.mat): 1e-6 while t <= 0.15, 1.225e-2 at t = 10.65
    (the survey's figures).  The GPU propagator is checked on the whole curve in tests/test_gpu_lindblad.py."""
    import warnings
    import scipy.sparse.linalg as spl
    from helpers import load_golden
    from qsextra_b200 import ChromophoreSystem
    meta, g = load_golden('heom_example3')
    W = 4
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        s = ChromophoreSystem(energies=meta['energies'], couplings=meta['coupling'], dipole_moments=[1., 1.],
                              frequencies_pseudomode=[0.] * W, levels_pseudomode=[2] * W,
                              couplings_ep=[float(np.sqrt(meta['Gamma'] / W * meta['Omega'] / 2))] * W)
    prog = LindbladProgram.from_system(s, rates=[2 * meta['Omega']] * W)
    blk = prog.block(1, 1)
    gen = prog.generator(blk)
    rho = np.zeros(blk.size, complex)
    rho[0] = 1.0                                        # site 0 excited, pseudomodes empty
    early = spl.expm_multiply(gen, rho, start=0.0, stop=0.15, num=4, endpoint=True)
    pop0 = [np.real(np.diag(v.reshape(blk.dim_a, blk.dim_b))).reshape(blk.n_a, -1).sum(1)[0] for v in early]
    assert np.abs(np.array(pop0) - g['rho11'][:4]).max() < 1e-6
    late = spl.expm_multiply(gen, early[-1], start=0.0, stop=10.5, num=2, endpoint=True)[-1]
    pop = np.real(np.diag(late.reshape(blk.dim_a, blk.dim_b))).reshape(blk.n_a, -1).sum(1)[0]
    k = int(np.argmin(np.abs(g['t'] - 10.65)))
    assert abs(abs(pop - g['rho11'][k]) - 1.225e-2) < 2e-4
