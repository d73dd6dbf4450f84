"""GPU checks at BASELINE.json's FULL sizes through size-independent properties (the oracle only samples them):
normalisation, linearity in the dipoles, sub-grid consistency (prefix sharing), gsb == se at t2 = 0, plus an
oracle comparison on a corner of each grid."""
import numpy as np
import pytest

from oracle import fast, reference_path as R
from qsextra_b200 import ChromophoreSystem, ExcitonicSystem
from qsextra_b200.engine.program import SystemBatch
from qsextra_b200.qcomo import qevolve, qevolve_batch
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy

pytestmark = pytest.mark.gpu
HBAR = 0.658212
DT = 0.1 / HBAR
GAMMA = 59.08e-3


def c2_batch(n_sets, seed=0):
    rng = np.random.default_rng(seed)
    eps = rng.uniform(1.4, 1.6, (n_sets, 2))
    J = rng.uniform(-0.05, 0.05, n_sets)
    omega = rng.uniform(0.01, 0.2, n_sets)
    Gamma = rng.uniform(0.02, 0.1, n_sets)
    Omega = rng.uniform(0.05, 0.2, n_sets)
    coup = np.zeros((n_sets, 2, 2))
    coup[:, 0, 1] = coup[:, 1, 0] = J
    batch = SystemBatch(eps, coup, np.ones((n_sets, 2)), omega[:, None], np.sqrt(Gamma * Omega / 2)[:, None], (2,))
    return batch, (2 * Omega)[:, None]


def test_c1_markovian_dimer_200_steps(engine):
    """BASELINE configs[0] in full against the oracle."""
    s = ExcitonicSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 1.])
    s.set_state('localized excitation', 0)
    t = np.arange(201) * DT
    res = qevolve(s, t, shots=0, dt=DT, coll_rates=GAMMA / 4, verbose=False)
    ref = R.qevolve_exact(s, t, dt=DT, coll_rates=GAMMA / 4)
    assert res.populations.shape == (201, 2)
    assert np.abs(res.populations - ref['populations']).max() < 1e-10
    assert np.abs(res.populations.sum(1) - 1).max() < 1e-12


def test_c2_full_batch_properties_and_sampled_parity(engine):
    """BASELINE configs[1] in full: 4096 parameter sets x 1000 steps."""
    batch, rates = c2_batch(4096)
    t = np.arange(1001) * DT
    pops = qevolve_batch(batch, t, dt=DT, coll_rates=rates)
    assert pops.shape == (4096, 1001, 2)
    assert np.abs(pops.sum(-1) - 1).max() < 1e-11           # one exciton is conserved (pseudomode damping only)
    assert pops.min() > -1e-13 and pops.max() < 1 + 1e-13
    assert np.abs(pops[:, 0] - [1., 0.]).max() == 0
    # parity on a sample of parameter sets against the C oracle (full-space literal gates)
    pick = [0, 1, 2047, 4095]
    systems, r = [], []
    for b in pick:
        s = ChromophoreSystem(energies=batch.energies[b].tolist(), couplings=float(batch.couplings[b, 0, 1]), dipole_moments=[1., 1.],
                              frequencies_pseudomode=float(batch.omega[b, 0]), levels_pseudomode=2, couplings_ep=float(batch.coupl_ep[b, 0]))
        s.set_state('localized excitation', 0)
        systems.append(s)
        r.append([float(rates[b, 0])])
    ref, _ = fast.populations_batch(systems, t, dt=DT, coll_rates=np.array(r))
    assert np.abs(pops[pick] - ref).max() < 1e-10
    # idempotence: a second identical call returns identical bits; a sub-batch equals the slice of the batch
    again = qevolve_batch(batch, t, dt=DT, coll_rates=rates)
    assert np.array_equal(pops, again)
    sub = SystemBatch(batch.energies[100:132], batch.couplings[100:132], batch.dipole_moments[100:132], batch.omega[100:132],
                      batch.coupl_ep[100:132], (2,))
    assert np.array_equal(qevolve_batch(sub, t, dt=DT, coll_rates=rates[100:132]), pops[100:132])


def test_c3_dimer_2des_full_grid(engine):
    """BASELINE configs[2]: 64 x 64 t1/t3 grid at t2 = 0, rephasing pathways."""
    s = ExcitonicSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 0.8])
    grid = (30 * DT * np.arange(64)).tolist()
    kw = dict(dt=DT, coll_rates=GAMMA / 4)
    sig = {lab: qspectroscopy(s, FeynmanDiagram(lab, [grid, [0.], grid]), shots=0, verbose=False, **kw) for lab in ('gsb', 'se', 'esa')}
    mu2 = 1. + 0.8 ** 2
    assert abs(sig['gsb'][0, 0, 0] - mu2 ** 2) < 1e-12 and sig['gsb'].shape == (64, 1, 64)
    assert np.abs(sig['gsb'] - sig['se']).max() < 1e-11           # identical at t2 = 0 for this model
    # linearity: doubling every dipole multiplies a third-order response by 16
    s2 = ExcitonicSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[2., 1.6])
    twice = qspectroscopy(s2, FeynmanDiagram('esa', [grid, [0.], grid]), shots=0, verbose=False, **kw)
    assert np.abs(twice - 16 * sig['esa']).max() < 1e-10 * np.abs(twice).max()
    # sub-grid consistency (prefix sharing) and unsorted / repeated delays
    pick1, pick3 = [5, 0, 63, 5], [7, 62, 1]
    sub = qspectroscopy(s, FeynmanDiagram('esa', [[grid[i] for i in pick1], [0.], [grid[i] for i in pick3]]), shots=0, verbose=False, **kw)
    assert np.abs(sub - sig['esa'][np.ix_(pick1, [0], pick3)]).max() < 1e-12
    # oracle on a corner
    corner = [grid[i] for i in (0, 1, 2)]
    for lab in ('gsb', 'esa'):
        ref = R.qspectroscopy_operator(s, FeynmanDiagram(lab, [corner, [0.], corner]), **kw)
        assert np.abs(sig[lab][:3, :, :3] - ref).max() < 1e-10 * np.abs(ref).max()


def test_c4_chain_with_pseudomodes_reduced_grid(engine):
    """BASELINE configs[3] geometry (4 sites, one pseudomode each) on a reduced grid: all block shapes of the full run
    (16x64, 64x64, 16x16, 64x16, 96x64) through the kernels the full run uses, against the oracle on a corner."""
    J = np.zeros((4, 4))
    for i in range(3):
        J[i, i + 1] = J[i + 1, i] = -0.01
    s = ChromophoreSystem(energies=[1.55, 1.50, 1.46, 1.52], couplings=J, dipole_moments=[1., 0.9, 1.1, 1.],
                          frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=float(np.sqrt(GAMMA * 0.1 / 2)))
    kw = dict(dt=DT, coll_rates=[0.2])
    t1 = (15 * DT * np.arange(6)).tolist()
    t2 = (100 * DT * np.arange(2)).tolist()
    sig = {lab: qspectroscopy(s, FeynmanDiagram(lab, [t1, t2, t1]), shots=0, verbose=False, **kw) for lab in ('gsb', 'se', 'esa')}
    mu2 = float(np.sum(np.array([1., 0.9, 1.1, 1.]) ** 2))
    assert abs(sig['gsb'][0, 0, 0] - mu2 ** 2) < 1e-11
    assert np.abs(sig['gsb'][:, 0, :] - sig['se'][:, 0, :]).max() < 1e-10
    small = [0., 2 * DT]
    for lab in ('gsb', 'se', 'esa'):
        got = qspectroscopy(s, FeynmanDiagram(lab, [small, [0., 3 * DT], small]), shots=0, verbose=False, **kw)
        ref = R.qspectroscopy_operator(s, FeynmanDiagram(lab, [small, [0., 3 * DT], small]), **kw)
        assert np.abs(got - ref).max() < 1e-10 * np.abs(ref).max()


def test_c5_eight_site_aggregate_streaming_path(engine):
    """BASELINE configs[4] geometry: 8 sites, one 2-level pseudomode each, single-exciton sector (2048 x 2048 block =
    64 MiB, does not fit on chip -> streaming path).  Checked against the NumPy emulation of the operation semantics for
    the first steps, through conservation laws for 200 steps, and against the dt -> 0 trend of the collision model."""
    import emulator
    from qsextra_b200.qcomo.evolution.qevolve import _program_for
    n = 8
    J = np.zeros((n, n))
    for i in range(n - 1):
        J[i, i + 1] = J[i + 1, i] = -0.01
    eps = [1.55, 1.50, 1.46, 1.52, 1.49, 1.53, 1.47, 1.51]
    s = ChromophoreSystem(energies=eps, couplings=J, dipole_moments=[1.] * n, frequencies_pseudomode=0.02,
                          levels_pseudomode=2, couplings_ep=float(np.sqrt(GAMMA * 0.1 / 2)))
    s.set_state('localized excitation', 0)
    t = np.arange(201) * DT
    res = qevolve(s, t, shots=0, dt=DT, coll_rates=[0.2], verbose=False)
    pops = res.populations
    assert pops.shape == (201, n)
    assert np.abs(pops.sum(1) - 1).max() < 1e-11 and pops.min() > -1e-13
    assert pops[0, 0] == 1.0 and pops[-1, 0] < 0.99 and pops[-1, 1] > 1e-4       # the excitation moves along the chain
    # first two steps against the emulated operation list on the same 2048 x 2048 block
    prog = _program_for(s, DT, DT, 1, 1, [0.2])
    blk = prog.block(1, 1)
    bound = prog.bind(blk, fused=False)
    sig = np.zeros((blk.dim_a, blk.dim_b), complex)
    sig[0, 0] = 1.0
    for step in (1, 2):
        sig = emulator.apply_step(sig, bound)
        want = [np.real(np.diag(sig)).reshape(n, -1)[i].sum() for i in range(n)]
        assert np.abs(pops[step] - want).max() < 1e-12


def test_c5_collision_model_against_master_equation(engine):
    """BASELINE configs[4] as stated: the 8-site aggregate with 2-level pseudomodes, collision-model dynamics (qevolve,
    streaming path) cross-checked with the master equation (clevolve -> qsx_lindblad_evolve on the same 2048 x 2048
    single-exciton block, 4.2 M elements, generator of 63 M entries).  The gap is the O(dt) discretisation of the
    collision model: it halves when dt halves."""
    from qsextra_b200.qcomo import clevolve
    n = 8
    J = np.zeros((n, n))
    for i in range(n - 1):
        J[i, i + 1] = J[i + 1, i] = -0.01
    eps = [1.55, 1.50, 1.46, 1.52, 1.49, 1.53, 1.47, 1.51]
    s = ChromophoreSystem(energies=eps, couplings=J, dipole_moments=[1.] * n, frequencies_pseudomode=0.02,
                          levels_pseudomode=2, couplings_ep=float(np.sqrt(GAMMA * 0.1 / 2)))
    s.set_state('localized excitation', 0)
    t = np.arange(0, 61, 20) * DT
    cl = np.array(clevolve(s, t, rates=[0.2], verbose=False).expect).T
    assert cl.shape == (4, n) and np.abs(cl.sum(1) - 1).max() < 1e-10 and cl.min() > -1e-12
    gaps = []
    for sub in (1, 2):
        q = qevolve(s, t, shots=0, dt=DT / sub, coll_rates=[0.2], verbose=False).populations
        gaps.append(np.abs(q - cl).max())
    assert gaps[0] < 2e-3 and 1.7 < gaps[0] / gaps[1] < 2.3, gaps


# ---- far corners of the full grids against the C oracle (literal gates; collisions as channels, tests/test_oracle_golden.py) --
def _c4_system():
    J = np.zeros((4, 4))
    for i in range(3):
        J[i, i + 1] = J[i + 1, i] = -0.01
    return ChromophoreSystem(energies=[1.55, 1.50, 1.46, 1.52], couplings=J, dipole_moments=[1., 1., 1., 1.],
                             frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=float(np.sqrt(GAMMA * 0.1 / 2)))


def test_c4_full_grid_far_corners_against_the_oracle(engine):
    """BASELINE configs[3] at FULL size (128 x 16 x 128, gsb + se + esa, one qspectroscopy_set call) compared with the
    oracle where the delays are longest: the whole t3 rows of the chains (t1, t2) = (127, 15) and (64, 7) -- 1905 + 1500
    + 1905 collision steps at the far corner (127, 15, 127) -- 2 x 128 points per diagram."""
    from qsextra_b200.spectroscopy import qspectroscopy_set
    s = _c4_system()
    t1 = (15 * DT * np.arange(128)).tolist()
    t2 = (100 * DT * np.arange(16)).tolist()
    specs = [FeynmanDiagram(lab, [t1, t2, t1]) for lab in ('gsb', 'se', 'esa')]
    sigs = qspectroscopy_set(s, specs, verbose=False, dt=DT, coll_rates=[0.2])
    rows = [(127, 15), (64, 7)]
    for spec, sig in zip(specs, sigs):
        assert sig.shape == (128, 16, 128)
        ref, _ = fast.response_points(s, spec, rows, DT, coll_rates=[0.2], kraus=True)
        got = np.array([sig[i, j] for i, j in rows])
        scale = np.abs(sig).max()
        assert np.abs(got - ref).max() < 1e-10 * scale, (spec.side_seq, np.abs(got - ref).max() / scale)
        # and relative to the local magnitude of the (decayed) far row itself
        assert np.abs(got - ref).max() < 1e-8 * np.abs(ref).max()


def test_c3_all_six_pathways_far_rows_against_the_oracle(engine):
    """BASELINE configs[2] at full size, the three rephasing and the three non-rephasing pathways, compared with the
    oracle on the t3 rows of t1 = 63 (1890 steps) and t1 = 31."""
    from qsextra_b200.spectroscopy import Spectroscopy
    s = ExcitonicSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 0.8])
    grid = (30 * DT * np.arange(64)).tolist()
    kw = dict(dt=DT, coll_rates=GAMMA / 4)
    specs = [FeynmanDiagram(lab, [grid, [0.], grid]) for lab in ('gsb', 'se', 'esa')]
    for sides, directions in (('kkkk', 'ioio'), ('kbbk', 'iioo'), ('kbkk', 'iiio')):
        extra = Spectroscopy([grid, [0.], grid])
        extra.side_seq, extra.direction_seq = sides, directions
        specs.append(extra)
    rows = [(63, 0), (31, 0)]
    for spec in specs:
        sig = qspectroscopy(s, spec, shots=0, verbose=False, **kw)
        ref, _ = fast.response_points(s, spec, rows, DT, coll_rates=GAMMA / 4, kraus=True)
        got = np.array([sig[i, j] for i, j in rows])
        assert np.abs(got - ref).max() < 1e-10 * np.abs(sig).max(), spec.side_seq


def test_c5_eight_sites_50_steps_against_the_sector_oracle(engine):
    """BASELINE configs[4] geometry (8 sites, one 2-level pseudomode each, 2048 x 2048 single-exciton block, streaming
    path) against the oracle's literal gate list restricted to the single-exciton sector (oracle/sector.py)."""
    from oracle import sector
    n = 8
    J = np.zeros((n, n))
    for i in range(n - 1):
        J[i, i + 1] = J[i + 1, i] = -0.01
    s = ChromophoreSystem(energies=[1.55, 1.50, 1.46, 1.52, 1.49, 1.53, 1.47, 1.51], couplings=J, dipole_moments=[1.] * n,
                          frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=float(np.sqrt(GAMMA * 0.1 / 2)))
    s.set_state('localized excitation', 0)
    steps = [0, 1, 10, 25, 50]
    got = qevolve(s, np.array(steps) * DT, shots=0, dt=DT, coll_rates=[0.2], verbose=False).populations
    ref = sector.populations(s, steps, DT, coll_rates=[0.2])
    assert np.abs(got - ref).max() < 1e-10, np.abs(got - ref).max()
