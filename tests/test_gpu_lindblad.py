"""Master-equation comparison path on the GPU (``clevolve`` / ``clspectroscopy`` -> qsx_lindblad_evolve) against the
goldens the unmodified reference produced over the solver stand-in, and against the dense oracle (oracle/lindblad.py).
Tolerance 1e-10 relative (FP64; the propagation is a Taylor series exact to rounding)."""
import glob
import os

import numpy as np
import pytest
import torch

import emulator
from helpers import build_system, load_golden
from oracle import lindblad as oracle_lindblad
from qsextra_b200.engine.lindblad import LindbladProgram
from qsextra_b200.qcomo import clevolve, qevolve
from qsextra_b200.spectroscopy import FeynmanDiagram, clspectroscopy, qspectroscopy
from test_lindblad_host import GOLDEN, SYSTEMS, spectroscopy_of

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize('name', ['ex1_dephasing', 'ex1_unitary', 'ex2_pseudomode'])
def test_clevolve_golden(engine, name):
    meta, g = load_golden('clevolve_' + name)
    res = clevolve(build_system(meta['system']), meta['time'], rates=meta['rates'], verbose=False)
    got = np.array(res.expect).T
    assert got.shape == g['populations'].shape and np.abs(got - g['populations']).max() < TOL
    assert res.times == list(meta['time']) and res.states == []


def test_clspectroscopy_goldens(engine):
    seen = 0
    for path in sorted(glob.glob(os.path.join(GOLDEN, 'qspectroscopy_*.npz'))):
        meta, g = load_golden(os.path.basename(path)[:-4])
        if 'signal_cl' not in g:
            continue
        got = clspectroscopy(build_system(meta['system']), spectroscopy_of(meta), verbose=False,
                             rates=meta['kwargs'].get('coll_rates'))
        assert got.shape == g['signal_cl'].shape
        assert np.abs(got - g['signal_cl']).max() < TOL * max(1.0, np.abs(g['signal_cl']).max()), path
        seen += 1
    assert seen >= 10


@pytest.mark.parametrize('name', sorted(SYSTEMS))
def test_states_against_dense_oracle(engine, name):
    sysdef, rates = SYSTEMS[name]
    sd = dict(sysdef)
    n = len(sd['energies'])
    rng = np.random.default_rng(11)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    system = build_system(sd)
    times = [0.0, 0.25, 0.3, 1.1]                         # not uniform
    # general ket over every sector: populations and full states
    want = oracle_lindblad.clevolve_exact(system, times, rates, state=oracle_lindblad.Space(system).with_modes(psi))
    res = clevolve(system, times, rates=rates, state_overwrite=psi, verbose=False)
    assert np.abs(np.array(res.expect).T - want['populations']).max() < TOL
    full = clevolve(system, times, rates=rates, measure_populations=False, state_overwrite=psi, verbose=False)
    assert len(full.states) == len(times) and full.expect == []
    for t, state in enumerate(full.states):
        rho = state.full() if not state.isket else state.full() @ state.full().conj().T
        assert np.abs(rho - want['states'][t]).max() < TOL
        assert (rates is None) == bool(state.isket)
    # reduced state of chromophore 0 (the notebooks' ptrace(N - i - 1)[1, 1])
    last = full.states[-1]
    assert abs(last.ptrace(n - 1)[1, 1].real - want['populations'][-1, 0]) < TOL
    # an operator as the initial state, given on the electronic space
    rho_e = rng.normal(size=(2 ** n, 2 ** n)) + 1j * rng.normal(size=(2 ** n, 2 ** n))
    sp = oracle_lindblad.Space(system)
    vac = sp.with_modes(np.eye(2 ** n)[0])[: sp.dim // 2 ** n] if sp.levels else np.ones(1)
    rho_full = np.kron(rho_e, np.outer(vac, vac.conj()))
    want_op = oracle_lindblad.clevolve_exact(system, times, rates, state=rho_full)['states']
    got_op = clevolve(system, times, rates=rates, measure_populations=False, state_overwrite=rho_e, verbose=False).states
    assert max(np.abs(got_op[t].full() - want_op[t]).max() for t in range(len(times))) < TOL * 10


def test_kernel_against_emulation_on_random_blocks(engine):
    sysdef, rates = SYSTEMS['dimer_mode_d3']
    program = LindbladProgram.from_system(build_system(sysdef), rates)
    rng = np.random.default_rng(3)
    for ka, kb in [(1, 1), (2, 1), (0, 1)]:
        block = program.block(ka, kb)
        bound = program.bind(block)
        dev = engine.upload_lindblad(bound)
        n_inst = 7                                           # not a multiple of the instances a thread handles
        sigma = rng.normal(size=(3, block.size)) + 1j * rng.normal(size=(3, block.size))
        src = np.array([0, 2, 1, 1, 0, 2, 2], dtype=np.int32)
        seg = [0.0, 0.2, 0.0, 0.9]
        obs = engine.upload_observables(*block.emission_observable(np.array([1.0, 0.7]))) if abs(ka - kb) == 1 else None
        out, snap = engine.evolve_lindblad(dev, engine.to_device(sigma), n_inst, seg, obs=obs, snapshots=True,
                                           inst_src=engine.to_device(src, torch.int32))
        snap = snap.cpu().numpy()
        for i in range(n_inst):
            want = emulator.apply_lindblad(bound, sigma[src[i]], seg)
            assert np.abs(snap[i] - want).max() < 1e-12 * max(1.0, np.abs(want).max())
            if obs is not None:
                ptr, idx, w = obs['host']
                trace = np.array([[np.sum(w[ptr[o]:ptr[o + 1]] * row[idx[ptr[o]:ptr[o + 1]]]) for o in range(len(ptr) - 1)]
                                  for row in want])
                assert np.abs(out[i].cpu().numpy() - trace).max() < 1e-12 * max(1.0, np.abs(trace).max())


def test_collision_model_approaches_the_master_equation(engine):
    """The reason the comparison path exists (Examples 1-2): the q -> cl gap shrinks linearly with dt."""
    system = build_system(dict(energies=[1., 2.], couplings=1., dipole_moments=[1., 1.], frequencies_pseudomode=0.,
                               levels_pseudomode=2, couplings_ep=1., state=['localized excitation', 0]))
    t = np.arange(0, 3.01, 0.5)
    cl = np.array(clevolve(system, t, rates=[2.], verbose=False).expect).T
    gaps = []
    for dt in (0.05, 0.025):
        q = qevolve(system, t, shots=0, dt=dt, coll_rates=[2.], verbose=False).populations
        gaps.append(np.abs(q - cl).max())
    assert gaps[0] < 0.05 and 1.6 < gaps[0] / gaps[1] < 2.4
    spec = FeynmanDiagram('a', np.arange(0, 5., 0.5))
    esys = build_system(dict(energies=[1.55, 1.46], couplings=0.01, dipole_moments=[1., 0.8]))
    s_cl = clspectroscopy(esys, spec, verbose=False, rates=0.02)
    s_q = qspectroscopy(esys, spec, shots=0, verbose=False, dt=0.05, coll_rates=0.02)
    assert np.abs(s_cl - s_q).max() < 5e-3


def test_argument_errors(engine):
    from qsextra_b200.engine import cabi
    esys = build_system(dict(energies=[1., 2.], couplings=1., dipole_moments=[1., 1.], state=['localized excitation', 0]))
    with pytest.raises(TypeError):
        clevolve(esys, [0., 1.], rates=[0.1, 0.2], verbose=False)
    with pytest.raises(TypeError):
        clevolve('system', [0., 1.])
    csys = build_system(dict(energies=[1., 2.], couplings=1., dipole_moments=[1., 1.], frequencies_pseudomode=[0.1],
                             levels_pseudomode=[2], couplings_ep=[0.1], state=['localized excitation', 0]))
    with pytest.raises(ValueError):
        clevolve(csys, [0., 1.], rates=[0.1, 0.2], verbose=False)
    with pytest.raises(ValueError):
        clspectroscopy(csys, FeynmanDiagram('a', [0., 1.]), verbose=False, rates=[0.1, 0.2])
    program = LindbladProgram.from_system(esys, 0.1)
    dev = engine.upload_lindblad(program.bind(program.block(1, 1)))
    sigma = torch.zeros((1, 4), dtype=torch.complex128, device='cuda')
    with pytest.raises(cabi.QsxError, match='nothing to emit'):
        engine.evolve_lindblad(dev, sigma, 1, [0.1])
    with pytest.raises(cabi.QsxError, match='durations'):
        engine.evolve_lindblad(dev, sigma, 1, [-0.1], snapshots=True)
    with pytest.raises(cabi.QsxError, match='Taylor'):
        engine.evolve_lindblad(dev, sigma, 1, [0.1], snapshots=True, terms=1)


def test_clspectroscopy_goldens_with_the_read_out_propagated(engine, monkeypatch):
    """Same goldens with the last delay in the Heisenberg picture (generator transposed), forced for every grid."""
    monkeypatch.setenv('QSX_HEISENBERG_MIN', '1')
    seen = 0
    for path in sorted(glob.glob(os.path.join(GOLDEN, 'qspectroscopy_*.npz'))):
        meta, g = load_golden(os.path.basename(path)[:-4])
        if 'signal_cl' not in g:
            continue
        got = clspectroscopy(build_system(meta['system']), spectroscopy_of(meta), verbose=False,
                             rates=meta['kwargs'].get('coll_rates'))
        assert np.abs(got - g['signal_cl']).max() < TOL * max(1.0, np.abs(g['signal_cl']).max()), path
        seen += 1
    assert seen >= 10


def test_clevolve_against_the_heom_curve_of_example_3(engine):
    """Known-answer test K-HEOM: ``clevolve`` with four 2-level pseudomodes per site against the HEOM
    curve the reference ships (Benchmark Example 3/data.mat, fixture tests/golden/heom_example3.npz).  The pseudomode
    model approximates the HEOM bath, so the agreement is a physics band: 1.5e-2 overall (the survey measured 1.23e-2
    at t = 10.65) and 1e-6 for t <= 0.15, where bath memory has not built up yet."""
    from qsextra_b200 import ChromophoreSystem
    meta, g = load_golden('heom_example3')
    W = 4
    s = ChromophoreSystem(energies=meta['energies'], couplings=meta['coupling'], dipole_moments=[1., 1.],
                          frequencies_pseudomode=[0.] * W, levels_pseudomode=[2] * W,
                          couplings_ep=[float(np.sqrt(meta['Gamma'] / W * meta['Omega'] / 2))] * W)
    s.set_state('localized excitation', 0)
    t = g['t']
    res = clevolve(s, t.tolist(), rates=[2 * meta['Omega']] * W, verbose=False)
    rho11 = np.array(res.expect[0])
    gap = np.abs(rho11 - g['rho11'])
    assert gap.max() < 1.5e-2, gap.max()
    assert gap[t <= 0.15 + 1e-12].max() < 1e-6, gap[:4]
    assert 5e-3 < gap.max()                                # (it IS an approximation of the bath: not a parity check)
