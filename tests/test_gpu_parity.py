"""GPU parity: the CUDA path (through the public API and the C ABI) against the golden fixtures produced by the
unmodified reference, and against the CPU oracle on seeded inputs.  Tolerance: 1e-10 relative (FP64), as
BASELINE.json's north_star states."""
import numpy as np
import pytest

from helpers import build_spectroscopy, build_system, golden_names, load_golden, time_arg
from oracle import reference_path as R
from qsextra_b200.qcomo import qevolve, qevolve_batch
from qsextra_b200.spectroscopy import qspectroscopy

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def close(a, b, scale=None):
    scale = max(1.0, np.abs(b).max()) if scale is None else scale
    assert np.abs(np.asarray(a) - np.asarray(b)).max() <= RTOL * scale, np.abs(np.asarray(a) - np.asarray(b)).max()


@pytest.mark.parametrize('name', golden_names('qevolve_'))
def test_qevolve_matches_reference_golden(engine, name):
    meta, arrays = load_golden(name)
    system = build_system(meta['system'])
    kw = dict(meta['kwargs'])
    t = kw.pop('time')
    res = qevolve(system, time_arg(t), shots=0, verbose=False, **kw)
    close(res.distribution, arrays['distribution'])
    n = system.system_size
    pops = np.array([[arrays['distribution'][k][[(v >> i) & 1 == 1 for v in range(2 ** n)]].sum() for i in range(n)]
                     for k in range(len(res.steps))])
    close(res.populations, pops)


@pytest.mark.parametrize('name', golden_names('qspectroscopy_'))
def test_qspectroscopy_matches_reference_golden(engine, name):
    meta, arrays = load_golden(name)
    system = build_system(meta['system'])
    spec = build_spectroscopy(meta['label'], meta['delay_time'])
    signal = qspectroscopy(system, spec, shots=0, verbose=False, **dict(meta['kwargs']))
    assert signal.shape == arrays['signal'].shape
    close(signal, arrays['signal'], scale=max(1.0, np.abs(arrays['signal']).max()))


def test_qevolve_shots_counts(engine):
    meta, arrays = load_golden('qevolve_ex1_dephasing')
    system = build_system(meta['system'])
    kw = dict(meta['kwargs'])
    t = kw.pop('time')
    res = qevolve(system, t, shots=20000, verbose=False, **kw)
    counts = res.get_counts()
    assert len(counts) == len(t)
    n = system.system_size
    for c, p in zip(counts, arrays['distribution']):
        assert sum(c.values()) == 20000
        for i in range(n):
            key = '{:b}'.format(1 << i).zfill(n)
            assert abs(c.get(key, 0) / 20000 - p[1 << i]) < 5 * np.sqrt(max(p[1 << i] * (1 - p[1 << i]), 1e-4) / 20000)


def test_batch_random_parameter_sets_vs_oracle(engine):
    """C2-shaped workload at test size: non-Markovian dimers with one pseudomode per site, seeded parameters."""
    from qsextra_b200 import ChromophoreSystem
    rng = np.random.default_rng(0)
    B, steps = 6, 40
    dt = 0.1 / 0.658212
    systems, rates = [], []
    for _ in range(B):
        gam, om = rng.uniform(0.02, 0.1), rng.uniform(0.05, 0.2)
        s = ChromophoreSystem(energies=rng.uniform(1.4, 1.6, 2).tolist(), couplings=float(rng.uniform(-0.05, 0.05)),
                              dipole_moments=[1., 1.], frequencies_pseudomode=float(rng.uniform(0.01, 0.2)),
                              levels_pseudomode=2, couplings_ep=float(np.sqrt(gam * om / 2)))
        s.set_state('localized excitation', 0)
        systems.append(s)
        rates.append([2 * om])
    time = np.arange(steps + 1) * dt
    pops = qevolve_batch(systems, time, dt=dt, coll_rates=np.array(rates))
    assert pops.shape == (B, steps + 1, 2)
    for b in range(B):
        ref = R.qevolve_exact(systems[b], time, dt=dt, coll_rates=rates[b])
        close(pops[b], ref['populations'])


def test_trimer_two_excitons_vs_oracle(engine):
    """ESA on a trimer with pseudomodes touches the two-exciton sector; compare with the operator-form oracle."""
    from qsextra_b200 import ChromophoreSystem
    from qsextra_b200.spectroscopy import FeynmanDiagram
    s = ChromophoreSystem(energies=[1.55, 1.46, 1.50], couplings=[[0, -.01, .004], [-.01, 0, -.008], [.004, -.008, 0]],
                          dipole_moments=[1., 0.7, -0.4], frequencies_pseudomode=0.02, levels_pseudomode=2,
                          couplings_ep=0.054)
    dt = 0.5
    spec = FeynmanDiagram('esa', [[0., 2., 1.], [1., 0.], [0., 3., 3.]])
    kw = dict(dt=dt, coll_rates=[0.2])
    close(qspectroscopy(s, spec, shots=0, verbose=False, **kw), R.qspectroscopy_operator(s, spec, **kw))


def test_restart_from_reference_checkpoint(engine, tmp_path, monkeypatch):
    """Wire-compat with an interrupted reference run (qspectroscopy.py:103-111, :185-190): finished grid points are
    kept from signal.npy, the rest is computed, the folder is removed."""
    import os
    from qsextra_b200 import ExcitonicSystem
    from qsextra_b200.spectroscopy import FeynmanDiagram
    monkeypatch.chdir(tmp_path)
    s = ExcitonicSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 1.])
    dt = 0.1 / 0.658212
    spec = FeynmanDiagram('gsb', [[0., 10 * dt], [0.], [0., 5 * dt, 10 * dt]])
    kw = dict(dt=dt, coll_rates=59.08e-3 / 4)
    full = qspectroscopy(s, spec, shots=0, verbose=False, **kw)
    os.mkdir('ckpt')
    partial = np.zeros_like(full)
    partial.reshape(-1)[:4] = 77.0 + 1j                      # what the "interrupted run" had stored for points 0..3
    np.save('ckpt/signal.npy', partial)
    np.save('ckpt/last_time_indices_number.npy', 3)
    resumed = qspectroscopy(s, spec, shots=0, verbose=False, restart_from_checkpoint='ckpt', **kw)
    assert np.all(resumed.reshape(-1)[:4] == 77.0 + 1j)
    close(resumed.reshape(-1)[4:], full.reshape(-1)[4:])
    assert not os.path.exists('ckpt')
    again = qspectroscopy(s, spec, shots=0, verbose=False, checkpoint=True, **kw)
    close(again, full)
    assert [d for d in os.listdir('.') if os.path.isdir(d)] == []


@pytest.mark.parametrize('name', golden_names('qspectroscopy_'))
def test_qspectroscopy_golden_with_the_last_delay_in_the_heisenberg_picture(engine, name, monkeypatch):
    """Same goldens with the read-out propagated backwards for the last delay (engine/response.py), which the big
    grids use; forced here for every grid size and with the register-resident forward path switched off."""
    monkeypatch.setenv('QSX_HEISENBERG_MIN', '1')
    monkeypatch.setenv('QSX_HEISENBERG_ALWAYS', '1')
    monkeypatch.setenv('QSX_DISABLE_RR', '1')
    meta, arrays = load_golden(name)
    system = build_system(meta['system'])
    spec = build_spectroscopy(meta['label'], meta['delay_time'])
    signal = qspectroscopy(system, spec, shots=0, verbose=False, **dict(meta['kwargs']))
    assert signal.shape == arrays['signal'].shape
    close(signal, arrays['signal'], scale=max(1.0, np.abs(arrays['signal']).max()))


def test_heisenberg_and_forward_agree_on_a_chain_with_pseudomodes(engine, monkeypatch):
    from qsextra_b200 import ChromophoreSystem
    from qsextra_b200.spectroscopy import FeynmanDiagram
    J = np.zeros((4, 4))
    for i in range(3):
        J[i, i + 1] = J[i + 1, i] = -0.01 * (1 + 0.3 * i)
    s = ChromophoreSystem(energies=[1.55, 1.50, 1.46, 1.52], couplings=J, dipole_moments=[1., 0.8, -0.6, 1.1],
                          frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=0.054)
    dt = 0.15
    delays = [(dt * 3 * np.arange(6)).tolist(), (dt * 5 * np.arange(3)).tolist(), (dt * 3 * np.arange(7)).tolist()]
    for label in ('esa', 'se', 'gsb'):
        spec = FeynmanDiagram(label, delays)
        backward = qspectroscopy(s, spec, shots=0, verbose=False, dt=dt, coll_rates=[0.2])      # 18 chains >= 8
        monkeypatch.setenv('QSX_DISABLE_HEISENBERG', '1')
        forward = qspectroscopy(s, spec, shots=0, verbose=False, dt=dt, coll_rates=[0.2])
        monkeypatch.delenv('QSX_DISABLE_HEISENBERG')
        close(backward, forward, scale=max(1.0, np.abs(forward).max()))


def test_populations_of_a_chain_with_pseudomodes_use_the_cta_wide_kernel(engine):
    """qevolve / qevolve_batch on the single-exciton block of 3- and 4-site chains with one pseudomode per site: the
    population read-out of the CTA-wide register-resident kernel against the oracle, several parameter sets at once."""
    from qsextra_b200 import ChromophoreSystem
    from qsextra_b200.engine.runtime import get_engine
    rng = np.random.default_rng(21)
    eng = get_engine()
    for n in (3, 4):
        systems, rates = [], []
        for _ in range(3):
            J = np.zeros((n, n))
            for i in range(n - 1):
                J[i, i + 1] = J[i + 1, i] = rng.uniform(-0.05, 0.05)
            s = ChromophoreSystem(energies=rng.uniform(1.4, 1.6, n).tolist(), couplings=J, dipole_moments=[1.] * n,
                                  frequencies_pseudomode=float(rng.uniform(0.01, 0.2)), levels_pseudomode=2,
                                  couplings_ep=float(rng.uniform(0.02, 0.08)))
            s.set_state('localized excitation', int(rng.integers(n)))
            systems.append(s)
            rates.append([float(rng.uniform(0.1, 0.4))])
        dt = 0.15
        time = np.arange(0, 31, 5) * dt
        before = eng.launches
        pops = qevolve_batch(systems, time, dt=dt, coll_rates=np.array(rates))
        assert pops.shape == (3, 7, n) and eng.launches == before + 1          # one launch for the whole batch
        for b in range(3):
            ref = R.qevolve_exact(systems[b], time, dt=dt, coll_rates=rates[b])
            close(pops[b], ref['populations'])


def test_diagrams_of_one_spectrum_share_their_common_levels(engine):
    """``shared=``: gsb, se and esa of the same system and grid computed with one dict equal the independent calls
    (bit for bit: the shared levels are the same launches) and need fewer launches; another system is refused."""
    from qsextra_b200 import ChromophoreSystem
    from qsextra_b200.spectroscopy import FeynmanDiagram
    J = np.zeros((4, 4))
    for i in range(3):
        J[i, i + 1] = J[i + 1, i] = -0.01
    s = ChromophoreSystem(energies=[1.55, 1.50, 1.46, 1.52], couplings=J, dipole_moments=[1., 0.8, -0.6, 1.1],
                          frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=0.054)
    dt = 0.15
    delays = [(dt * 3 * np.arange(6)).tolist(), (dt * 5 * np.arange(3)).tolist(), (dt * 3 * np.arange(7)).tolist()]
    kw = dict(shots=0, verbose=False, dt=dt, coll_rates=[0.2])
    before = engine.launches
    alone = {lab: qspectroscopy(s, FeynmanDiagram(lab, delays), **kw) for lab in ('gsb', 'se', 'esa')}
    launches_alone = engine.launches - before
    shared = {}
    before = engine.launches
    together = {lab: qspectroscopy(s, FeynmanDiagram(lab, delays), shared=shared, **kw) for lab in ('gsb', 'se', 'esa')}
    launches_together = engine.launches - before
    for lab in alone:
        assert np.array_equal(alone[lab], together[lab]), lab
    assert launches_together < launches_alone
    other = ChromophoreSystem(energies=[1.55, 1.50, 1.46, 1.53], couplings=J, dipole_moments=[1., 0.8, -0.6, 1.1],
                              frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=0.054)
    with pytest.raises(ValueError, match='shared dict'):
        qspectroscopy(other, FeynmanDiagram('gsb', delays), shared=shared, **kw)
    # the set API: read-outs of all diagrams enqueued up front
    from qsextra_b200.spectroscopy import qspectroscopy_set
    as_set = qspectroscopy_set(s, [FeynmanDiagram(lab, delays) for lab in ('gsb', 'se', 'esa')], **kw)
    for lab, got in zip(('gsb', 'se', 'esa'), as_set):
        assert np.array_equal(alone[lab], got), lab
    with pytest.raises(ValueError, match='same delay_time'):
        qspectroscopy_set(s, [FeynmanDiagram('gsb', delays), FeynmanDiagram('se', [delays[0], delays[1], delays[2][:-1]])], **kw)


def test_batch_with_different_initial_states_equals_single_runs(engine):
    """qevolve_batch over systems that differ in parameters AND in their (multi-sector) initial states."""
    from qsextra_b200 import ExcitonicSystem
    rng = np.random.default_rng(33)
    systems = []
    for b in range(5):
        j01, j12 = rng.uniform(-.2, .2, 2)
        s = ExcitonicSystem(energies=rng.uniform(1.3, 1.7, 3).tolist(),
                            couplings=[[0, j01, 0], [j01, 0, j12], [0, j12, 0]], dipole_moments=[1., 1., 1.])
        c = rng.normal(size=8) + 1j * rng.normal(size=8)
        s.set_state('state', (c / np.linalg.norm(c)).tolist())
        systems.append(s)
    dt = 0.2
    time = np.arange(0, 9, 2) * dt
    rates = rng.uniform(0.05, 0.3, 5)
    pops = qevolve_batch(systems, time, dt=dt, coll_rates=rates)
    assert pops.shape == (5, 5, 3)
    for b, s in enumerate(systems):
        single = qevolve(s, time, shots=0, dt=dt, coll_rates=float(rates[b]), verbose=False)
        close(pops[b], single.populations)
        ref = R.qevolve_exact(s, time, dt=dt, coll_rates=float(rates[b]))
        close(single.populations, ref['populations'])
    # a batch that never leaves the ground state
    for s in systems:
        s.set_state('ground')
    assert not np.any(qevolve_batch(systems, time, dt=dt, coll_rates=rates))


def test_shots_add_the_estimator_sampling_error(engine):
    """F2 (qspectroscopy.py:114, :180-181): shots > 0 returns exact signal + sampling error of the stated sigma,
    reproducible with a seed, on the host and on the device; shots = 0 / None stay exact."""
    from qsextra_b200 import ExcitonicSystem
    from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy_set
    from qsextra_b200.spectroscopy.simulation.qspectroscopy import estimator_noise_sigma
    s = ExcitonicSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 0.8])
    dt = 0.1 / 0.658212
    grid = (dt * np.arange(40)).tolist()
    spec = FeynmanDiagram('gsb', [grid, [0.], grid])
    kw = dict(verbose=False, dt=dt, coll_rates=59.08e-3 / 4)
    exact = qspectroscopy(s, spec, shots=0, **kw)
    assert np.array_equal(exact, qspectroscopy(s, spec, shots=None, **kw))
    noisy = qspectroscopy(s, spec, shots=4000, seed=3, **kw)
    sigma = estimator_noise_sigma(spec.side_seq, spec.direction_seq, [1., 0.8], 4000)
    dev = noisy - exact
    assert abs(dev.real.std() / sigma - 1) < 0.1 and abs(dev.imag.std() / sigma - 1) < 0.1
    assert np.array_equal(noisy, qspectroscopy(s, spec, shots=4000, seed=3, **kw))
    on_device = qspectroscopy(s, spec, shots=4000, seed=3, device_output=True, **kw)
    dev2 = on_device.cpu().numpy() - exact
    assert abs(dev2.real.std() / sigma - 1) < 0.1
    default = qspectroscopy(s, spec, **kw)                    # the reference's default: shots = 1000
    assert abs((default - exact).real.std() / (sigma * 2) - 1) < 0.1
    both = qspectroscopy_set(s, [spec], **kw)                 # the set API defaults to exact
    assert np.array_equal(both[0], exact)


def test_basis_substitution_on_the_gpu(engine, monkeypatch):
    """4-site chain with pseudomodes, 80 t1 values: the t2 level propagates the 64 basis elements of the (0,1) block's
    support instead of 80 chains (engine/support.py); same signal as the chain-by-chain tree and as the oracle rows."""
    from qsextra_b200 import ChromophoreSystem
    from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy_set
    from oracle import fast
    J = np.zeros((4, 4))
    for i in range(3):
        J[i, i + 1] = J[i + 1, i] = -0.01
    s = ChromophoreSystem(energies=[1.55, 1.50, 1.46, 1.52], couplings=J, dipole_moments=[1., 0.8, -0.6, 1.1],
                          frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=0.054)
    dt = 0.15
    delays = [(dt * 2 * np.arange(80)).tolist(), (dt * 7 * np.arange(3)).tolist(), (dt * 3 * np.arange(9)).tolist()]
    specs = [FeynmanDiagram(lab, delays) for lab in ('gsb', 'se', 'esa')]
    kw = dict(verbose=False, dt=dt, coll_rates=[0.2])
    got = qspectroscopy_set(s, specs, **kw)
    monkeypatch.setenv('QSX_DISABLE_BASIS', '1')
    plain = qspectroscopy_set(s, specs, **kw)
    rows = [(79, 2), (40, 1), (0, 0)]
    for spec, a, b in zip(specs, got, plain):
        scale = np.abs(b).max()
        assert np.abs(a - b).max() < 1e-11 * scale
        ref, _ = fast.response_points(s, spec, rows, dt, coll_rates=[0.2], kraus=True)
        assert np.abs(np.array([a[i, j] for i, j in rows]) - ref).max() < 1e-10 * scale


def test_qsx_evolve_driven_by_the_c_program_builder(engine):
    """A caller that is not Python: the step program comes from qsx_step_program_build (C, host side) instead of
    engine/program.py, goes to qsx_evolve as it is, and the populations equal the oracle's."""
    import torch
    from test_cabi import c_built_program
    from qsextra_b200 import ChromophoreSystem
    from qsextra_b200.engine.sectors import population_observables
    s = ChromophoreSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 1.], frequencies_pseudomode=0.02,
                          levels_pseudomode=2, couplings_ep=0.054)
    s.set_state('localized excitation', 0)
    dt = 0.1 / 0.658212
    steps = np.arange(0, 61, 10)
    want = R.qevolve_exact(s, steps * dt, dt=dt, coll_rates=[0.2])['populations']
    from qsextra_b200.qcomo.evolution.qevolve import _program_for, _validate_rates
    blk = _program_for(s, dt, dt, 1, 1, _validate_rates(s, [0.2])).block(1, 1)
    ops, params, cfg_a, cfg_b, n_bits = c_built_program(s, 1, 1, dt, [0.2], 1)
    dev = dict(block=engine.device_block(blk), ops=engine.to_device(ops, torch.int32), params=engine.to_device(params[None, :]),
               n_ops=len(ops), stride=len(params), n_sets=1, n_passes=0, team=0, bound=None)
    sigma0 = np.zeros((1, blk.size), dtype=complex)
    sigma0[0, 0] = 1.0                                   # |site 0 excited, modes 0><same|
    obs = engine.upload_observables(*population_observables(blk), kind='populations')
    got, _ = engine.evolve(dev, engine.to_device(sigma0), 1, int(steps.max()), steps, obs=obs, obs_real=True)
    assert np.abs(got.cpu().numpy()[0] - want).max() < 1e-10


def test_ensemble_signals_copied_to_pinned_host_memory_on_the_way(engine):
    """``qspectroscopy_ensemble(host_out=...)``: every diagram's rows leave for the pinned host tensors as soon as the
    diagram is complete (ascending delays), or from the finished signal (unsorted delays: the rows are not the caller's
    grid yet).  Both must equal the returned signals; eager evaluation, captured plan and replay alike."""
    import torch
    from qsextra_b200 import ChromophoreSystem
    from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy_ensemble
    rng = np.random.default_rng(5)
    members = [ChromophoreSystem(energies=(np.array([1.55, 1.50, 1.46]) + 0.01 * rng.normal(size=3)).tolist(),
                                 couplings=[[0, -.01, 0], [-.01, 0, -.01], [0, -.01, 0]], dipole_moments=[1., 0.8, 1.1],
                                 frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=0.054) for _ in range(5)]
    dt = 0.15
    kw = dict(verbose=False, dt=dt, coll_rates=[0.2])
    for t2 in ([0., 3 * dt], [3 * dt, 0.]):
        delays = [(dt * np.arange(12)).tolist(), t2, (2 * dt * np.arange(9)).tolist()]
        specs = [FeynmanDiagram(lab, delays) for lab in ('gsb', 'se', 'esa')]
        for _ in range(3):
            pinned = [torch.zeros((5, 12, 2, 9), dtype=torch.complex128).pin_memory() for _ in specs]
            signals = qspectroscopy_ensemble(members, specs, host_out=pinned, **kw)
            for got, want in zip(pinned, signals):
                assert np.abs(want).max() > 0 and np.array_equal(got.numpy(), want)
