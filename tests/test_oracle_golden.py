"""Pin the oracle: its restatement of the reference's circuit construction must reproduce, to rounding, the
fixtures written by the UNMODIFIED reference running over oracle/refshim (tests/golden/make_golden.py), plus the
notebook known-answers listed in SURVEY.md section 4."""

import numpy as np
import pytest

from helpers import build_spectroscopy, build_system, golden_names, load_golden, time_arg
from oracle import circuits as C
from oracle import fast, reference_path as R
from oracle.densesim import DensityMatrix, pauli_matrix


@pytest.mark.parametrize('name', golden_names('qevolve_'))
def test_qevolve_oracle_matches_reference(name):
    meta, arrays = load_golden(name)
    system = build_system(meta['system'])
    kw = dict(meta['kwargs'])
    t = kw.pop('time')
    out = R.qevolve_exact(system, time_arg(t), **kw)
    assert np.abs(out['distribution'] - arrays['distribution']).max() < 1e-13
    assert np.abs(out['trace'] - 1).max() < 1e-12


@pytest.mark.parametrize('name', golden_names('qspectroscopy_'))
def test_qspectroscopy_operator_form_matches_reference(name):
    meta, arrays = load_golden(name)
    system = build_system(meta['system'])
    spec = build_spectroscopy(meta['label'], meta['delay_time'])
    signal = R.qspectroscopy_operator(system, spec, **dict(meta['kwargs']))
    assert np.abs(signal - arrays['signal']).max() < 1e-12 * max(1.0, np.abs(arrays['signal']).max())


@pytest.mark.parametrize('name', ['qspectroscopy_abs_dimer_markov', 'qspectroscopy_custom_kbbk_dimer'])
def test_qspectroscopy_literal_form_matches_reference(name):
    """The circuit-by-circuit restatement (every site pathway, X/Y variants, ketbra ancilla)."""
    meta, arrays = load_golden(name)
    system = build_system(meta['system'])
    delays = meta['delay_time']
    if name.startswith('qspectroscopy_abs'):
        delays = delays[:3]
        want = arrays['signal'][:3]
    else:
        want = arrays['signal']
    spec = build_spectroscopy(meta['label'], delays)
    signal = R.qspectroscopy_literal(system, spec, **dict(meta['kwargs']))
    assert np.abs(signal - want).max() < 1e-12 * max(1.0, np.abs(want).max())


def test_circuits_per_pathway_counts():
    """Flag logic of qspectroscopy.py:142-176: 1 ('a'), 3 ('gsb'), 2 ('se'), 2 ('esa') circuits per site pathway."""
    system = build_system(dict(energies=[1., 1.1], couplings=0.1, dipole_moments=[1., 1.]))
    for label, n in (('a', 1), ('gsb', 3), ('se', 2), ('esa', 2)):
        spec = build_spectroscopy(label, [0.] if label == 'a' else [[0.], [0.], [0.]])
        path = (0,) * len(spec.side_seq)
        circuits = C.pathway_circuits(system, spec.side_seq, spec.direction_seq, [[]] * len(spec.delay_time), path)
        assert len(circuits) == n


def test_mode_operator_dictionaries_match_reference():
    meta, _ = load_golden('mode_operators')
    for d, dicts in meta.items():
        got = C.mode_operator_terms(int(d))
        for want, have in zip(dicts, got):
            assert [k for k, _ in want] == list(have.keys())
            for (_, (re, im)), v in zip(want, have.values()):
                assert abs(complex(re, im) - v) < 1e-15


def test_known_answers_from_notebooks():
    # K-abs (Absorption notebooks): signal(0) = sum mu^2 = 2, signal(10 fs) ~ -0.763 + 0.634i
    meta, arrays = load_golden('qspectroscopy_abs_dimer_markov')
    assert abs(arrays['signal'][0] - 2.0) < 1e-12
    assert abs(arrays['signal'][1] - (-0.763 + 0.634j)) < 2e-3
    assert np.abs(arrays['signal'] - arrays['signal_cl']).max() < 2e-3         # q -> cl gap is O(dt)
    # K-plots (Example 2): P_0(t = 1, 2, 3) = 0.428, 0.528, 0.641
    meta, arrays = load_golden('qevolve_ex2_pseudomode')
    p0 = arrays['distribution'][:, 1]
    for t, want in ((1., 0.428), (2., 0.528), (3., 0.641)):
        k = int(round(t / 0.5))
        assert abs(p0[k] - want) < 1.5e-3
    _, cl = load_golden('clevolve_ex2_pseudomode')
    assert np.abs(cl['populations'][:, 0] - p0).max() < 5e-4                    # collision model vs Lindblad
    # Example 1: dephased populations of the N = 4 network stay normalised and in [0, 1]
    _, arrays = load_golden('qevolve_ex1_dephasing')
    assert np.abs(arrays['distribution'].sum(1) - 1).max() < 1e-12
    assert arrays['distribution'].min() > -1e-14


def test_densesim_rotation_against_dense_matrices():
    rng = np.random.default_rng(3)
    n = 3
    dm = DensityMatrix(n)
    psi = rng.normal(size=8) + 1j * rng.normal(size=8)
    psi /= np.linalg.norm(psi)
    dm.rho = np.outer(psi, psi.conj())
    rho = dm.rho.copy()
    for label in ('XYZ', 'IZI', 'YIY', 'XXI', 'ZZZ'):
        theta = rng.normal()
        dm.pauli_rotation(label, theta)
        P = pauli_matrix(label)
        G = np.cos(theta) * np.eye(8) - 1j * np.sin(theta) * P
        rho = G @ rho @ G.conj().T
        assert np.abs(dm.rho - rho).max() < 1e-14
        assert abs(dm.expectation_pauli(label) - np.trace(P @ rho)) < 1e-14
    dm.reset(1)
    t = rho.reshape([2] * 6)
    red = np.einsum('abcdbf->acdf', t)
    want = np.zeros_like(t)
    want[:, 0, :, :, 0, :] = red
    assert np.abs(dm.rho - want.reshape(8, 8)).max() < 1e-14
    dm.cx(0, 2)
    dm.cy(2, 1)
    dm.h(0)
    assert abs(dm.trace() - 1) < 1e-13


def test_c_oracle_matches_numpy_oracle():
    meta, _ = load_golden('qevolve_dimer_two_modes')
    system = build_system(meta['system'])
    kw = dict(meta['kwargs'])
    t = kw.pop('time')
    want = R.qevolve_exact(system, t, **kw)['populations']
    got, threads = fast.populations_batch([system, system], t, **kw)
    assert threads >= 1
    assert np.abs(got[0] - want).max() < 1e-13 and np.abs(got[1] - want).max() < 1e-13


def test_collision_model_converges_to_lindblad_with_dt():
    """q -> cl gap shrinks ~ linearly with dt (sanity band of SURVEY 8(c), not a parity bar)."""
    _, cl = load_golden('clevolve_ex1_dephasing')
    meta, _ = load_golden('qevolve_ex1_dephasing')
    sd = dict(meta['system'])
    sd['state'] = ['localized excitation', 0]
    system = build_system(sd)
    t = np.arange(0, 2.01, 0.1)
    gaps = []
    for dt in (0.02, 0.01):
        pops = R.qevolve_exact(system, t, dt=dt, coll_rates=0.1)['populations']
        gaps.append(np.abs(pops - cl['populations']).max())
    assert 0.4 * gaps[0] < gaps[1] < 0.6 * gaps[0] and gaps[1] < 5e-3        # first order in dt


def _golden_response_in_c(name, kraus):
    """Signal of a qspectroscopy golden through the C restatement (oracle/csrc: qsxo_response_chains)."""
    import itertools
    meta, arrays = load_golden(name)
    kw = dict(meta['kwargs'])
    if 'dt' not in kw:
        pytest.skip('the C port evolves with one step size (dt given)')
    if kw.get('Trotter_order', 1) != 1 and kraus:
        pytest.skip('Suzuki splitting of the collision gates has no two-gate pattern')
    system = build_system(meta['system'])
    spec = build_spectroscopy(meta['label'], meta['delay_time'])
    shape = [len(d) for d in spec.delay_time]
    rows = list(itertools.product(*[range(n) for n in shape[:-1]]))
    dt = kw.pop('dt')
    got, _ = fast.response_points(system, spec, rows, dt, kraus=kraus, **kw)
    return got.reshape(shape), arrays['signal']


@pytest.mark.parametrize('name', golden_names('qspectroscopy_'))
def test_c_response_chains_match_reference_goldens(name):
    """The C port used for far delay-grid points (bench.py cpu_baseline, tests/test_gpu_full_size.py) reproduces every
    spectroscopy fixture written by the unmodified reference, with literal collision gates + ancilla reset."""
    got, want = _golden_response_in_c(name, kraus=False)
    assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize('name', golden_names('qspectroscopy_'))
def test_c_response_chains_kraus_form_match_reference_goldens(name):
    """... and with every collision replaced by the channel it induces once the ancilla is traced out (state without
    the ancilla qubit: what makes the 5 000-step corners of BASELINE configs[3] affordable on the CPU)."""
    got, want = _golden_response_in_c(name, kraus=True)
    assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())


def test_kraus_form_rewrites_only_collision_patterns():
    system = build_system(dict(energies=[1., 1.1], couplings=0.1, dipole_moments=[1., 1.], frequencies_pseudomode=[0.1],
                               levels_pseudomode=[4], couplings_ep=[0.05]))
    layout = C.Layout(system, True)
    ops = C.step_ops(system, layout, 0.1, 0.1, 1, 1, [0.2])
    assert fast.kraus_form(ops, layout) is None                 # d = 4 collisions are not the two-gate pattern
    dimer = build_system(dict(energies=[1., 1.1], couplings=0.1, dipole_moments=[1., 1.]))
    layout = C.Layout(dimer, True)
    ops = fast.kraus_form(C.step_ops(dimer, layout, 0.1, 0.1, 1, 1, 0.3), layout)
    assert [op[0] for op in ops[-2:]] == ['dephase', 'dephase']
    assert all(len(op[1]) == layout.n_qubits - 1 for op in ops if op[0] == 'rot')
    assert abs(ops[-1][2] - np.sqrt(0.3 * 0.1)) < 1e-15


def test_sector_oracle_equals_full_space_simulation():
    """oracle/sector.py (the literal gate list kept inside one exciton-number sector; used for the 8-site BASELINE
    configs[4], whose full space no CPU can hold) against the full-space dense simulator on 3 sites."""
    from oracle import sector
    J = np.array([[0, -0.01, 0.005], [-0.01, 0, 0.02], [0.005, 0.02, 0]])
    chrom = build_system(dict(energies=[1.55, 1.50, 1.46], couplings=J.tolist(), dipole_moments=[1., 1., 1.],
                              frequencies_pseudomode=[0.02], levels_pseudomode=[2], couplings_ep=[0.054],
                              state=['delocalized excitation', [0.6, [0.0, 0.8], 0.0]]))
    steps = [0, 1, 2, 7, 20]
    dt = 0.15
    ref = R.qevolve_exact(chrom, np.array(steps) * dt, dt=dt, coll_rates=[0.2])['populations']
    assert np.abs(sector.populations(chrom, steps, dt, coll_rates=[0.2]) - ref).max() < 1e-12
    exc = build_system(dict(energies=[1.55, 1.50, 1.46], couplings=J.tolist(), dipole_moments=[1., 1., 1.],
                            state=['localized excitation', 1]))
    ref = R.qevolve_exact(exc, np.array(steps) * dt, dt=dt, coll_rates=0.05, Trotter_number=2)['populations']
    assert np.abs(sector.populations(exc, steps, dt, coll_rates=0.05, Trotter_number=2) - ref).max() < 1e-12
