"""Host-side logic on the CPU: the step-program compiler (masks, lowering, pass regrouping, sector tables) checked
against the reference goldens through the NumPy emulation of the engine's operation semantics (tests/emulator.py)."""
import numpy as np
import pytest

import emulator
from helpers import build_spectroscopy, build_system, golden_names, load_golden, time_arg
from qsextra_b200.engine.program import StepProgram, SystemBatch, mode_operator_terms, resolve_time_grid
from qsextra_b200.engine.response import interaction_plan
from qsextra_b200.engine.sectors import Block, sector_configs, sector_rank, split_state_by_sector
from qsextra_b200.qcomo.evolution.qevolve import _program_for, initial_amplitudes
from qsextra_b200.spectroscopy.simulation.qspectroscopy import delay_step_counts


def _program(system, kw, time=None):
    dt, dtt, tn, counts = resolve_time_grid(time if time is not None else kw['dt'], kw.get('dt'), kw.get('Trotter_number', 1))
    return _program_for(system, dt, dtt, tn, kw.get('Trotter_order', 1), kw.get('coll_rates')), counts


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('name', golden_names('qevolve_'))
def test_step_program_reproduces_reference_distribution(name, fused):
    meta, arrays = load_golden(name)
    system = build_system(meta['system'])
    kw = dict(meta['kwargs'])
    t = kw.pop('time')
    prog, counts = _program(system, kw, time_arg(t))
    n = system.system_size
    wanted = sorted(set(int(c) for c in counts))
    dist = np.zeros((len(wanted), 1 << n))
    for k, amp in split_state_by_sector(n, initial_amplitudes(system)).items():
        blk = prog.block(k, k)
        bound = prog.bind(blk, fused=fused)
        sig = np.zeros((blk.dim_a, blk.dim_b), complex)
        at = np.arange(blk.n_a) << blk.n_bits
        sig[np.ix_(at, at)] = np.outer(amp, amp.conj())
        step = emulator.apply_pass_program if fused and bound.passes is not None else emulator.apply_step
        for s in range(max(wanted) + 1):
            if s:
                sig = step(sig, bound)
            if s in wanted:
                dist[wanted.index(s), list(sector_configs(n, k))] += np.real(np.diag(sig)).reshape(blk.n_a, -1).sum(1)
    assert np.abs(dist - arrays['distribution']).max() < 1e-12


@pytest.mark.parametrize('name', golden_names('qspectroscopy_'))
def test_response_tree_reproduces_reference_signal(name):
    meta, arrays = load_golden(name)
    system = build_system(meta['system'])
    spec = build_spectroscopy(meta['label'], meta['delay_time'])
    kw = dict(meta['kwargs'])
    if kw.get('dt') is None:
        pytest.skip('dt=None: one step program per delay, no shared tree (checked on the GPU against the same golden)')
    counts, (dt, dtt, tn, _) = delay_step_counts(spec.delay_time, kw.get('dt'), kw.get('Trotter_number', 1))
    prog = _program_for(system, dt, dtt, tn, kw.get('Trotter_order', 1), kw.get('coll_rates'))
    got = emulator.response_function(prog, system.dipole_moments, spec.side_seq, spec.direction_seq, counts)
    assert np.abs(got - arrays['signal']).max() < 1e-12 * max(1.0, np.abs(arrays['signal']).max())


def test_mode_operator_masks_match_reference_dictionaries():
    """Same Pauli strings, same order, same coefficients as qevolve.py:107-195 (fixture from the reference)."""
    meta, _ = load_golden('mode_operators')

    def label(x, z, nq):
        return ''.join('IXZY'[((x >> q) & 1) | (((z >> q) & 1) << 1)] for q in range(nq - 1, -1, -1))
    for d, dicts in meta.items():
        got = mode_operator_terms(int(d))
        nq = len(dicts[0][0][0])
        for want, have in zip(dicts, got):
            assert [k for k, _ in want] == [label(x, z, nq) for (x, z) in have]
            for (_, (re, im)), v in zip(want, have.values()):
                assert abs(complex(re, im) - v) < 1e-15


def test_pass_program_equals_operation_list_on_random_blocks():
    rng = np.random.default_rng(5)
    sd = dict(energies=[1.55, 1.50, 1.46, 1.52], couplings=[[0, -.01, 0, 0], [-.01, 0, -.01, 0], [0, -.01, 0, -.01], [0, 0, -.01, 0]],
              dipole_moments=[1.] * 4, frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=0.054)
    system = build_system(sd)
    prog = _program_for(system, 0.15, 0.15, 1, 1, [0.2])
    for ka, kb in ((0, 1), (1, 1), (2, 1), (1, 0)):
        blk = prog.block(ka, kb)
        for width in (1, 2):
            bound = prog.bind(blk, kb=width)
            assert {int(p[0]) for p in bound.passes} == {0, 1}            # fully fused: one cfg pass + mode passes
            assert len(bound.passes) == 1 + (4 if width == 1 else 2)
            sig = rng.normal(size=(blk.dim_a, blk.dim_b)) + 1j * rng.normal(size=(blk.dim_a, blk.dim_b))
            assert np.abs(emulator.apply_step(sig, bound) - emulator.apply_pass_program(sig, bound)).max() < 1e-13


def test_batched_parameters_equal_single_compiles():
    rng = np.random.default_rng(2)
    systems = []
    for _ in range(3):
        s = build_system(dict(energies=rng.uniform(1.4, 1.6, 2).tolist(), couplings=float(rng.uniform(-.05, .05)),
                              dipole_moments=[1., 1.], frequencies_pseudomode=float(rng.uniform(.01, .2)),
                              levels_pseudomode=2, couplings_ep=float(rng.uniform(.01, .1))))
        systems.append(s)
    rates = rng.uniform(0.1, 0.4, (3, 1))
    batch_prog = StepProgram(SystemBatch.from_systems(systems), 0.15, 1, 1, rates)
    blk = batch_prog.block(1, 1)
    together = batch_prog.bind(blk)
    for b, s in enumerate(systems):
        alone = _program_for(s, 0.15, 0.15, 1, 1, rates[b].tolist()).bind(blk)
        assert np.array_equal(alone.ops, together.ops)
        assert np.abs(alone.params[0] - together.params[b]).max() < 1e-15


def test_time_grid_rules():
    dt, dtt, tn, counts = resolve_time_grid([0., 0.3, 0.1, 0.3], 0.1, 2)
    assert (dt, dtt, tn) == (0.1, 0.05, 2) and counts.tolist() == [0, 3, 1, 3]
    dt, dtt, tn, counts = resolve_time_grid(np.arange(0, 1.01, 0.5), None, 5)      # dt from the grid, collisions every dt/5
    assert abs(dt - 0.1) < 1e-15 and tn == 1 and counts.tolist() == [0, 5, 10]
    with pytest.raises(ValueError):
        resolve_time_grid([0., 0.25], 0.1, 1)


def test_sector_tables_and_interaction_plan():
    assert sector_configs(4, 2) == (3, 5, 6, 9, 10, 12)
    rank = sector_rank(4, 2)
    assert rank[6] == 2 and rank[7] == -1
    blk = Block(4, 4, 2, 1)
    assert (blk.dim_a, blk.dim_b, blk.size) == (96, 64, 6144)
    assert interaction_plan('bbkk', 'ioio') == [('b', 'X'), ('b', 'raise'), ('k', 'X'), ('k', 'X')]
    assert interaction_plan('bkkk', 'iiio') == [('b', 'X'), ('k', 'X'), ('k', 'raise'), ('k', 'X')]
    assert interaction_plan('kk', 'io') == [('k', 'X'), ('k', 'X')]
    ptr, idx, w = Block(2, 0, 1, 0).emission_observable([2.0, 3.0])
    assert idx.tolist() == [0, 1] and w.tolist() == [2.0, 3.0]


def _chain(n, **kw):
    J = np.zeros((n, n))
    for i in range(n - 1):
        J[i, i + 1] = J[i + 1, i] = kw.get('J', -0.3)
    return build_system(dict(energies=(1.5 + 0.01 * np.arange(n)).tolist(), couplings=J.tolist(), dipole_moments=[1.] * n,
                             frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=kw.get('g', 0.9)))


def test_register_resident_programs_equal_operation_list():
    """engine/rr.py (small power-of-two blocks) and engine/rrc.py (CTA-wide, large blocks) against the plain operation
    list, through the NumPy emulations of their kernels; strong couplings so that a wrong cosine fold shows up."""
    from qsextra_b200.engine import rr as RR, rrc as RRC
    rng = np.random.default_rng(11)
    dimer = build_system(dict(energies=[1.55, 1.46], couplings=-0.4, dipole_moments=[1., 0.8], frequencies_pseudomode=0.3,
                              levels_pseudomode=2, couplings_ep=0.8))
    prog = _program_for(dimer, 0.15, 0.15, 1, 1, [0.5])
    kinds = set()
    for ka, kb in ((1, 1), (1, 0), (0, 1), (2, 1), (1, 2), (0, 0)):
        blk = prog.block(ka, kb)
        bound = prog.bind(blk, fused=False)
        for rb in (None, 2, 3, 4):
            rrp = RR.build(prog, bound, reg_bits=rb)
            assert rrp is not None
            kinds |= {int(o[0]) for o in rrp.ops}
            sig = rng.normal(size=(blk.dim_a, blk.dim_b)) + 1j * rng.normal(size=(blk.dim_a, blk.dim_b))
            assert np.abs(emulator.apply_step(sig, bound) - emulator.apply_rr_program(sig, rrp)).max() < 1e-13
    assert RR.RR_ROTT in kinds                       # the dimer folds its cosines (uniform pending factors)
    # N = 4 without pseudomodes: two configuration bits, hops pair configurations with different pending factors
    four = build_system(dict(energies=[1., 2., 3., 4.], couplings=[[0, 1, 0, 1], [1, 0, 1, 0], [0, 1, 0, 1], [1, 0, 1, 0]],
                             dipole_moments=[1.] * 4))
    prog4 = _program_for(four, 0.1, 0.1, 1, 1, 0.1)
    for ka, kb in ((1, 0), (0, 1), (1, 1)):
        blk = prog4.block(ka, kb)
        bound = prog4.bind(blk, fused=False)
        rrp = RR.build(prog4, bound)
        if rrp is None:
            continue
        sig = rng.normal(size=(blk.dim_a, blk.dim_b)) + 1j * rng.normal(size=(blk.dim_a, blk.dim_b))
        assert np.abs(emulator.apply_step(sig, bound) - emulator.apply_rr_program(sig, rrp)).max() < 1e-13
    for n in (3, 4):
        progn = _program_for(_chain(n), 0.15, 0.15, 1, 1, [0.2])
        for ka, kb in ((2, 1), (1, 1), (1, 0), (0, 1), (0, 0), (1, 2)):
            blk = progn.block(ka, kb)
            bound = progn.bind(blk, fused=False)
            big = RRC.build(progn, bound)
            assert big is not None and big.threads == 4 ** n
            sig = rng.normal(size=(blk.dim_a, blk.dim_b)) + 1j * rng.normal(size=(blk.dim_a, blk.dim_b))
            assert np.abs(emulator.apply_step(sig, bound) - emulator.apply_rrc_program(sig, big)).max() < 1e-13
            back = RRC.build(progn, bound.transposed())        # the read-out propagated backwards
            assert back is not None and [o[0] for o in back.ops].count(RRC.OP_ADT) == n and back.ops[-1][0] == RRC.OP_DIAG
            assert np.abs(emulator.apply_step(sig, bound.transposed()) - emulator.apply_rrc_program(sig, back)).max() < 1e-13


def test_generated_program_tables_are_current():
    """csrc/*.inc are generated from the host compiler; regenerate in memory and compare (a stale table would silently
    send launches to the generic kernels)."""
    import importlib.util, os, io, contextlib
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for tool, inc in (('gen_rr_programs.py', 'qsx_rr_programs.inc'), ('gen_rrc_programs.py', 'qsx_rrc_programs.inc')):
        path = os.path.join(repo, 'qsextra_b200', 'csrc', inc)
        before = open(path).read()
        spec = importlib.util.spec_from_file_location('gen_' + inc, os.path.join(repo, 'tools', tool))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        with contextlib.redirect_stdout(io.StringIO()):
            mod.main()
        after = open(path).read()
        assert before == after, inc + ' is stale: run python tools/' + tool


@pytest.mark.parametrize('sysdef,kw,blocks', [
    (dict(energies=[1.5, 1.4], couplings=-0.1, dipole_moments=[1., 1.], frequencies_pseudomode=[0.2], levels_pseudomode=[2],
          couplings_ep=[0.3]), dict(coll_rates=[0.4]), [(1, 1), (2, 1), (1, 0)]),
    (dict(energies=[1.5, 1.4, 1.6], couplings=[[0, .1, .05], [.1, 0, .2], [.05, .2, 0]], dipole_moments=[1.] * 3),
     dict(coll_rates=0.2, Trotter_order=2), [(1, 1), (2, 1)]),
    (dict(energies=[1.5, 1.4], couplings=-0.1, dipole_moments=[1., 1.], frequencies_pseudomode=[0.2], levels_pseudomode=[3],
          couplings_ep=[0.3]), dict(coll_rates=[0.4]), [(1, 1), (1, 0)]),           # Y rotations and resets
])
def test_transposed_program_propagates_the_read_out(sysdef, kw, blocks):
    """<W, Phi^n sigma> = <(Phi^T)^n W, sigma> with the transposed program (reverse order, QSX_OPF_TRANSPOSE)."""
    from qsextra_b200.qcomo.evolution.qevolve import _program_for
    system = build_system(sysdef)
    tn = kw.get('Trotter_number', 1)
    prog = _program_for(system, 0.2, 0.2 / tn, tn, kw.get('Trotter_order', 1), kw.get('coll_rates'))
    rng = np.random.default_rng(8)
    for ka, kb in blocks:
        blk = prog.block(ka, kb)
        fwd = prog.bind(blk, fused=False)
        bwd = fwd.transposed()
        assert bwd.passes is None and np.all(bwd.ops[:, 7] == 1) and np.array_equal(bwd.ops[::-1, :7], fwd.ops[:, :7])
        sigma = rng.normal(size=(blk.dim_a, blk.dim_b)) + 1j * rng.normal(size=(blk.dim_a, blk.dim_b))
        w = rng.normal(size=(blk.dim_a, blk.dim_b)) + 1j * rng.normal(size=(blk.dim_a, blk.dim_b))
        s3, w3 = sigma, w
        for _ in range(3):
            s3 = emulator.apply_step(s3, fwd)
            w3 = emulator.apply_step(w3, bwd)
        assert abs(np.sum(w * s3) - np.sum(w3 * sigma)) < 1e-12 * np.abs(w).sum()


def test_cta_wide_programs_refuse_collisions_that_empty_a_pseudomode():
    """The damping-diagonal forms of the CTA-wide kernel divide by the product of the damping cosines at emissions
    (csrc/qsx_rrc3.cuh, qsx_rrc3t.cuh): engine/rrc.py hands blocks whose collision (almost) empties a pseudomode in one
    step to the general kernel instead (cos phi < 1e-3), and keeps ordinary rates."""
    from qsextra_b200.engine import rrc as RRC
    from qsextra_b200.engine.program import StepProgram, SystemBatch
    J = np.zeros((1, 4, 4))
    for i in range(3):
        J[0, i, i + 1] = J[0, i + 1, i] = -0.3
    batch = SystemBatch(np.full((1, 4), 1.5), J, np.ones((1, 4)), np.full((1, 1), 0.3), np.full((1, 1), 0.7), (2,))
    # (the collision angle is sqrt(rate dt): pi / 2 moves the whole excitation of the pseudomode out in one step)
    for rate, served in ((0.5, True), (30.0, True), ((np.pi / 2) ** 2 / 0.15, False)):
        prog = StepProgram(batch, 0.15, 1, 1, np.full((1, 1), rate))
        for ka, kb in ((1, 1), (0, 1)):
            bound = prog.bind(prog.block(ka, kb), fused=False)
            assert (RRC.build(prog, bound) is not None) == served, (rate, ka, kb)
            assert (RRC.build(prog, bound.transposed()) is not None) == served, (rate, ka, kb)
