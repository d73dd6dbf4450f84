"""Differential test of every kernel family through the C ABI: random complex (non-physical) blocks are evolved a few
steps by (a) the operation list, (b) the pass program, (c) the generic register-resident kernel, (d) its compile-time
variant and (e) the CTA-wide register-resident kernel, and each result is compared element by element with the NumPy
emulation of the operation semantics (tests/emulator.py), which the CPU suite ties to the reference goldens."""
import os
import zlib

import numpy as np
import pytest
import torch

import emulator
from helpers import build_system
from qsextra_b200.qcomo.evolution.qevolve import _program_for

pytestmark = pytest.mark.gpu
STEPS = 3


def chain(n, modes=True, levels=2, w=1, J=-0.3, g=0.7):
    coup = np.zeros((n, n))
    for i in range(n - 1):
        coup[i, i + 1] = coup[i + 1, i] = J * (1 + 0.1 * i)
    sd = dict(energies=(1.5 + 0.07 * np.arange(n)).tolist(), couplings=coup.tolist(), dipole_moments=[1.] * n)
    if modes:
        sd.update(frequencies_pseudomode=[0.3 + 0.1 * k for k in range(w)], levels_pseudomode=[levels] * w,
                  couplings_ep=[g * (1 + 0.2 * k) for k in range(w)])
    return build_system(sd)


def ring(n):
    """All-to-all coupled aggregate with one 2-level pseudomode per site."""
    coup = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            coup[i, j] = coup[j, i] = -0.2 / (1 + j - i)
    return build_system(dict(energies=(1.5 + 0.07 * np.arange(n)).tolist(), couplings=coup.tolist(), dipole_moments=[1.] * n,
                             frequencies_pseudomode=[0.3], levels_pseudomode=[2], couplings_ep=[0.7]))


CASES = [
    ('dimer_modes', chain(2), dict(coll_rates=[0.5]), [(1, 1), (1, 0), (2, 1), (0, 0)]),
    ('dimer_markov', chain(2, modes=False), dict(coll_rates=0.2), [(1, 1), (1, 0), (2, 1)]),
    ('dimer_two_modes', chain(2, w=2), dict(coll_rates=[0.5, 0.2]), [(1, 1), (0, 1)]),
    ('dimer_d3', chain(2, levels=3), dict(coll_rates=[0.4]), [(1, 1), (1, 0)]),
    ('trimer_modes', chain(3), dict(coll_rates=[0.5]), [(1, 1), (2, 1), (1, 0)]),
    ('trimer_markov_trotter2', chain(3, modes=False), dict(coll_rates=0.2, Trotter_number=2), [(1, 1), (2, 1)]),
    ('chain4_modes', chain(4), dict(coll_rates=[0.5]), [(2, 1), (1, 1), (0, 1)]),
    ('chain4_suzuki2', chain(4, modes=False), dict(coll_rates=0.1, Trotter_order=2), [(1, 1), (2, 2)]),
    # couplings outside the compile-time menus of the CTA-wide kernel: it must hand over to the pass program
    ('trimer_ring_modes', ring(3), dict(coll_rates=[0.5]), [(1, 1), (2, 1)]),
    ('tetramer_ring_modes', ring(4), dict(coll_rates=[0.5]), [(1, 1), (1, 0)]),
]


def run_engine(engine, prog, bound, sigma0, use_rr, env=None):
    old = {}
    for key, val in (env or {}).items():
        old[key] = os.environ.get(key)
        os.environ[key] = val
    try:
        dev = engine.upload_program(bound, prog if use_rr else None)
        sig = engine.to_device(sigma0.reshape(1, -1).astype(complex))
        emit = np.arange(STEPS + 1)
        _, snap = engine.evolve(dev, sig, 1, STEPS, emit, snapshots=True)
        torch.cuda.synchronize()
        return snap.cpu().numpy().reshape(STEPS + 1, *sigma0.shape), dev
    finally:
        for key, val in old.items():
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val


@pytest.mark.parametrize('name,system,kw,blocks', CASES, ids=[c[0] for c in CASES])
def test_all_kernel_families_agree_with_emulated_operations(engine, name, system, kw, blocks):
    dt = 0.15
    tn = kw.get('Trotter_number', 1)
    prog = _program_for(system, dt, dt / tn, tn, kw.get('Trotter_order', 1), kw.get('coll_rates'))
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    seen = set()
    for ka, kb in blocks:
        blk = prog.block(ka, kb)
        sigma0 = rng.normal(size=(blk.dim_a, blk.dim_b)) + 1j * rng.normal(size=(blk.dim_a, blk.dim_b))
        plain = prog.bind(blk, fused=False)
        want = [sigma0]
        for _ in range(STEPS):
            want.append(emulator.apply_step(want[-1], plain))
        want = np.array(want)
        scale = np.abs(want).max()
        variants = [('op list', plain, False, None), ('pass program', prog.bind(blk), False, None),
                    ('register-resident, generic', plain, True, {'QSX_RR_NO_STATIC': '1', 'QSX_DISABLE_RRC': '1'}),
                    ('register-resident, best', plain, True, {'QSX_RRC_V2': '0', 'QSX_RRC_V3': '0'}),
                    ('register-resident, throughput form of the CTA-wide kernel', plain, True, {'QSX_RRC_V2': '1', 'QSX_RRC_V3': '0'}),
                    ('register-resident, damping-diagonal form of the CTA-wide kernel', plain, True, {'QSX_RRC_V2': '1', 'QSX_RRC_V3': '1'})]
        for label, bound, use_rr, env in variants:
            got, dev = run_engine(engine, prog, bound, sigma0, use_rr, env)
            err = np.abs(got - want).max()
            assert err < 1e-12 * scale, (name, (ka, kb), label, err)
            seen.add(label + (' [rr]' if 'rr' in dev else ' [rrc]' if dev.get('rrc') is not None else ''))
    # the register-resident forms must actually have been exercised where the structure allows it
    if name in ('dimer_modes', 'dimer_markov'):
        assert any('[rr]' in s for s in seen)
    if name in ('trimer_modes', 'chain4_modes', 'trimer_ring_modes', 'tetramer_ring_modes'):
        assert any('[rrc]' in s for s in seen)          # rings: kernels specialised at run time (engine/jit.py)


def test_parameter_sets_assembled_on_the_device_match_the_host_recipe(engine):
    """qsx_rr_tables against RRProgram.params (NumPy) for a batch of random parameter sets, with and without folds."""
    from qsextra_b200.engine import rr as rr_mod
    from qsextra_b200.engine.program import StepProgram, SystemBatch
    rng = np.random.default_rng(17)
    B = 37
    batch = SystemBatch(rng.uniform(1.4, 1.6, (B, 2)), np.broadcast_to(np.array([[0., -0.05], [-0.05, 0.]]), (B, 2, 2)).copy(),
                        np.ones((B, 2)), rng.uniform(0.05, 0.2, (B, 1)), rng.uniform(0.05, 0.2, (B, 1)), (2,))
    markov = SystemBatch(rng.uniform(1.4, 1.6, (B, 2)), np.broadcast_to(np.array([[0., -0.05], [-0.05, 0.]]), (B, 2, 2)).copy(),
                         np.ones((B, 2)))
    cases = [(StepProgram(batch, 0.15, 1, 1, rng.uniform(0.1, 0.4, (B, 1))), [(1, 1), (2, 1), (1, 0)]),
             (StepProgram(markov, 0.15, 1, 1, rng.uniform(0.01, 0.1, B)), [(1, 1), (2, 1)])]
    seen_fold = False
    for prog, blocks in cases:
        for ka, kb in blocks:
            bound = prog.bind(prog.block(ka, kb), fused=False)
            rrp = rr_mod.build(prog, bound)
            assert rrp is not None
            seen_fold |= bool(rrp.fold_classes)
            got = rrp.device_params(engine).cpu().numpy()
            want = rrp.params
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= 2e-15 * max(1.0, np.abs(want).max())
    assert seen_fold


@pytest.mark.parametrize('name,system,kw,blocks', CASES, ids=[c[0] for c in CASES])
def test_parameter_rows_evaluated_on_the_device_match_the_host(engine, name, system, kw, blocks):
    """qsx_bind_params against ParamRecipe.evaluate (NumPy): same recipe, device sincos instead of libm."""
    tn = kw.get('Trotter_number', 1)
    prog = _program_for(system, 0.15, 0.15 / tn, tn, kw.get('Trotter_order', 1), kw.get('coll_rates'))
    for ka, kb in blocks:
        bound = prog.bind(prog.block(ka, kb), fused=False)
        assert bound._params is None                       # nothing evaluated on the host yet
        got = engine.bind_params(bound).cpu().numpy()
        want = bound.params
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-15


@pytest.mark.parametrize('name,system,kw,blocks', CASES, ids=[c[0] for c in CASES])
def test_transposed_programs_agree_with_emulated_operations(engine, name, system, kw, blocks):
    """The step run backwards with transposed operations (read-out propagation, QSX_OPF_TRANSPOSE): operation list
    and, where a compile-time program exists, the CTA-wide kernel, against the emulation."""
    dt = 0.15
    tn = kw.get('Trotter_number', 1)
    prog = _program_for(system, dt, dt / tn, tn, kw.get('Trotter_order', 1), kw.get('coll_rates'))
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 1)
    seen = set()
    for ka, kb in blocks:
        blk = prog.block(ka, kb)
        w0 = rng.normal(size=(blk.dim_a, blk.dim_b)) + 1j * rng.normal(size=(blk.dim_a, blk.dim_b))
        back = prog.bind(blk, fused=False).transposed()
        want = [w0]
        for _ in range(STEPS):
            want.append(emulator.apply_step(want[-1], back))
        want = np.array(want)
        for label, env in (('op list', {'QSX_DISABLE_RRC': '1'}), ('best', None)):
            got, dev = run_engine(engine, prog, back, w0, True, env)
            assert 'rr' not in dev
            assert np.abs(got - want).max() < 1e-12 * np.abs(want).max(), (name, (ka, kb), label)
            seen.add('rrc' if dev.get('rrc') is not None else 'ops')
    if name in ('trimer_modes', 'chain4_modes'):
        assert 'rrc' in seen


@pytest.mark.parametrize('m,n,k', [(1, 1, 1), (7, 5, 3), (33, 65, 17), (64, 128, 1024), (100, 3, 6144), (2048, 128, 96)])
def test_readout_gemm_against_numpy(engine, m, n, k):
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    a = rng.normal(size=(m, k)) + 1j * rng.normal(size=(m, k))
    b = rng.normal(size=(n, k)) + 1j * rng.normal(size=(n, k))
    got = engine.readout_gemm(torch.as_tensor(a, device='cuda'), torch.as_tensor(b, device='cuda')).cpu().numpy()
    want = a @ b.T
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-12 * np.sqrt(k) * max(1.0, np.abs(want).max())


def test_throughput_form_of_the_cta_wide_kernel_with_many_instances(engine):
    """qsx_rrc2.cuh (two CTAs per SM, folded damping scalings, shear hops) and qsx_rrc3.cuh (damping diagonalised by a
    change of coordinates that is undone around every emission) against the first form and the emulated
    operations: 300 instances of the 4-site chain's (1,1) block with different parameter sets and shared source rows,
    snapshots and trace read-outs at irregular steps (step 0 included: no folded scaling is pending there)."""
    from qsextra_b200.engine.program import StepProgram, SystemBatch
    from qsextra_b200.engine.sectors import population_observables
    rng = np.random.default_rng(11)
    n_sets, n_src, n_inst = 5, 7, 300
    J = np.zeros((n_sets, 4, 4))
    for i in range(3):
        J[:, i, i + 1] = J[:, i + 1, i] = -0.3 + 0.02 * rng.normal(size=n_sets)
    batch = SystemBatch(1.5 + 0.1 * rng.normal(size=(n_sets, 4)), J, np.ones((n_sets, 4)),
                        0.3 + 0.05 * rng.normal(size=(n_sets, 1)), np.full((n_sets, 1), 0.7), (2,))
    prog = StepProgram(batch, 0.15, 1, 1, 0.5 + 0.1 * rng.random((n_sets, 1)))
    for ka, kb in ((1, 1), (0, 1), (1, 0)):
        blk = prog.block(ka, kb)
        bound = prog.bind(blk, fused=False)
        sigma0 = rng.normal(size=(n_src, blk.size)) + 1j * rng.normal(size=(n_src, blk.size))
        inst_set = torch.as_tensor(rng.integers(0, n_sets, n_inst), dtype=torch.int32, device=engine.device)
        inst_src = torch.as_tensor(rng.integers(0, n_src, n_inst), dtype=torch.int32, device=engine.device)
        emit = [0, 1, 4, 9]
        results = {}
        for flag, v3 in (('0', '0'), ('1', '0'), ('3', '1')):
            os.environ['QSX_RRC_V2'] = '0' if flag == '0' else '1'
            os.environ['QSX_RRC_V3'] = v3
            try:
                dev = engine.upload_program(bound, prog)
                assert dev.get('rrc') is not None
                _, snap = engine.evolve(dev, engine.to_device(sigma0), n_inst, 9, emit, snapshots=True, inst_set=inst_set,
                                        inst_src=inst_src)
                obs = None
                if ka == kb:
                    obs = engine.upload_observables(*population_observables(blk), kind='populations')
                    pops, _ = engine.evolve(dev, engine.to_device(sigma0), n_inst, 9, emit, obs=obs, obs_real=True,
                                            inst_set=inst_set, inst_src=inst_src)
                else:
                    mu = np.array([1., 0.8, -0.6, 1.1])
                    obs = engine.upload_observables(*blk.emission_observable(mu), kind='emission', mu=mu)
                    pops, _ = engine.evolve(dev, engine.to_device(sigma0), n_inst, 9, emit, obs=obs, inst_set=inst_set,
                                            inst_src=inst_src)
                torch.cuda.synchronize()
                results[flag] = (snap.cpu().numpy(), pops.cpu().numpy())
            finally:
                os.environ.pop('QSX_RRC_V2', None)
                os.environ.pop('QSX_RRC_V3', None)
        scale = np.abs(results['0'][0]).max()
        for flag in ('1', '3'):
            assert np.abs(results['0'][0] - results[flag][0]).max() < 1e-12 * scale, (ka, kb, flag)
            assert np.abs(results['0'][1] - results[flag][1]).max() < 1e-12 * np.abs(results['0'][1]).max(), (ka, kb, flag)
        sets, srcs = inst_set.cpu().numpy(), inst_src.cpu().numpy()
        for inst in (0, 151, 299):                           # the emulated operations, a few instances
            sigma = sigma0[srcs[inst]].reshape(blk.dim_a, blk.dim_b)
            for step in range(10):
                if step:
                    sigma = emulator.apply_step(sigma, bound, int(sets[inst]))
                if step in emit:
                    for flag in ('1', '3'):
                        got = results[flag][0][inst, emit.index(step)].reshape(blk.dim_a, blk.dim_b)
                        assert np.abs(got - sigma).max() < 1e-12 * scale, (ka, kb, flag, inst, step)


def test_segment_scheduling_of_the_damping_diagonal_form(engine, monkeypatch):
    """With more instances than resident CTAs and nothing but snapshots to write, qsx_rrc3.cuh lets its CTAs claim
    (instance, segment) pairs and continue an instance from the snapshot of its previous segment (qsx_rrc_args.progress).
    Same snapshots as one CTA per instance from start to end: 700 instances of the (1,1) and (0,1) blocks, different
    parameter sets, shared source rows, irregular emission steps with and without step 0."""
    from qsextra_b200.engine.program import StepProgram, SystemBatch
    rng = np.random.default_rng(23)
    n_sets, n_src, n_inst = 5, 7, 700
    J = np.zeros((n_sets, 4, 4))
    for i in range(3):
        J[:, i, i + 1] = J[:, i + 1, i] = -0.3 + 0.02 * rng.normal(size=n_sets)
    batch = SystemBatch(1.5 + 0.1 * rng.normal(size=(n_sets, 4)), J, np.ones((n_sets, 4)),
                        0.3 + 0.05 * rng.normal(size=(n_sets, 1)), np.full((n_sets, 1), 0.7), (2,))
    prog = StepProgram(batch, 0.15, 1, 1, 0.5 + 0.1 * rng.random((n_sets, 1)))
    monkeypatch.setenv('QSX_RRC_V2', '1')
    monkeypatch.setenv('QSX_RRC_V3', '1')
    for ka, kb in ((1, 1), (0, 1)):
        blk = prog.block(ka, kb)
        bound = prog.bind(blk, fused=False)
        sigma0 = rng.normal(size=(n_src, blk.size)) + 1j * rng.normal(size=(n_src, blk.size))
        inst_set = torch.as_tensor(rng.integers(0, n_sets, n_inst), dtype=torch.int32, device=engine.device)
        inst_src = torch.as_tensor(rng.integers(0, n_src, n_inst), dtype=torch.int32, device=engine.device)
        for emit in ([0, 3, 4, 9, 17], [2, 5, 17]):
            snaps = {}
            for mode in ('segments', 'whole instances'):
                if mode == 'whole instances':
                    monkeypatch.setenv('QSX_DISABLE_SEGMENTS', '1')
                else:
                    monkeypatch.delenv('QSX_DISABLE_SEGMENTS', raising=False)
                dev = engine.upload_program(bound, prog)
                _, snap = engine.evolve(dev, engine.to_device(sigma0), n_inst, 17, emit, snapshots=True, inst_set=inst_set,
                                        inst_src=inst_src)
                torch.cuda.synchronize()
                assert dev.get('rrc') is not None
                snaps[mode] = snap.cpu().numpy()
            scale = np.abs(snaps['whole instances']).max()
            assert np.abs(snaps['segments'] - snaps['whole instances']).max() < 1e-12 * scale, (ka, kb, emit)
    monkeypatch.delenv('QSX_DISABLE_SEGMENTS', raising=False)


def test_damping_diagonal_form_of_the_transposed_steps(engine):
    """qsx_rrc3t.cuh (read-out propagations: z(1,1) = W(1,1) - W(0,0) per damped pseudomode, dampings folded into the
    diagonal multiply that ends the step, transform undone and redone around every emission) against the first form
    and the emulated transposed operations: the (1,0), (0,1) blocks (ket rotations on disjoint rows: one shared round)
    and the two-exciton block (2,1) (every row belongs to two ket rotations: one round per rotation), several parameter
    sets, snapshots at irregular steps including step 0 and consecutive steps."""
    from qsextra_b200.engine.program import StepProgram, SystemBatch
    rng = np.random.default_rng(17)
    n_sets, n_inst = 4, 9
    J = np.zeros((n_sets, 4, 4))
    for i in range(3):
        J[:, i, i + 1] = J[:, i + 1, i] = -0.3 + 0.02 * rng.normal(size=n_sets)
    batch = SystemBatch(1.5 + 0.1 * rng.normal(size=(n_sets, 4)), J, np.ones((n_sets, 4)),
                        0.3 + 0.05 * rng.normal(size=(n_sets, 1)), np.full((n_sets, 1), 0.7), (2,))
    prog = StepProgram(batch, 0.15, 1, 1, 0.5 + 0.1 * rng.random((n_sets, 1)))
    emit = [0, 1, 2, 7, 12]
    for ka, kb in ((1, 0), (0, 1), (2, 1)):
        blk = prog.block(ka, kb)
        bound = prog.bind(blk, fused=False).transposed()
        sigma0 = rng.normal(size=(n_inst, blk.size)) + 1j * rng.normal(size=(n_inst, blk.size))
        inst_set = torch.as_tensor(rng.integers(0, n_sets, n_inst), dtype=torch.int32, device=engine.device)
        results = {}
        for v3 in ('0', '1'):
            os.environ['QSX_RRC_V2'], os.environ['QSX_RRC_V3'] = '0', v3
            try:
                dev = engine.upload_program(bound, prog)
                assert dev.get('rrc') is not None
                _, snap = engine.evolve(dev, engine.to_device(sigma0), n_inst, 12, emit, snapshots=True, inst_set=inst_set)
                torch.cuda.synchronize()
                assert dev.get('rrc') is not None
                results[v3] = snap.cpu().numpy()
            finally:
                os.environ.pop('QSX_RRC_V2', None)
                os.environ.pop('QSX_RRC_V3', None)
        scale = np.abs(results['0']).max()
        assert np.abs(results['0'] - results['1']).max() < 1e-12 * scale, (ka, kb)
        sets = inst_set.cpu().numpy()
        for inst in (0, 4, 8):
            sigma = sigma0[inst].reshape(blk.dim_a, blk.dim_b)
            for step in range(13):
                if step:
                    sigma = emulator.apply_step(sigma, bound, int(sets[inst]))
                if step in emit:
                    got = results['1'][inst, emit.index(step)].reshape(blk.dim_a, blk.dim_b)
                    assert np.abs(got - sigma).max() < 1e-12 * scale, (ka, kb, inst, step)


def test_structures_outside_the_menu_are_specialised_at_run_time(engine, tmp_path, monkeypatch):
    """A 4-site ring and a chain with a missing bond have no compile-time program in csrc/qsx_rrc_programs.inc: the
    CTA-wide kernels are compiled for them with NVRTC on first use (engine/jit.py), registered through the C ABI and
    serve the forward, the throughput and the transposed form -- no silent drop to the shared-memory pass program."""
    from qsextra_b200.engine import jit
    from qsextra_b200.engine.program import StepProgram, SystemBatch
    monkeypatch.setenv('QSX_JIT_CACHE', str(tmp_path))
    rng = np.random.default_rng(21)
    n_sets = 3
    for label, bonds in (('ring', [(0, 1), (1, 2), (2, 3), (0, 3)]), ('broken chain', [(0, 1), (2, 3)]),
                         ('long range', [(0, 1), (1, 2), (2, 3), (0, 2)])):
        J = np.zeros((n_sets, 4, 4))
        for i, j in bonds:
            J[:, i, j] = J[:, j, i] = -0.3 + 0.03 * rng.normal(size=n_sets)
        batch = SystemBatch(1.5 + 0.1 * rng.normal(size=(n_sets, 4)), J, np.ones((n_sets, 4)),
                            0.3 + 0.05 * rng.normal(size=(n_sets, 1)), np.full((n_sets, 1), 0.7), (2,))
        prog = StepProgram(batch, 0.15, 1, 1, 0.5 + 0.1 * rng.random((n_sets, 1)))
        for ka, kb, transposed in ((1, 1, False), (2, 1, False), (2, 1, True)):
            blk = prog.block(ka, kb)
            plain = prog.bind(blk, fused=False)
            bound = plain.transposed() if transposed else plain
            n_inst = 5
            sigma0 = rng.normal(size=(n_inst, blk.size)) + 1j * rng.normal(size=(n_inst, blk.size))
            inst_set = torch.as_tensor(rng.integers(0, n_sets, n_inst), dtype=torch.int32, device=engine.device)
            for v2 in ('0', '1'):
                monkeypatch.setenv('QSX_RRC_V2', v2)
                dev = engine.upload_program(bound, prog)
                assert dev.get('rrc') is not None, label
                assert engine.has_static_rrc(dev) and dev['rrc_family'] == 'specialised at run time', (label, ka, kb)
                _, snap = engine.evolve(dev, engine.to_device(sigma0), n_inst, 3, [0, 1, 3], snapshots=True, inst_set=inst_set)
                torch.cuda.synchronize()
                assert dev.get('rrc') is not None                   # the specialised kernel ran (no hand-over)
                got = snap.cpu().numpy()
                for inst in range(n_inst):
                    sigma = sigma0[inst].reshape(blk.dim_a, blk.dim_b)
                    for step in range(4):
                        if step:
                            sigma = emulator.apply_step(sigma, bound, int(inst_set[inst]))
                        if step in (0, 1, 3):
                            err = np.abs(got[inst, [0, 1, 3].index(step)].reshape(blk.dim_a, blk.dim_b) - sigma).max()
                            assert err < 1e-12 * np.abs(sigma0).max(), (label, ka, kb, transposed, v2, inst, step, err)
    assert len([f for f in os.listdir(str(tmp_path)) if f.endswith('.cubin')]) >= 6          # cached on disk
    monkeypatch.delenv('QSX_RRC_V2')
