"""The response tree (engine/response.py: levels, prefix sharing, sharding over ranks, read-out propagation for the last
delay, gather) run on the CPU with a NumPy stand-in for the engine (tests/emulator.py emulates each kernel), alone and
under a world_size-2 gloo group, against the goldens of the unmodified reference."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import emulator
from helpers import build_spectroscopy, build_system, load_golden
from qsextra_b200.engine.response import response_function
from qsextra_b200.qcomo.evolution.qevolve import _program_for, _validate_rates
from qsextra_b200.spectroscopy.simulation.qspectroscopy import delay_step_counts

CASES = ['qspectroscopy_esa_dimer_markov', 'qspectroscopy_se_dimer_t2', 'qspectroscopy_custom_kbkk_dimer',
         'qspectroscopy_gsb_dimer_modes', 'qspectroscopy_se_trimer_markov']


class NumpyEngine:
    """The part of ``runtime.Engine`` the response tree uses, with every kernel replaced by its NumPy emulation."""

    def __init__(self):
        self.device = torch.device('cpu')
        self.transposed_runs = 0

    def to_device(self, array, dtype=None):
        t = torch.as_tensor(np.ascontiguousarray(array))
        return t if dtype is None else t.to(dtype)

    def apply_dipole(self, in_block, out_block, side, mu, sigma):
        rows = sigma.numpy().reshape(-1, in_block.dim_a, in_block.dim_b)
        out = [emulator.apply_dipole(r, in_block, out_block, side, mu.numpy()) for r in rows]
        return torch.as_tensor(np.array(out).reshape(len(out), -1))

    def upload_program(self, bound, program=None):
        return dict(bound=bound)

    def upload_observables(self, ptr, idx, w, kind=None, mu=None):
        return dict(host=(np.asarray(ptr), np.asarray(idx), np.asarray(w)))

    def evolve(self, prog, sigma_in, n_inst, n_steps, emit_steps, obs=None, snapshots=False, inst_set=None,
               inst_src=None, **_):
        bound = prog['bound']
        self.transposed_runs += int(bound.is_transposed)
        blk = bound.block
        emit = [int(e) for e in np.asarray(emit_steps)]
        rows = sigma_in.numpy().reshape(-1, blk.dim_a, blk.dim_b)
        snaps = np.zeros((n_inst, len(emit), blk.size), dtype=complex)
        for i in range(n_inst):
            sigma = rows[i if inst_src is None else int(inst_src[i])]
            member = 0 if inst_set is None else int(inst_set[i])
            for step in range(n_steps + 1):
                if step:
                    sigma = emulator.apply_step(sigma, bound, member)
                for k, e in enumerate(emit):
                    if e == step:
                        snaps[i, k] = sigma.reshape(-1)
        out_obs = None
        if obs is not None:
            ptr, idx, w = obs['host']
            out_obs = torch.as_tensor((snaps[:, :, idx[ptr[0]:ptr[1]]] * w[ptr[0]:ptr[1]]).sum(-1)[:, :, None])
        return out_obs, (torch.as_tensor(snaps) if snapshots else None)

    def readout_gemm(self, sigma, readouts):
        return sigma @ readouts.transpose(-1, -2)


def signal_of(name, engine):
    meta, arrays = load_golden(name)
    system = build_system(meta['system'])
    spec = build_spectroscopy(meta['label'], meta['delay_time'])
    kw = dict(meta['kwargs'])
    counts, (dt, dt_trotter, tn, _) = delay_step_counts(spec.delay_time, kw['dt'], kw.get('Trotter_number', 1))
    program = _program_for(system, dt, dt_trotter, tn, kw.get('Trotter_order', 1), _validate_rates(system, kw.get('coll_rates')))
    got = response_function(program, np.asarray(system.dipole_moments, dtype=float), spec.side_seq, spec.direction_seq,
                            counts, engine=engine)
    return got, arrays['signal']


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('heisenberg', [False, True])
def test_tree_on_the_cpu_stand_in(name, heisenberg, monkeypatch):
    monkeypatch.setenv('QSX_DISABLE_RR', '1')
    monkeypatch.setenv('QSX_HEISENBERG_MIN', '1' if heisenberg else '1000000')
    engine = NumpyEngine()
    got, want = signal_of(name, engine)
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-10 * max(1.0, np.abs(want).max())
    assert (engine.transposed_runs > 0) == heisenberg


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, size, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), QSX_DISABLE_RR='1', QSX_HEISENBERG_MIN='2')
    dist.init_process_group('gloo', rank=rank, world_size=size)
    try:
        worst, transposed = 0.0, 0
        for name in ('qspectroscopy_esa_dimer_t2', 'qspectroscopy_se_trimer_markov'):
            engine = NumpyEngine()
            got, want = signal_of(name, engine)
            worst = max(worst, float(np.abs(got - want).max()))
            transposed += engine.transposed_runs
        np.save(os.path.join(out_dir, 'r%d.npy' % rank), np.array([worst, transposed]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('size', [2, 3])
def test_tree_sharded_over_a_gloo_group(tmp_path, size):
    mp.spawn(_worker, args=(size, _free_port(), str(tmp_path)), nprocs=size, join=True)
    for r in range(size):
        worst, transposed = np.load(os.path.join(str(tmp_path), 'r%d.npy' % r))
        assert worst < 1e-10                                  # every rank holds the full, correct signal


def test_set_of_diagrams_shares_levels_and_read_outs(monkeypatch):
    """response_function_set on the stand-in engine: identical signals, fewer propagations."""
    from qsextra_b200.engine.response import response_function_set
    from qsextra_b200.spectroscopy import FeynmanDiagram
    monkeypatch.setenv('QSX_DISABLE_RR', '1')
    monkeypatch.setenv('QSX_HEISENBERG_MIN', '1')
    meta, _ = load_golden('qspectroscopy_esa_dimer_t2')
    system = build_system(meta['system'])
    kw = dict(meta['kwargs'])
    diagrams = [FeynmanDiagram(lab, meta['delay_time']) for lab in ('gsb', 'se', 'esa')]
    counts, (dt, dt_trotter, tn, _) = delay_step_counts(diagrams[0].delay_time, kw['dt'], kw.get('Trotter_number', 1))
    program = _program_for(system, dt, dt_trotter, tn, kw.get('Trotter_order', 1), _validate_rates(system, kw.get('coll_rates')))
    mu = np.asarray(system.dipole_moments, dtype=float)

    class Counting(NumpyEngine):
        evolves = 0

        def evolve(self, *a, **k):
            Counting.evolves += 1
            return super().evolve(*a, **k)

    alone = [response_function(program, mu, d.side_seq, d.direction_seq, counts, engine=Counting()) for d in diagrams]
    separately, Counting.evolves = Counting.evolves, 0
    together = response_function_set(program, mu, [(d.side_seq, d.direction_seq) for d in diagrams], counts, engine=Counting())
    assert all(np.array_equal(a, b) for a, b in zip(alone, together))
    assert Counting.evolves < separately                       # 5 instead of 9 propagations
    assert Counting.evolves == 5 and separately == 9


# ---- ensembles: the parameter sets of one step program through one tree -----------------------------------------
def _ensemble(n_members):
    import warnings
    from qsextra_b200 import ChromophoreSystem
    rng = np.random.default_rng(5)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return [ChromophoreSystem(energies=(np.array([1.55, 1.46]) + 0.02 * rng.standard_normal(2)).tolist(),
                                  couplings=-0.01 + 0.002 * rng.standard_normal(), dipole_moments=[1., 0.8],
                                  frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=0.054)
                for _ in range(n_members)]


def _ensemble_signals(systems, labels, delays, engine):
    from qsextra_b200.engine.program import StepProgram, SystemBatch
    from qsextra_b200.engine.response import response_function_set
    from qsextra_b200.spectroscopy import FeynmanDiagram
    dt = 0.15
    diagrams = [FeynmanDiagram(lab, delays) for lab in labels]
    counts, (dt_, dt_trotter, tn, _) = delay_step_counts(diagrams[0].delay_time, dt, 1)
    rates = np.array([[0.2 + 0.01 * b] for b in range(len(systems))])
    program = StepProgram(SystemBatch.from_systems(systems), dt_, tn, 1, rates, dt_trotter=dt_trotter)
    mu = np.asarray(systems[0].dipole_moments, dtype=float)
    return response_function_set(program, mu, [(d.side_seq, d.direction_seq) for d in diagrams], counts, engine=engine)


_ENSEMBLE_DELAYS = [[0., 0.3, 0.15], [0.15, 0.45], [0., 0.15, 0.6]]


@pytest.mark.parametrize('heisenberg', [False, True])
def test_ensemble_equals_member_by_member(heisenberg, monkeypatch):
    monkeypatch.setenv('QSX_DISABLE_RR', '1')
    monkeypatch.setenv('QSX_HEISENBERG_MIN', '1' if heisenberg else '1000000')
    systems = _ensemble(3)
    engine = NumpyEngine()
    together = _ensemble_signals(systems, ('gsb', 'esa'), _ENSEMBLE_DELAYS, engine)
    assert (engine.transposed_runs > 0) == heisenberg
    for d, sig in enumerate(together):
        assert sig.shape == (3, 3, 2, 3)
        for b, system in enumerate(systems):
            # a member alone, with ITS rate: a one-set program
            from qsextra_b200.engine.program import StepProgram, SystemBatch
            from qsextra_b200.spectroscopy import FeynmanDiagram
            diagram = FeynmanDiagram(('gsb', 'esa')[d], _ENSEMBLE_DELAYS)
            counts, (dt_, dt_trotter, tn, _) = delay_step_counts(diagram.delay_time, 0.15, 1)
            program = StepProgram(SystemBatch.from_systems([system]), dt_, tn, 1, np.array([[0.2 + 0.01 * b]]), dt_trotter=dt_trotter)
            alone = response_function(program, np.asarray(system.dipole_moments, dtype=float), diagram.side_seq,
                                      diagram.direction_seq, counts, engine=NumpyEngine())
            assert np.abs(sig[b] - alone).max() < 1e-12


def _ensemble_worker(rank, size, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), QSX_DISABLE_RR='1', QSX_HEISENBERG_MIN='2')
    dist.init_process_group('gloo', rank=rank, world_size=size)
    try:
        sigs = _ensemble_signals(_ensemble(3), ('se', 'esa'), _ENSEMBLE_DELAYS, NumpyEngine())
        np.save(os.path.join(out_dir, 'e%d.npy' % rank), np.array(sigs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('size', [2, 4])
def test_ensemble_members_sharded_over_a_gloo_group(tmp_path, size, monkeypatch):
    """Members over ranks (4 ranks for 3 members: one rank holds nothing), one gather, every rank the whole result."""
    monkeypatch.setenv('QSX_DISABLE_RR', '1')
    monkeypatch.setenv('QSX_HEISENBERG_MIN', '2')
    want = np.array(_ensemble_signals(_ensemble(3), ('se', 'esa'), _ENSEMBLE_DELAYS, NumpyEngine()))
    mp.spawn(_ensemble_worker, args=(size, _free_port(), str(tmp_path)), nprocs=size, join=True)
    for r in range(size):
        got = np.load(os.path.join(str(tmp_path), 'e%d.npy' % r))
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-12


# ---- basis substitution: many chains entering a level replaced by the basis of their structural support ----------
_BASIS_DELAYS = [(0.15 * np.arange(14)).tolist(), [0., 0.3, 0.15], [0., 0.45, 0.15, 0.9]]


def _basis_case(labels, engine, n_members=1):
    systems = _ensemble(n_members)
    return systems, _ensemble_signals(systems, labels, _BASIS_DELAYS, engine)


class _CountingEngine(NumpyEngine):
    def __init__(self):
        super().__init__()
        self.instances = []

    def evolve(self, prog, sigma_in, n_inst, *a, **k):
        self.instances.append(int(n_inst))
        return super().evolve(prog, sigma_in, n_inst, *a, **k)


@pytest.mark.parametrize('heisenberg', [False, True])
@pytest.mark.parametrize('n_members', [1, 3])
def test_basis_substitution_equals_chain_by_chain(heisenberg, n_members, monkeypatch):
    """14 t1 values enter the t2 level of a dimer with pseudomodes; the (0,1) / (1,0) block they live in has a support
    of 8 elements (the pseudomodes of the ground side stay empty): 8 basis elements are propagated instead of 14
    chains and the signal is contracted with the coefficients.  Same numbers as the plain tree and as the oracle."""
    monkeypatch.setenv('QSX_DISABLE_RR', '1')
    monkeypatch.setenv('QSX_HEISENBERG_MIN', '1' if heisenberg else '1000000')
    labels = ('gsb', 'se', 'esa')
    with_basis = _CountingEngine()
    systems, got = _basis_case(labels, with_basis, n_members)
    monkeypatch.setenv('QSX_DISABLE_BASIS', '1')
    plain = _CountingEngine()
    _, want = _basis_case(labels, plain, n_members)
    assert max(with_basis.instances) == 8 * 3 * n_members and max(plain.instances) == 14 * 3 * n_members or heisenberg
    assert sum(with_basis.instances) < sum(plain.instances)
    for a, b in zip(got, want):
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-12
    if n_members == 1:
        from oracle import reference_path as R
        from qsextra_b200.spectroscopy import FeynmanDiagram
        for lab, sig in zip(labels, got):
            ref = R.qspectroscopy_operator(systems[0], FeynmanDiagram(lab, _BASIS_DELAYS), dt=0.15, coll_rates=[0.2])
            assert np.abs(sig - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())


def _basis_worker(rank, size, port, out_dir, n_members):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), QSX_DISABLE_RR='1', QSX_HEISENBERG_MIN='2')
    dist.init_process_group('gloo', rank=rank, world_size=size)
    try:
        _, sigs = _basis_case(('gsb', 'esa'), NumpyEngine(), n_members)
        np.save(os.path.join(out_dir, 'b%d.npy' % rank), np.array(sigs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('size,n_members', [(2, 1), (3, 1), (2, 3)])
def test_basis_rows_sharded_over_a_gloo_group(tmp_path, size, n_members, monkeypatch):
    """One system: the basis rows are the sharded units and the contraction follows the gather; an ensemble: members are
    sharded and every rank contracts its own."""
    monkeypatch.setenv('QSX_DISABLE_RR', '1')
    monkeypatch.setenv('QSX_HEISENBERG_MIN', '2')
    monkeypatch.setenv('QSX_DISABLE_BASIS', '1')
    _, want = _basis_case(('gsb', 'esa'), NumpyEngine(), n_members)
    monkeypatch.delenv('QSX_DISABLE_BASIS')
    mp.spawn(_basis_worker, args=(size, _free_port(), str(tmp_path), n_members), nprocs=size, join=True)
    for r in range(size):
        got = np.load(os.path.join(str(tmp_path), 'b%d.npy' % r))
        assert np.abs(got - np.array(want)).max() < 1e-12


def test_structural_support_is_closed_under_the_step():
    """The closure of engine/support.py contains every element the emulated step populates, and is minimal for the
    block after the first interaction (empty pseudomodes on the ground side)."""
    from qsextra_b200.engine import support
    from qsextra_b200.engine.sectors import Block
    system = _ensemble(1)[0]
    counts, (dt_, dt_trotter, tn, _) = delay_step_counts([[0.15]], 0.15, 1)
    program = _program_for(system, dt_, dt_trotter, tn, 1, _validate_rates(system, [0.2]))
    ground = program.block(0, 0)
    for (ka, kb, side), expected in (((0, 1, 'b'), 8), ((1, 0, 'k'), 8)):
        block = Block(program.n_sites, program.n_bits, ka, kb)
        mask = support.step_closure(program, block, support.dipole_support(support.ground_support(ground), ground, block, side))
        assert int(mask.sum()) == expected
        sigma = np.zeros((block.dim_a, block.dim_b), dtype=complex)
        sigma[mask] = 1.0 + 0.5j
        bound = program.bind(block)
        for _ in range(12):
            sigma = emulator.apply_step(sigma, bound, 0)
        assert np.all(mask[np.abs(sigma) > 0])


# ---- ensembles split by the caller: every rank lowers its own members, the tree runs unsharded, one gather --------
def _presplit_worker(rank, size, port, out_dir, n_members):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), QSX_DISABLE_RR='1', QSX_HEISENBERG_MIN='2')
    dist.init_process_group('gloo', rank=rank, world_size=size)
    try:
        from qsextra_b200.engine import dist as qdist
        from qsextra_b200.engine.program import StepProgram, SystemBatch
        from qsextra_b200.engine.response import _finish, empty_share, response_function_set
        from qsextra_b200.spectroscopy import FeynmanDiagram
        systems = _ensemble(n_members)
        lo, hi = qdist.shard_bounds(n_members, rank, size)
        diagrams = [FeynmanDiagram(lab, _BASIS_DELAYS) for lab in ('gsb', 'esa')]
        sequences = [(d.side_seq, d.direction_seq) for d in diagrams]
        counts, (dt_, dt_trotter, tn, _) = delay_step_counts(diagrams[0].delay_time, 0.15, 1)
        engine = NumpyEngine()
        if hi > lo:
            rates = np.array([[0.2 + 0.01 * b] for b in range(n_members)])[lo:hi]
            program = StepProgram(SystemBatch.from_systems(systems[lo:hi]), dt_, tn, 1, rates, dt_trotter=dt_trotter)
            mu = np.asarray(systems[0].dipole_moments, dtype=float)
            sigs = response_function_set(program, mu, sequences, counts, engine=engine, global_sets=n_members)
        else:
            sigs = [_finish(engine, out, info, False) for out, info in zip(*empty_share(engine, sequences, counts, n_members))]
        np.save(os.path.join(out_dir, 'p%d.npy' % rank), np.array(sigs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('size,n_members', [(2, 3), (4, 3), (2, 2)])
def test_ensemble_split_by_the_caller(tmp_path, size, n_members, monkeypatch):
    """``global_sets``: ranks hold 2 + 1 members, 1 + 1 + 1 + 0 members (one rank only joins the gathers) or one member
    each (a one-set program locally); every rank ends up with all members' signals."""
    monkeypatch.setenv('QSX_DISABLE_RR', '1')
    monkeypatch.setenv('QSX_HEISENBERG_MIN', '2')
    want = np.array(_ensemble_signals(_ensemble(n_members), ('gsb', 'esa'), _BASIS_DELAYS, NumpyEngine()))
    mp.spawn(_presplit_worker, args=(size, _free_port(), str(tmp_path), n_members), nprocs=size, join=True)
    for r in range(size):
        got = np.load(os.path.join(str(tmp_path), 'p%d.npy' % r))
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-12
