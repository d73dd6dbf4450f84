"""Response-function evaluation on the device: the Feynman-pathway sum of ``qspectroscopy`` as a tree of
batched evolutions.

The reference (qspectroscopy.py:117-183) builds, for every delay-grid point, N**(order+1) site pathways x {1,2,3}
X/Y circuits, each re-evolved from t = 0.  Because everything is linear in sigma = |ket><bra|:

  * the sum over sites of one interaction is ONE collective operator  sum_s mu_s O_s  (``qsx_apply_dipole``);
    O = X for the first interaction on a side, else |0><1| if (side == 'k') xor (direction == 'i'), else |1><0|
    -- the coefficient rules mu, mu/2, +-i mu/2 of :144-174;  ket side: sigma -> O sigma, bra side: sigma -> sigma O;
  * <X_kb> + i<Y_kb> of the ketbra ancilla (:90-91, :180-183) is the trace of the block after the emission
    operator (a read-out of ``qsx_evolve``);
  * chains that share a prefix along t1 -> t2 -> t3 are evolved once: level n evolves every instance produced by
    level n-1 and snapshots it at each distinct step count of the n-th delay list;
  * the LAST delay is evaluated in the Heisenberg picture when it holds many chains: <W, Phi^n sigma> =
    <(Phi^T)^n W, sigma>, so the emission read-out W is propagated once with the transposed step (on a side stream,
    while the earlier delays run) and the signal is one [chains x E] . [E x emissions] product
    (``qsx_readout_gemm``) instead of one propagation per chain.

Independent chains are sharded over ranks at the first level that has enough of them; one all-gather assembles
the signal.

Ensembles.  A ``StepProgram`` bound to S > 1 parameter sets (static-disorder realisations of one aggregate: same
structure, same dipole moments, different energies / couplings / bath parameters) is evaluated in ONE tree: level n
holds S x (chains per set) instances, every instance carries the index of its parameter set (``inst_set``), each set
propagates its own read-out and the last level is one batched product.  The sets are what is sharded over the ranks
then (contiguous slices, one all-gather); the signal gets a leading axis of length S.
"""
from __future__ import annotations

import contextlib
import os

import numpy as np
import torch

from . import dist as qdist
from . import support as qsupport
from .program import StepProgram
from .sectors import Block


def interaction_plan(side_seq: str, direction_seq: str):
    """[(side, kind)] with kind in {'X', 'lower', 'raise'}; the last entry is the emission."""
    first = {'k': True, 'b': True}
    plan = []
    n_int = len(side_seq)
    for n, (side, direction) in enumerate(zip(side_seq, direction_seq)):
        if first[side]:
            kind = 'X'
            first[side] = False
        elif n == n_int - 1:
            kind = 'X'
        elif (side == 'k') ^ (direction == 'i'):
            kind = 'lower'
        else:
            kind = 'raise'
        plan.append((side, kind))
    return plan


def _fingerprint(program) -> bytes:
    """Digest of everything a propagation depends on: the structure of the step and every parameter value."""
    import hashlib
    h = hashlib.sha1()
    if getattr(program, 'continuous', False):
        for part in (program.energies, program.couplings, program.omega, program.coupl_ep, program.rates, program.levels):
            h.update(repr(None if part is None else np.asarray(part).tolist()).encode())
        return b'L' + h.digest()
    h.update(repr(program.structure_key()).encode())
    for op in program.lowered:
        for value in vars(op).values():
            if isinstance(value, np.ndarray):
                h.update(np.ascontiguousarray(value).tobytes())
            elif isinstance(value, (list, tuple)):
                for item in value:
                    for piece in (item if isinstance(item, (list, tuple)) else (item,)):
                        if isinstance(piece, np.ndarray):
                            h.update(np.ascontiguousarray(piece).tobytes())
    return b'S' + h.digest()


def n_parameter_sets(program) -> int:
    batch = getattr(program, 'batch', None)
    return int(batch.batch) if batch is not None and hasattr(batch, 'batch') else 1


def _last_level(program, plan, uniq, rank, size):
    """(block of the last delay or None if a sector is empty, chains PER PARAMETER SET that enter it, over all ranks)
    -- the bookkeeping of ``response_function`` without any propagation.  Rank-independent on purpose: every rank must
    take the same decisions from it (Heisenberg picture or not, and with it whether the last level is sharded)."""
    ka = kb = 0
    chains = 1
    block = None
    for level, (side, kind) in enumerate(plan[:-1]):
        delta = +1 if kind == 'X' or (kind == 'raise') == (side == 'k') else -1
        ka, kb = (ka + delta, kb) if side == 'k' else (ka, kb + delta)
        block = Block(program.n_sites, program.n_bits, ka, kb)
        if not block.valid():
            return None, 0
        if level < len(plan) - 2:
            chains *= len(uniq[level])
    return block, chains


def _propagate_readout(eng, program, mu, last_block, chains, emit, continuous, stream_index, fork_event=None):
    """Enqueue the propagation of the emission read-out of ``last_block`` on a side stream, if the Heisenberg picture
    is planned for this level.  -> (read-outs [n_emit, E] on the device, completion event | None) or None."""
    n_sets = n_parameter_sets(program)
    set_lo, set_hi = qdist.shard_bounds(n_sets, *qdist.world()) if n_sets > 1 else (0, 1)
    local_sets = set_hi - set_lo
    if local_sets == 0:
        return None                                     # more ranks than parameter sets: nothing to do on this one
    if (last_block is None or chains < int(os.environ.get('QSX_HEISENBERG_MIN', '8'))
            or os.environ.get('QSX_DISABLE_HEISENBERG')):
        return None
    back = None
    if not continuous:                                  # the register-resident small-block kernel is faster forwards
        from . import rr as rr_mod
        back = program.bind(last_block, fused=False)
        if not os.environ.get('QSX_DISABLE_RR') and rr_mod.build(program, back) is not None:
            return None
    ptr, idx, w = last_block.emission_observable(np.asarray(mu, dtype=float))
    readout = np.zeros(last_block.size, dtype=complex)
    np.add.at(readout, np.asarray(idx[ptr[0]:ptr[1]]), np.asarray(w[ptr[0]:ptr[1]]))
    on_gpu = eng.device.type == 'cuda'                   # (the CPU stand-in of the tests has no streams)
    side_stream = eng.side_stream_of(stream_index) if on_gpu else None
    if on_gpu:                                          # fork: after the caller's fork point if it has one (so that the
        if fork_event is not None:                      # propagation does not queue behind kernels enqueued since),
            side_stream.wait_event(fork_event)          # else after everything on the current stream
        else:
            side_stream.wait_stream(torch.cuda.current_stream(eng.device))
    readouts = None
    with (torch.cuda.stream(side_stream) if on_gpu else contextlib.nullcontext()):
        start = eng.to_device(readout.reshape(1, -1))
        if continuous:
            from .lindblad import segments_from_times
            dev_back = eng.upload_lindblad(program.bind(last_block, transposed=True))
            _, snap = eng.evolve_lindblad(dev_back, start, 1, segments_from_times(emit), snapshots=True)
            readouts = snap[0]
        else:
            dev_back = eng.upload_program(back.transposed(), program)
            # One read-out walks the steps alone (latency-bound); the chains walk them side by side.  Without a
            # compile-time program the transposed step is a plain operation list, several times slower per step than
            # the fused forward passes: then it only pays once the chains outnumber the SMs.
            fast = (bool(os.environ.get('QSX_HEISENBERG_ALWAYS')) or
                    getattr(eng, 'has_static_rrc', lambda p: True)(dev_back))
            if fast or chains * n_sets >= 4 * getattr(eng, 'sm_count', 1):
                if n_sets > 1:                          # one read-out per parameter set of this rank, same start
                    sets = torch.arange(set_lo, set_hi, dtype=torch.int32, device=eng.device)
                    _, readouts = eng.evolve(dev_back, start, local_sets, int(emit.max()), emit, snapshots=True,
                                             inst_set=sets, inst_src=torch.zeros_like(sets))
                else:
                    _, snap = eng.evolve(dev_back, start, 1, int(emit.max()), emit, snapshots=True)
                    readouts = snap[0]
    if readouts is None:
        return None
    ready = None
    if side_stream is not None:
        ready = torch.cuda.Event()
        ready.record(side_stream)
    return readouts, ready


def _unique_delays(step_counts, continuous):
    uniq, inverse = [], []
    for c in step_counts:
        u, inv = np.unique(np.asarray(c, dtype=np.float64 if continuous else np.int64), return_inverse=True)
        uniq.append(u if continuous else u.astype(np.int32))
        inverse.append(inv)
    return uniq, inverse


class ResponsePlan:
    """The device work of ``response_function_set`` for ONE (program, dipoles, diagrams, delay grid), captured in a
    CUDA graph: every propagation, dipole application and read-out product of the tree, on the main and the side
    streams, becomes one ``cudaGraphLaunch`` (a C4 evaluation is ~30 launches on 4 streams; enqueueing them from Python
    costs more than their critical path).  The graph ends with the rank-local signal rows; the gather over ranks and
    the reshaping to the caller's grid stay outside it.  Tables the tree uploads are kept resident by the plan."""

    def __init__(self, graph, local, infos, uploads, launches):
        self.graph, self.local, self.infos, self.uploads, self.launches = graph, local, infos, uploads, launches

    @classmethod
    def build(cls, eng, program, mu, sequences, step_counts, global_sets=None):
        """Warm pass (fills the upload memo, loads modules, sizes the allocator), then the captured pass."""
        eng._memo = {}
        try:
            _tree(eng, program, mu, sequences, step_counts, global_sets)
            torch.cuda.synchronize(eng.device)
            before = eng.launches
            graph = torch.cuda.CUDAGraph()
            eng._capturing = True
            with torch.cuda.graph(graph):
                local, infos = _tree(eng, program, mu, sequences, step_counts, global_sets)
            return cls(graph, local, infos, eng._memo, eng.launches - before)
        finally:
            eng._memo, eng._capturing = None, False

    def run(self, eng, device_output: bool, host_out=None):
        """Replay -> list of signals (fresh tensors / arrays: the graph's own output buffers are reused by the next
        replay)."""
        self.graph.replay()
        eng.launches += self.launches
        for k, (out, info) in enumerate(zip(self.local, self.infos)):
            _copy_local_rows(host_out, k, out, info)
        return [_finish(eng, out.clone() if device_output and not info['sharded'] else out, info, device_output)
                for out, info in zip(self.local, self.infos)]


def plan_key(program, mu, sequences, step_counts):
    rank, size = qdist.world()
    return (_fingerprint(program), np.asarray(mu, dtype=float).tobytes(), tuple((a, b) for a, b in sequences),
            tuple(np.asarray(c).tobytes() for c in step_counts), rank, size,
            tuple(os.environ.get(k) for k in ('QSX_DISABLE_BASIS', 'QSX_DISABLE_HEISENBERG', 'QSX_HEISENBERG_MIN',
                                              'QSX_HEISENBERG_ALWAYS', 'QSX_DISABLE_RR', 'QSX_DISABLE_RRC',
                                              'QSX_DISABLE_PULLBACK')))


class HostRows:
    """Pinned host tensors that receive the rank-local signal rows of the diagrams of ``response_function_set`` while
    the tree still runs; ``copied[k]`` tells whether diagram k's copy was started (the caller synchronises, and copies
    the finished signal itself where it was not)."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        self.copied = [False] * len(self.tensors)


def _copy_local_rows(host_out, k, out, info):
    """Start the copy of diagram k's rank-local signal rows into ``host_out.tensors[k]`` on the current stream.  Only rows
    that already are the caller's grid (ascending distinct delays, nothing left to contract) can go out this early."""
    if (host_out is None or host_out.tensors[k] is None or out is None or not out.is_cuda
            or os.environ.get('QSX_DISABLE_EARLY_COPY')):
        return
    plain = all(len(inv) == n and np.array_equal(inv, np.arange(n)) for inv, n in zip(info['inverse'], info['uniq_len']))
    target = host_out.tensors[k]
    if info['zero'] or info.get('pending') is not None or not plain or target.numel() != out.numel():
        return
    target.view(out.shape).copy_(out, non_blocking=True)
    host_out.copied[k] = True


def response_function_set(program: StepProgram, mu, sequences, step_counts, engine=None, device_output: bool = False,
                          global_sets: int | None = None, host_out=None):
    """Several diagrams ``sequences = [(side_seq, direction_seq), ...]`` of the same system on the same delay grid (the
    contributions of one 2-D spectrum).  All distinct read-out propagations are enqueued first, each on its own side
    stream, so the longest one (the two-exciton block of esa) runs while the shorter diagrams are evaluated; the delay
    levels the diagrams have in common are computed once (``shared`` of ``response_function``).  -> list of signals.

    The second evaluation of the same content on a GPU builds a ``ResponsePlan`` (CUDA graph); later ones replay it
    (``QSX_DISABLE_GRAPH`` switches this off).

    ``global_sets``: the parameter sets of ``program`` are THIS RANK's contiguous share (``dist.shard_bounds``) of an
    ensemble of that many members which the caller split over the ranks before lowering anything (so that no rank
    computes tables for members it does not hold); the tree then runs without further sharding and one all-gather per
    diagram assembles all members' signals on every rank.

    ``host_out``: a ``HostRows`` with one pinned complex128 host tensor per diagram (or None entries), each with as
    many elements as this rank's share of the signal.  The rank-local rows of a diagram are copied into it on the
    diagram's own stream as soon as they are complete, i.e. while the other diagrams still run (gsb is finished long
    before esa); the caller synchronises before reading and handles the diagrams whose ``copied`` flag stayed False
    (zero signal, unsorted or duplicate delays, basis rows sharded over ranks)."""
    from .runtime import get_engine
    eng = engine or get_engine()
    continuous = bool(getattr(program, 'continuous', False))
    on_gpu = eng.device.type == 'cuda'
    if on_gpu and not continuous and not os.environ.get('QSX_DISABLE_GRAPH') and hasattr(eng, 'plans'):
        key = plan_key(program, mu, sequences, step_counts) + (global_sets,)
        plan = eng.plans.get(key)
        if plan is None and key not in eng.plans:
            seen = eng.plan_seen[key] = eng.plan_seen.get(key, 0) + 1
            if len(eng.plan_seen) > 64:
                eng.plan_seen.clear()
            if seen >= 2:
                try:
                    plan = ResponsePlan.build(eng, program, mu, sequences, step_counts, global_sets)
                except Exception as exc:                  # not capturable (e.g. the streaming path synchronises): stay eager
                    import warnings
                    warnings.warn('response tree not captured in a CUDA graph: {!r}'.format(exc))
                    plan = None
                if len(eng.plans) >= 8:
                    eng.plans.pop(next(iter(eng.plans)))
                eng.plans[key] = plan
        if plan is not None:
            return plan.run(eng, device_output, host_out)
    local, infos = _tree(eng, program, mu, sequences, step_counts, global_sets, host_out)
    return [_finish(eng, out, info, device_output) for out, info in zip(local, infos)]


def _tree(eng, program, mu, sequences, step_counts, global_sets=None, host_out=None):
    """``_set_local``, or -- for an ensemble the caller already split over the ranks -- ``_set_local`` without any
    sharding inside, with the bookkeeping ``_finish`` needs to gather the members of all ranks."""
    if global_sets is None or qdist.world()[1] == 1:
        return _set_local(eng, program, mu, sequences, step_counts, host_out)
    with qdist.local_only():
        local, infos = _set_local(eng, program, mu, sequences, step_counts, host_out)
    local_sets = n_parameter_sets(program)
    out_rows = []
    for out, info in zip(local, infos):
        if info['zero']:
            out = torch.zeros([global_sets] + info['shape'], dtype=torch.complex128, device=eng.device)
        else:
            info.update(sharded=True, shard_units=global_sets, rows_per_unit=out.shape[0] // local_sets)
        info['lead'] = [global_sets]
        out_rows.append(out)
    return out_rows, infos


def empty_share(eng, sequences, step_counts, global_sets):
    """What a rank WITHOUT members contributes to the gathers of a pre-split ensemble: no rows, the same bookkeeping."""
    uniq, inverse = _unique_delays(step_counts, False)
    rows_per_unit = int(np.prod([len(u) for u in uniq[:-1]])) if len(uniq) > 1 else 1
    info = dict(lead=[global_sets], shape=[len(c) for c in step_counts], uniq_len=[len(u) for u in uniq], inverse=inverse,
                zero=False, sharded=True, shard_units=global_sets, rows_per_unit=rows_per_unit, pending=None)
    empty = torch.zeros((0, len(uniq[-1])), dtype=torch.complex128, device=eng.device)
    return [empty for _ in sequences], [dict(info) for _ in sequences]


def _set_local(eng, program, mu, sequences, step_counts, host_out=None):
    """Enqueue the whole tree of a set of diagrams; -> (rank-local signal rows per diagram, bookkeeping per diagram).
    ``host_out``: see ``response_function_set`` (the copy of a diagram's rows starts on the diagram's stream)."""
    continuous = bool(getattr(program, 'continuous', False))
    uniq, _ = _unique_delays(step_counts, continuous)
    rank, size = qdist.world()
    shared = {}
    pending = []
    for side_seq, direction_seq in sequences:
        plan = interaction_plan(side_seq, direction_seq)
        if len(plan) != len(step_counts) + 1:
            raise ValueError('side_seq must have one more entry than delay_time')
        last_block, chains = _last_level(program, plan, uniq, rank, size)
        key = ('readouts', last_block, uniq[-1].tobytes())
        if key not in shared and key not in [k for k, _ in pending]:
            pending.append((key, (last_block, chains)))
    # longest first: the block size orders the propagation times
    pending.sort(key=lambda item: -(item[1][0].size if item[1][0] is not None else 0))
    on_gpu = eng.device.type == 'cuda' and not os.environ.get('QSX_ONE_STREAM')

    fork = None
    if on_gpu:
        fork = torch.cuda.Event()
        fork.record(torch.cuda.current_stream(eng.device))

    def start_readouts(items, first_index):
        for index, (key, (last_block, chains)) in enumerate(items):
            got = _propagate_readout(eng, program, mu, last_block, chains, uniq[-1], continuous, first_index + index, fork)
            if got is not None:
                shared[key] = got

    if not on_gpu or len(sequences) == 1:
        start_readouts(pending, 0)
        pairs = [_response_local(program, mu, side_seq, direction_seq, step_counts, eng, shared)
                 for side_seq, direction_seq in sequences]
        for k, pair in enumerate(pairs):
            _copy_local_rows(host_out, k, *pair)
        return [p[0] for p in pairs], [p[1] for p in pairs]

    # Order of enqueueing on the GPU (it decides how early the long kernels start when the host is the slower side):
    # the longest read-out propagation, then the forward levels of the heaviest diagram (largest block before its last
    # delay), then the other read-outs and diagrams.  Each diagram after the first runs on its own stream; levels shared
    # between diagrams carry the event of their completion, so e.g. the cheap ground-state level of gsb fills the SMs the
    # long single-exciton level of se / esa leaves idle in its last wave.
    def weight(seq):
        ka = kb = 0
        heavy = 0
        for side, kind in interaction_plan(*seq)[:-2]:
            delta = +1 if kind == 'X' or (kind == 'raise') == (side == 'k') else -1
            ka, kb = (ka + delta, kb) if side == 'k' else (ka, kb + delta)
            blk = Block(program.n_sites, program.n_bits, ka, kb)
            heavy = max(heavy, blk.size if blk.valid() else 0)
        return heavy

    order = sorted(range(len(sequences)), key=lambda k: -weight(sequences[k]))
    main = torch.cuda.current_stream(eng.device)
    streams = [main] + [eng.side_stream_of(4 + (k - 1) % 4) for k in range(1, len(sequences))]
    for st in streams[1:]:
        st.wait_stream(main)                            # fork before anything of this call is enqueued on main
    shared['events'] = True                             # shared entries carry (stream, event) of their producer
    # Read-outs before forward levels: their few CTAs must be resident BEFORE a level fills every SM with CTAs that stay
    # for the whole launch -- enqueued after it, a read-out waits for the first of them to retire.  One system: only the
    # longest read-out first (it is the critical path, the others fit in later).  An ensemble: all of them (its long
    # level is scheduled by segments and leaves no gap until it ends).
    early = len(pending) if n_parameter_sets(program) >= 4 else 1
    start_readouts(pending[:early], 0)
    pairs = [None] * len(sequences)
    for position, (st, k) in enumerate(zip(streams, order)):
        side_seq, direction_seq = sequences[k]
        with torch.cuda.stream(st):
            if position == 0 and len(pending) > early:
                # the first diagram's early levels before the remaining read-outs (which only the products at the very
                # end need): its tree is entered with the read-outs it will find in `shared` later, so split the call
                pair = _response_local(program, mu, side_seq, direction_seq, step_counts, eng, shared,
                                       before_last=lambda: start_readouts(pending[early:], early))
            else:
                pair = _response_local(program, mu, side_seq, direction_seq, step_counts, eng, shared)
            _copy_local_rows(host_out, k, *pair)
            if st is not main and pair[0] is not None:
                pair[0].record_stream(main)
        pairs[k] = pair
    for st in streams[1:]:
        main.wait_stream(st)                            # join
    return [p[0] for p in pairs], [p[1] for p in pairs]


def _publish(eng, shared, key, value, tensors):
    """Store a level in ``shared`` together with the event that marks its completion on the producing stream."""
    if eng.device.type == 'cuda' and shared.get('events'):
        event = torch.cuda.Event()
        event.record(torch.cuda.current_stream(eng.device))
        shared[key] = (value, event, torch.cuda.current_stream(eng.device), tensors)
    else:
        shared[key] = (value, None, None, tensors)


def _consume(eng, shared, key):
    value, event, producer, tensors = shared[key]
    if event is not None:
        here = torch.cuda.current_stream(eng.device)
        if here != producer:
            here.wait_event(event)
            for t in tensors:
                if t is not None and t.is_cuda:
                    t.record_stream(here)
    return value


def response_function(program: StepProgram, mu, side_seq: str, direction_seq: str, step_counts, engine=None,
                      device_output: bool = False, shared: dict | None = None):
    """Signal over the delay grid.  ``step_counts``: one integer list per delay (any order, duplicates allowed); for
    a continuous-time program (``engine.lindblad.LindbladProgram``, the comparison path clspectroscopy.py:18-132) the
    lists hold the delay times themselves.
    Returns a complex ndarray of shape [len(c) for c in step_counts], identical on every rank (``device_output``: a
    complex128 CUDA tensor instead, e.g. for the device post-processing).

    ``shared``: a dict the caller passes to successive calls for diagrams of the SAME system, parameters and delay
    grid (e.g. gsb, se and esa of one 2-D spectrum).  Levels whose interaction prefix and delays coincide (all three
    rephasing diagrams share the t1 level, se and esa also the t2 level) and read-outs of the same final block are
    then computed once and kept on the device in it.  A dict filled for another program raises ``ValueError``.
"""
    from .runtime import get_engine
    eng = engine or get_engine()
    out, info = _response_local(program, mu, side_seq, direction_seq, step_counts, eng, shared)
    return _finish(eng, out, info, device_output)


def _response_local(program: StepProgram, mu, side_seq: str, direction_seq: str, step_counts, eng,
                    shared: dict | None = None, before_last=None):
    """``_response_tree`` with the caller's current stream restored whatever happens inside (the tree may move to a
    branch stream half way)."""
    home = torch.cuda.current_stream(eng.device) if eng.device.type == 'cuda' else None
    try:
        return _response_tree(program, mu, side_seq, direction_seq, step_counts, eng, shared, before_last)
    finally:
        if home is not None and torch.cuda.current_stream(eng.device) != home:
            torch.cuda.set_stream(home)


def _response_tree(program: StepProgram, mu, side_seq: str, direction_seq: str, step_counts, eng,
                   shared: dict | None = None, before_last=None):
    """Enqueue the tree of one diagram; -> (rank-local signal rows, bookkeeping for ``_finish``: gather over the
    ranks, caller's delay order, host copy).  Arguments as ``response_function``."""
    n_sites = program.n_sites
    plan = interaction_plan(side_seq, direction_seq)
    if len(plan) != len(step_counts) + 1:
        raise ValueError('side_seq must have one more entry than delay_time')
    shape = [len(c) for c in step_counts]
    continuous = bool(getattr(program, 'continuous', False))
    uniq, inverse = _unique_delays(step_counts, continuous)

    def upload(block):
        if continuous:
            return eng.upload_lindblad(program.bind(block))
        return eng.upload_program(program.bind(block), program)

    def propagate(dev, sigma, n_local, emit, **kw):
        if continuous:
            from .lindblad import segments_from_times
            return eng.evolve_lindblad(dev, sigma, n_local, segments_from_times(emit), **kw)
        return eng.evolve(dev, sigma, n_local, int(emit.max()), emit, **kw)
    mu_dev = eng.to_device(np.asarray(mu, dtype=np.float64), torch.float64)
    rank, size = qdist.world()

    if shared is not None:
        mark = (_fingerprint(program), np.asarray(mu, dtype=float).tobytes(), rank, size)
        if shared.setdefault('program', mark) != mark:
            raise ValueError('the shared dict was filled for a different system, parameter set or process group')

    def prefix_key(level):
        return ('level', tuple(plan[:level + 1]), tuple(u.tobytes() for u in uniq[:level + 1]))

    # Last delay in the Heisenberg picture: <W, Phi^n sigma> = <(Phi^T)^n W, sigma>.  ONE propagation of the emission
    # read-out W with the transposed step replaces one propagation per chain; the signal is then the product
    # [chains x E] . [E x emissions] (qsx_readout_gemm).  The read-out does not depend on the chains, so it is
    # propagated on a side stream while the earlier delays run.
    readouts = ready = None                             # propagated read-outs and the CUDA event of their completion
    last_block, chains = _last_level(program, plan, uniq, rank, size)
    readout_key = ('readouts', last_block, uniq[-1].tobytes())

    def find_readouts():
        if shared is not None and readout_key in shared:
            return shared[readout_key]
        got = _propagate_readout(eng, program, mu, last_block, chains, uniq[-1], continuous, 0)
        if got is not None and shared is not None:
            shared[readout_key] = got
        return got if got is not None else (None, None)

    if before_last is None:                             # (else: looked up when the last level is reached)
        readouts, ready = find_readouts()

    on_cuda = eng.device.type == 'cuda'
    home = torch.cuda.current_stream(eng.device) if on_cuda else None     # the stream this diagram was given
    start_event = base_sets = None                      # fork point of the branch stream; parameter-set ids at level 0
    branch = None                                       # stream the basis rows continue on (see the substitution)

    ka = kb = 0
    block = program.block(0, 0)
    n_sets = n_parameter_sets(program)
    if n_sets > 1 and continuous:
        raise ValueError('ensembles are evaluated by the collision-model tree only')
    ground = np.zeros((n_sets, block.size), dtype=complex)
    ground[:, 0] = 1.0                                  # |g><g|, pseudomodes and ancilla in |0>
    sigma = eng.to_device(ground)
    inst_set = torch.arange(n_sets, dtype=torch.int32, device=eng.device) if n_sets > 1 else None
    if n_sets > 1:                                      # the sets this rank holds (parameter sets are sharded first)
        set_lo, set_hi = qdist.shard_bounds(n_sets, rank, size) if size > 1 else (0, n_sets)
        base_sets = torch.arange(set_lo, set_hi, dtype=torch.int32, device=eng.device)
    if on_cuda and not os.environ.get('QSX_ONE_STREAM'):
        start_event = torch.cuda.Event()                # fork point: everything a branch stream needs exists by now
        start_event.record(home)
    n_global = n_sets                                   # instances across all ranks at the current level
    shard_units, rows_per_unit = 0, 1                   # units split over ranks, and rows each has grown into
    sharded = False
    zero = False
    out = None
    # structural support of the states of the current level (engine/support.py) and the pending basis substitution
    support = None if continuous else qsupport.ground_support(block)
    pending = None                                      # dict(coeff [S, n_in, d], rows_in, after_gather) once substituted
    inst_src = None
    for level, (side, kind) in enumerate(plan[:-1]):
        # ket side: |1><0| sigma raises the ket sector; bra side: sigma |0><1| raises the bra sector
        delta = +1 if kind == 'X' or (kind == 'raise') == (side == 'k') else -1
        nka, nkb = (ka + delta, kb) if side == 'k' else (ka, kb + delta)
        new_block = Block(n_sites, program.n_bits, nka, nkb)
        if not new_block.valid():
            zero = True
            break
        last = level == len(plan) - 2
        if last and before_last is not None:            # the caller's deferred work (other read-outs), then ours
            before_last()
            before_last = None
            readouts, ready = find_readouts()
        if shared is not None and not last and prefix_key(level) in shared:
            (sigma, n_global, sharded, shard_units, rows_per_unit, inst_set, pending,
             support) = _consume(eng, shared, prefix_key(level))
            block, ka, kb = new_block, nka, nkb
            continue
        # Basis substitution (linearity).  The n_in states per parameter set that enter this level all live on the
        # `d` elements of their structural support: when n_in clearly exceeds d, the d one-hot basis elements are
        # propagated instead and the signal is contracted with the coefficients [n_in, d] at the end.
        inst_src = None
        if (pending is None and level >= 1 and support is not None and sigma.shape[0] > 0
                and not os.environ.get('QSX_DISABLE_BASIS')):
            sup = np.flatnonzero(support.reshape(-1))
            local_sets = sigma.shape[0] * n_sets // n_global if sharded else n_sets
            n_in = sigma.shape[0] // max(local_sets, 1)
            if local_sets > 0 and n_in * 2 >= 3 * len(sup) and (n_sets > 1 or not sharded):
                d = len(sup)
                sup_dev = eng.to_device(sup, torch.int64)
                coeff = sigma.reshape(local_sets, n_in, block.size).index_select(2, sup_dev).contiguous()
                coeff_ready = None
                if start_event is not None:
                    # The basis rows do not depend on the earlier levels (only the coefficients do): they continue on
                    # their own stream, forked at the entry of this diagram, and run beside the earlier levels; the
                    # contraction at the end waits for the coefficients.
                    coeff_ready = torch.cuda.Event()
                    coeff_ready.record(torch.cuda.current_stream(eng.device))
                    eng._branches = getattr(eng, '_branches', 0) + 1
                    branch = eng.side_stream_of(8 + eng._branches % 4)
                    branch.wait_event(start_event)
                    torch.cuda.set_stream(branch)
                    if base_sets is not None:
                        base_sets.record_stream(branch)
                one_hot = np.zeros((d, block.size), dtype=complex)
                one_hot[np.arange(d), sup] = 1.0
                sigma = eng.to_device(one_hot)
                pending = dict(coeff=coeff, n_in=n_in, d=d, after_gather=False, ready=coeff_ready)
                if n_sets > 1:                          # every set propagates the same d source rows with its own tables
                    inst_set = base_sets.repeat_interleave(d)
                    inst_src = torch.arange(d, dtype=torch.int32, device=eng.device).repeat(local_sets)
                    n_global = n_sets * d
                else:
                    n_global = d
        # Chains / basis rows of ONE system are only worth sharding when a level holds more instances than the GPU
        # has SMs (below that every propagation is latency-bound and a rank gains nothing from holding fewer of them):
        # otherwise every rank evaluates the whole tree and no gather is needed.  Parameter sets are always sharded.
        shard_min = int(os.environ.get('QSX_SHARD_MIN', getattr(eng, 'sm_count', 1)))
        worth_it = not (last and readouts is not None) or 'QSX_SHARD_MIN' in os.environ    # (a product only: not worth it)
        if not sharded and size > 1 and (n_sets > 1 or (n_global >= size and n_global >= shard_min and worth_it)):
            lo, hi = qdist.shard_bounds(n_global, rank, size)
            sigma = sigma[lo:hi].contiguous()
            if inst_set is not None:
                inst_set = inst_set[lo:hi].contiguous()
            sharded, shard_units = True, n_global
            if pending is not None:                     # the basis rows of ONE system are what is sharded: contract
                pending = dict(pending, after_gather=True)                                  # after the gather
        n_local = sigma.shape[0] if inst_src is None else int(inst_src.numel())
        emit = uniq[level]
        if n_local == 0:                                # a rank without parameter sets (more ranks than sets)
            block, ka, kb = new_block, nka, nkb
            width = block.size if level < len(plan) - 2 else len(emit)
            sigma = out = torch.zeros((0, width), dtype=torch.complex128, device=eng.device)
            n_global *= len(emit)
            if sharded and level < len(plan) - 2:
                rows_per_unit *= len(emit)
            continue
        # Last interaction before a product with propagated read-outs: <W, O sigma> = <O^T W, sigma>.  When the interaction
        # ENLARGES the block (esa: 64 x 64 -> 96 x 64) the transposed dipole is applied to the few read-outs instead of the
        # many chains -- the collective operator between two sector blocks is real, and its transpose is the same operator
        # with the blocks exchanged -- and the product runs over the smaller block (QSX_DISABLE_PULLBACK: the chains).
        pull_back = (last and readouts is not None and new_block.size > block.size and not continuous
                     and not os.environ.get('QSX_DISABLE_PULLBACK'))
        if not pull_back:
            sigma = eng.apply_dipole(block, new_block, 0 if side == 'k' else 1, mu_dev, sigma)
        prev_block = block
        block, ka, kb = new_block, nka, nkb
        dev = None if last and readouts is not None else upload(block)
        sets_kw = {} if inst_set is None else dict(inst_set=inst_set)
        if inst_src is not None:
            sets_kw['inst_src'] = inst_src
        if level < len(plan) - 2:
            _, snap = propagate(dev, sigma, n_local, emit, snapshots=True, **sets_kw)
            sigma = snap.reshape(n_local * len(emit), block.size)
            if inst_set is not None:
                inst_set = inst_set.repeat_interleave(len(emit))
            inst_src = None
        elif readouts is not None:
            if ready is not None:
                main = torch.cuda.current_stream(eng.device)
                main.wait_event(ready)
                readouts.record_stream(main)
            if inst_src is not None:                    # basis rows shared by the sets: one copy per instance
                sigma = sigma.index_select(0, inst_src.long())
            weights, width = readouts, block.size
            if pull_back:
                width = prev_block.size
                weights = eng.apply_dipole(block, prev_block, 0 if side == 'k' else 1, mu_dev,
                                           readouts.reshape(-1, block.size).contiguous()).reshape(
                                               tuple(readouts.shape[:-1]) + (width,))
            if n_sets > 1:                              # whole sets per rank, equally many chains each: batched product
                local_sets = weights.shape[0]
                out = eng.readout_gemm(sigma.reshape(local_sets, n_local // local_sets, width).contiguous(),
                                       weights.contiguous()).reshape(n_local, len(emit))
            else:
                out = eng.readout_gemm(sigma.reshape(n_local, width).contiguous(), weights.contiguous())
        else:
            obs = eng.upload_observables(*block.emission_observable(np.asarray(mu, dtype=float)), kind='emission', mu=mu)
            out, _ = propagate(dev, sigma, n_local, emit, obs=obs, **sets_kw)
            out = out.reshape(n_local, len(emit))
        n_global *= len(emit)
        if sharded and level < len(plan) - 2:
            rows_per_unit *= len(emit)
        if support is not None:
            support = qsupport.step_closure(program, block, qsupport.dipole_support(support, prev_block, block, side))
        if shared is not None and not last:
            _publish(eng, shared, prefix_key(level), (sigma, n_global, sharded, shard_units, rows_per_unit, inst_set, pending,
                                                      support), [sigma, inst_set, None if pending is None else pending['coeff']])

    if pending is not None and not pending['after_gather'] and not zero and out is not None and out.shape[0] > 0:
        out = contract_basis(eng, out, pending)
    if branch is not None:                              # back to the diagram's own stream, which joins the branch
        torch.cuda.set_stream(home)
        home.wait_stream(branch)
        if out is not None and out.is_cuda:
            out.record_stream(home)
    lead = [n_sets] if n_sets > 1 else []               # ensembles: one signal per parameter set
    info = dict(lead=lead, shape=shape, uniq_len=[len(u) for u in uniq], inverse=inverse, zero=zero, sharded=sharded,
                shard_units=shard_units, rows_per_unit=rows_per_unit,
                # (contracted in _finish, after the streams of the tree have joined: no event to wait for there)
                pending=dict(pending, ready=None) if pending is not None and pending['after_gather'] else None)
    if zero:
        out = torch.zeros(lead + shape, dtype=torch.complex128, device=eng.device)
    return out, info


def contract_basis(eng, out, pending):
    """Rows of ``out`` = (set, basis element b, later delays r) x last delay -> (set, incoming state k, r) x last delay:
    out[s, k, r, e] = sum_b coeff[s, k, b] out[s, b, r, e]  (one batched product of the read-out kernel, K = d)."""
    coeff = pending['coeff']                            # [S, n_in, d]
    if pending.get('ready') is not None and coeff.is_cuda:
        here = torch.cuda.current_stream(eng.device)
        here.wait_event(pending['ready'])
        coeff.record_stream(here)
    n_sets_local, n_in, d = coeff.shape
    n_last = out.shape[1]
    per_basis = out.shape[0] // (n_sets_local * d)
    table = out.reshape(n_sets_local, d, per_basis * n_last).transpose(1, 2).contiguous()      # [S, R, d]
    if hasattr(eng, 'readout_gemm') and eng.device.type == 'cuda':
        full = eng.readout_gemm(coeff, table)                                                    # [S, n_in, R]
    else:
        full = torch.matmul(coeff, table.transpose(1, 2))
    return full.reshape(n_sets_local * n_in * per_basis, n_last)


def _finish(eng, out, info, device_output: bool):
    """Rank-local rows -> the signal on the caller's delay grid (one all-gather when the tree was sharded)."""
    lead, shape, inverse, uniq_len = info['lead'], info['shape'], info['inverse'], info['uniq_len']
    if info['zero']:
        return out if device_output else out.cpu().numpy()
    if info['sharded']:
        out = qdist.all_gather_rows(out, info['shard_units'], info['rows_per_unit'])
    if info.get('pending') is not None:                 # the sharded units were basis rows of one system
        out = contract_basis(eng, out, info['pending'])
    skip = len(lead)
    if device_output:
        table = out.reshape(lead + uniq_len)
        for axis, inv in enumerate(inverse):            # back to the caller's order/duplicates, axis by axis
            if len(inv) != table.shape[skip + axis] or np.any(inv != np.arange(len(inv))):
                table = table.index_select(skip + axis, torch.as_tensor(inv, device=eng.device))
        return table.reshape(lead + shape).contiguous()
    table = (eng.to_host(out) if hasattr(eng, 'to_host') else out.cpu().numpy()).reshape(lead + uniq_len)
    if all(len(inv) == n and np.array_equal(inv, np.arange(len(inv))) for inv, n in zip(inverse, uniq_len)):
        return table                                    # delays were ascending and distinct: nothing to reorder
    if lead:
        return table[np.ix_(np.arange(lead[0]), *inverse)].reshape(lead + shape)
    return table[np.ix_(*inverse)].reshape(shape)
