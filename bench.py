#!/usr/bin/env python
"""bench.py -- headline measurement of the hot path on B200 (contract: DESIGN.md "Measurement").

Workload (BASELINE.json configs[3], "C4", the north-star target): the 2-D electronic spectrum of the 4-chromophore
chain with one 2-level pseudomode per site on the 128 x 16 x 128 (t1, t2, t3) delay grid, the three rephasing diagrams
gsb + se + esa, for MEMBERS static-disorder realisations of the aggregate PER GPU (member 0 is the unperturbed BASELINE
system; 2-D spectra are always averaged over such realisations).  One bench "step" = one evaluation of all members'
signals: 3 x 128 x 16 x 128 grid points per member.  Members are sharded over the ranks (weak scaling) and ONE all-gather
over NCCL assembles every member's signal on every rank -- inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--members M]

metric = 2DES grid points per second (whole job).  `value`: CUDA events around the K evaluations (the response tree as
a CUDA graph + the gather), step program and parameter tables resident in HBM, L2 flushed between evaluations outside
the event pairs.  `e2e`: the public API `qspectroscopy_ensemble` from host objects (new realisations every step, so every
table is lowered and uploaded again) to the signals in pinned host memory.  `--impl reference`: the CPU oracle port
(oracle/csrc/qsx_oracle.c, all host cores) on a bounded sample of the same grid with the grid-average cost per point.
Other BASELINE configurations are reported as extras of the same line: `collision_steps` (configs[1], C2, with its own
roofline), `spectroscopy` (configs[2] C3; C4 as ONE system = the latency of a single spectrum; strong scaling of it).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

HBAR = 0.658212
DT = 0.1 / HBAR
GAMMA = 59.08e-3
METRIC = 'grid_points_per_sec'
UNIT = 'grid points/s'
N_T1, N_T2, N_T3 = 128, 16, 128
DIAGRAMS = ('gsb', 'se', 'esa')
POINTS_PER_MEMBER = len(DIAGRAMS) * N_T1 * N_T2 * N_T3
DISORDER_EV = 0.01
# C2 (extra block)
N_SETS = 4096
N_STEPS = 1000


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def c4_delays(n3=N_T3):
    return [(15 * DT * np.arange(N_T1)).tolist(), (100 * DT * np.arange(N_T2)).tolist(), (15 * DT * np.arange(n3)).tolist()]


def c4_member(delta=None):
    """4-site chain, eps = [1.55, 1.50, 1.46, 1.52], J = -0.01 between neighbours, mu = 1, one 2-level
    pseudomode per site (omega 0.02, Gamma 59.08e-3, Omega 0.1, g = sqrt(Gamma Omega / 2), rate 0.2); ``delta``: site-energy
    shifts of a disorder realisation."""
    import warnings
    from qsextra_b200 import ChromophoreSystem
    J = np.zeros((4, 4))
    for i in range(3):
        J[i, i + 1] = J[i + 1, i] = -0.01
    eps = np.array([1.55, 1.50, 1.46, 1.52]) + (0.0 if delta is None else np.asarray(delta))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return ChromophoreSystem(energies=eps.tolist(), couplings=J, dipole_moments=[1., 1., 1., 1.],
                                 frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=float(np.sqrt(GAMMA * 0.1 / 2)))


def c4_members(n_total, seed=0):
    """Global ensemble: member 0 (with seed 0) is the unperturbed system, the others carry Gaussian site-energy disorder."""
    rng = np.random.default_rng(seed)
    deltas = rng.normal(0.0, DISORDER_EV, (n_total, 4))
    if seed == 0:
        deltas[0] = 0.0
    return [c4_member(d) for d in deltas]


def c4_specs(n3=N_T3):
    from qsextra_b200.spectroscopy import FeynmanDiagram
    return [FeynmanDiagram(lab, c4_delays(n3)) for lab in DIAGRAMS]


def workload_config(n_gpus, members):
    return dict(workload='C4: 2DES of the 4-chromophore chain with one 2-level pseudomode per site, {} x {} x {} (t1, t2, t3) '
                         'grid, gsb + se + esa, {} static-disorder realisation(s) of the aggregate per GPU (member 0 = the '
                         'BASELINE system)'.format(N_T1, N_T2, N_T3, members),
                grid=[N_T1, N_T2, N_T3], diagrams=list(DIAGRAMS), members_per_gpu=members, global_members=members * n_gpus,
                grid_points_per_member=POINTS_PER_MEMBER, disorder_sigma_eV=DISORDER_EV,
                sharding='members over ranks (contiguous), one all_gather_into_tensor of the signals inside the timed region',
                cache='L2 flushed (256 MiB write) between timed evaluations, outside the event pairs',
                reference_arm_sample='rows of the same grid with the grid-average cost per point, one row per host thread '
                                     'and bench step (see cpu_baseline.sample)')


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port): a bounded sample with the grid-average cost per point
# ------------------------------------------------------------------------------------------------
CPU_ROW = (16, 2)           # n1 = 240, n2 = 200 steps
CPU_T3 = 32                 # first 32 t3 values: 465 more steps -> 905 steps per 32 points = 28.3 steps per point;
                            # the oracle's algorithm over the full grid: (952 + 750 + 1905) / 128 = 28.2 steps per point


def cpu_rows(n_threads):
    return [(CPU_ROW[0] + (k % 5) - 2, CPU_ROW[1]) for k in range(n_threads)]


def cpu_sample(diagrams=DIAGRAMS, n_threads=None):
    """Time the C oracle (full-space literal gates, one (t1, t2) row per thread, shared prefix along t3) on
    ``n_threads`` rows per diagram.  -> (grid points/s, threads, seconds, description)."""
    from oracle import fast
    from qsextra_b200.spectroscopy import FeynmanDiagram
    n_threads = n_threads or os.cpu_count() or 1
    system = c4_member()
    rows = cpu_rows(n_threads)
    points, used = 0, 1
    t0 = time.perf_counter()
    for lab in diagrams:
        sig, used = fast.response_points(system, FeynmanDiagram(lab, c4_delays(CPU_T3)), rows, DT, coll_rates=[0.2],
                                         kraus=True, n_threads=n_threads)
        points += sig.size
    sec = time.perf_counter() - t0
    what = ('{} rows (t1 index {}..{}, t2 index {}) x first {} t3 values x {} of the C4 grid: {} points, ~905 collision '
            'steps per 32 points = the grid-average 28 steps per point of the oracle algorithm (one chain per (t1, t2), '
            'prefix shared along t3); C oracle port, full-space literal gate list with the collisions as channels, one '
            'row per thread, {:.1f} s'.format(len(rows), CPU_ROW[0] - 2, CPU_ROW[0] + 2, CPU_ROW[1], CPU_T3,
                                            '+'.join(diagrams), points, sec))
    return points / sec, used, sec, what


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_threads = os.cpu_count() or 1
    cpu_sample(('gsb',), min(2, n_threads))                    # load the library, touch the code paths
    for w in range(args.warmup):
        cpu_sample((DIAGRAMS[w % 3],), n_threads)
    t0 = time.perf_counter()
    points = 0.0
    used, what = 1, ''
    for k in range(args.steps):
        rate, used, sec, what = cpu_sample((DIAGRAMS[k % 3],), n_threads)
        points += rate * sec
    sec = time.perf_counter() - t0
    value = points / sec
    line = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * sec / max(args.steps, 1), higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f64', data='synthetic', config=workload_config(args.gpus, args.members),
                cpu_baseline=dict(value=value, unit=UNIT, cores=used, kind='port',
                                  sample='per bench step (diagram rotating gsb, se, esa): ' + what),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm helpers
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling (every 20 ms) started BEFORE the warm-up; `stop(t0, t1)` keeps the samples whose timestamps
    fall inside the timed region [t0, t1] (wall clock) and, when the region is shorter than the sampling period, the
    samples taken under the same load right before it."""
    QUERY = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def wait_for_first_sample(self, timeout=3.0):
        t_end = time.time() + timeout
        while self.proc is not None and time.time() < t_end:
            if os.path.exists(self.path) and os.path.getsize(self.path) > 0:
                return
            time.sleep(0.02)

    def stop(self, t0, t1):
        import datetime
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.05)
        self.proc.terminate()
        self.proc.wait()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = []
        for row in open(self.path):
            f = [x.strip() for x in row.split(',')]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                rows.append((ts, float(f[1]), float(f[2]), [n for n, flag in zip(names, f[4:8]) if flag.lower().startswith('active')]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = [r for r in rows if t0 <= r[0] <= t1]
        where = 'timed region'
        if len(inside) < 3:                        # region shorter than a few sampling periods: add the warm-up samples
            inside = [r for r in rows if t0 - 0.5 <= r[0] <= t1 + 0.05]
            where = 'timed region and the 0.5 s of warm-up before it (region shorter than 3 sampling periods)'
        sm = [r[1] for r in inside]
        reasons = sorted({n for r in inside for n in r[3]})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max((r[2] for r in inside), default=None),
                    samples=len(inside), window=where, reasons=reasons)


def measured_peaks():
    try:
        return json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        return {}


def ncu_traffic(kernel_key):
    """dram bytes per launch of the dominant kernel from the committed ncu capture of this workload (profiles/), or
    None when the capture of the current round does not hold it."""
    try:
        table = json.load(open(os.path.join(REPO, 'profiles', 'ncu_traffic.json')))
    except (OSError, ValueError):
        return None, None
    hit = table.get(kernel_key)
    return (hit.get('dram_bytes_per_launch'), hit.get('source')) if hit else (None, None)


def ncu_capture(kernel_key):
    """Everything profiles/ncu_traffic.json holds about a kernel (per-instance figures of its committed ncu capture)."""
    try:
        return json.load(open(os.path.join(REPO, 'profiles', 'ncu_traffic.json'))).get(kernel_key) or {}
    except (OSError, ValueError):
        return {}


def survey_flops_per_step(n_sites, n_modes, stored_elements, n_bonds):
    """F_step = E (6 n_Z + 12 n_R + 2.5 n_AD) for the d = 2 pseudomode step (un-fused gate list)."""
    n_z = n_sites + n_sites * n_modes
    n_r = 2 * n_bonds + 2 * n_sites * n_modes
    n_ad = n_sites * n_modes
    return stored_elements * (6 * n_z + 12 * n_r + 2.5 * n_ad)


def trace_table(trace, t_ref):
    """[(name, meta, start, stop)] -> per-launch rows and per-kernel sums (device times from the CUDA events)."""
    rows = []
    for name, meta, start, stop in trace:
        ms = start.elapsed_time(stop)
        flops = meta.get('flops')
        if flops is None and 'executed_flops_per_step' in meta:
            flops = float(meta['executed_flops_per_step']) * meta['instances'] * meta['steps']
        rows.append(dict(kernel=name, start_ms=t_ref.elapsed_time(start), ms=ms, executed_flops=flops,
                         **{k: v for k, v in meta.items() if k in ('instances', 'steps', 'elements', 'threads',
                                                                    'fp64_instructions_per_step', 'executed_flops_per_step')}))
    sums = {}
    for r in rows:
        s = sums.setdefault(r['kernel'], dict(launches=0, ms=0.0, executed_flops=0.0, fp64_instructions=0.0))
        s['launches'] += 1
        s['ms'] += r['ms']
        s['executed_flops'] += r['executed_flops'] or 0.0
        if 'fp64_instructions_per_step' in r:
            s['fp64_instructions'] += float(r['fp64_instructions_per_step']) * r['instances'] * r['steps']
    return rows, sums


# ------------------------------------------------------------------------------------------------
# extras: BASELINE configs[1] (C2, collision steps/s) with its own roofline
# ------------------------------------------------------------------------------------------------
def c2_parameters(n_sets, seed=0):
    """eps~U(1.4,1.6), J~U(-.05,.05), omega~U(.01,.2), Gamma~U(.02,.1), Omega~U(.05,.2),
    g = sqrt(Gamma Omega / 2), rate = 2 Omega."""
    rng = np.random.default_rng(seed)
    eps = rng.uniform(1.4, 1.6, (n_sets, 2))
    J = rng.uniform(-0.05, 0.05, n_sets)
    omega = rng.uniform(0.01, 0.2, n_sets)
    Gamma = rng.uniform(0.02, 0.1, n_sets)
    Omega = rng.uniform(0.05, 0.2, n_sets)
    coup = np.zeros((n_sets, 2, 2))
    coup[:, 0, 1] = coup[:, 1, 0] = J
    return dict(energies=eps, couplings=coup, dipole_moments=np.ones((n_sets, 2)), omega=omega[:, None],
                coupl_ep=np.sqrt(Gamma * Omega / 2)[:, None], rates=(2 * Omega)[:, None])


def collision_steps_block(eng, world, rank, barrier, fp64_peak, sm_clock_hz, repeats=20):
    """configs[1]: 4096 parameter sets per GPU x 1000 collision steps of the non-Markovian dimer, populations after
    every step.  Device-timed rate, end-to-end rate through ``qevolve_batch`` and the roofline of its kernel."""
    import torch
    import torch.distributed as dist
    from qsextra_b200.engine.program import StepProgram, SystemBatch
    from qsextra_b200.engine.sectors import population_observables
    from qsextra_b200.qcomo import qevolve_batch
    params = c2_parameters(N_SETS * world)
    lo, hi = rank * N_SETS, (rank + 1) * N_SETS
    batch = SystemBatch(params['energies'][lo:hi], params['couplings'][lo:hi], params['dipole_moments'][lo:hi],
                        params['omega'][lo:hi], params['coupl_ep'][lo:hi], (2,))
    program = StepProgram(batch, DT, 1, 1, params['rates'][lo:hi])
    block = program.block(1, 1)
    bound = program.bind(block)
    dev = eng.upload_program(bound, program)
    sigma0 = torch.zeros((1, block.size), dtype=torch.complex128, device=eng.device)
    sigma0[0, 0] = 1.0                                          # |site 0 excited, modes 0>
    inst_set = torch.arange(N_SETS, dtype=torch.int32, device=eng.device)
    inst_src = torch.zeros(N_SETS, dtype=torch.int32, device=eng.device)
    emit = eng.to_device(np.arange(N_STEPS + 1), torch.int32)
    obs = eng.upload_observables(*population_observables(block), kind='populations')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)

    def hot_path():
        return eng.evolve(dev, sigma0, N_SETS, N_STEPS, emit, obs=obs, obs_real=True, inst_set=inst_set, inst_src=inst_src)[0]

    for _ in range(5):
        flush.fill_(1)
        hot_path()
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(repeats)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(repeats)]
    for i in range(repeats):
        flush.fill_(i & 0xff)
        starts[i].record()
        hot_path()
        stops[i].record()
    barrier()
    kernel_ms = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    t_dev = torch.tensor([sum(kernel_ms)], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    value = repeats * N_SETS * N_STEPS * world / (float(t_dev.item()) * 1e-3)
    # end to end: host arrays in, pinned host populations out, in chunks so that read-back overlaps the next kernel
    pinned = torch.empty((N_SETS, N_STEPS + 1, 2), dtype=torch.float64).pin_memory()
    host_batch = SystemBatch(np.array(batch.energies), np.array(batch.couplings), np.array(batch.dipole_moments),
                             np.array(batch.omega), np.array(batch.coupl_ep), (2,))
    time_grid = np.arange(N_STEPS + 1) * DT

    def e2e_once():
        res = qevolve_batch(host_batch, time_grid, dt=DT, coll_rates=params['rates'][lo:hi], device_output=True, shard=False)
        pinned.copy_(res, non_blocking=True)
        torch.cuda.synchronize()

    e2e_once()
    e2e_once()
    h2d0 = eng.h2d_bytes
    barrier()
    t0 = time.perf_counter()
    iters = 5
    for _ in range(iters):
        e2e_once()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = iters * N_SETS * N_STEPS * world / float(t_e2e.item())
    out = dict(metric='collision_steps_per_sec', value=value, unit='steps/s', kernel_ms_per_launch=float(np.mean(kernel_ms)),
               workload='C2: non-Markovian dimer, 1 two-level pseudomode/site, {} steps, {} parameter sets per GPU, populations '
                        'after every step'.format(N_STEPS, N_SETS),
               e2e=dict(value=e2e_value, unit='steps/s', h2d_bytes_per_step=int((eng.h2d_bytes - h2d0) // iters),
                        d2h_bytes_per_step=int(pinned.numel() * 8),
                        api='qevolve_batch(SystemBatch of host arrays) -> pinned host tensor'))
    if rank == 0 and 'rr' in dev:
        ex_flops, ex_instr = dev['rr'].executed_flops_per_step()
        per_ms = float(np.mean(kernel_ms))
        ex_tf = ex_flops * N_SETS * N_STEPS / (per_ms * 1e-3) / 1e12
        f_step = survey_flops_per_step(2, 1, block.size, 1)
        traffic, source = ncu_traffic('rr::evolve_rr_static_kernel<0,1>')
        out['roofline'] = dict(bound='fp64', kernel='rr::evolve_rr_static_kernel<0, 1>', achieved=ex_tf, peak=fp64_peak,
                               unit='TFLOP/s', frac=ex_tf / fp64_peak if fp64_peak else None,
                               note='executed flops (operation list of the register-resident kernel, read-outs excluded)',
                               fp64_pipe_utilisation_estimate=ex_instr * N_SETS * N_STEPS / (per_ms * 1e-3) / (64 * eng.sm_count * sm_clock_hz),
                               survey_formula=dict(flops_per_collision_step=f_step,
                                                   achieved=f_step * N_SETS * N_STEPS / (per_ms * 1e-3) / 1e12,
                                                   note='un-fused gate list of SURVEY 8(d); the kernel executes fewer flops'),
                               traffic=traffic, traffic_source=source)
    return out, (params, lo, hi, hot_path)


def c2_parity(params, hot_path, n_check=256):
    """The kernel's populations of the first ``n_check`` parameter sets against the C oracle."""
    import warnings
    from oracle import fast
    from qsextra_b200 import ChromophoreSystem
    systems, rates = [], []
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for b in range(n_check):
            s = ChromophoreSystem(energies=params['energies'][b].tolist(), couplings=float(params['couplings'][b, 0, 1]),
                                  dipole_moments=[1., 1.], frequencies_pseudomode=float(params['omega'][b, 0]),
                                  levels_pseudomode=2, couplings_ep=float(params['coupl_ep'][b, 0]))
            s.set_state('localized excitation', 0)
            systems.append(s)
            rates.append([float(params['rates'][b, 0])])
    pops, _ = fast.populations_batch(systems, np.arange(N_STEPS + 1) * DT, dt=DT, coll_rates=np.array(rates),
                                     n_threads=os.cpu_count() or 1)
    got = hot_path()[:n_check].cpu().numpy()
    return float(np.abs(got - pops).max())


# ------------------------------------------------------------------------------------------------
# extras: the other spectroscopy configurations, C4 as ONE system (latency), strong scaling of it
# ------------------------------------------------------------------------------------------------
def spectroscopy_extras(world, rank, barrier):
    import warnings
    import torch
    from qsextra_b200 import ExcitonicSystem
    from qsextra_b200.spectroscopy import FeynmanDiagram, Spectroscopy, qspectroscopy, qspectroscopy_set
    from qsextra_b200.spectroscopy.postprocessing import postprocessing_all_T2
    out = {}

    def timed(run, repeats=5):
        """Median wall-clock of ``repeats`` evaluations (each between barriers) and the last result."""
        samples, result = [], None
        for _ in range(repeats):
            barrier()
            t_start = time.perf_counter()
            result = run()
            barrier()
            samples.append(time.perf_counter() - t_start)
        return sorted(samples)[len(samples) // 2], samples, result

    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        try:
            dimer = ExcitonicSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 1.])
            grid = (30 * DT * np.arange(64)).tolist()
            delays = [grid, [0.], grid]
            specs = [FeynmanDiagram(lab, delays) for lab in DIAGRAMS]
            for sides, directions in (('kkkk', 'ioio'), ('kbbk', 'iioo'), ('kbkk', 'iiio')):
                extra = Spectroscopy(delays)
                extra.side_seq, extra.direction_seq = sides, directions
                specs.append(extra)
            kw = dict(dt=DT, coll_rates=GAMMA / 4)
            for _ in range(3):
                qspectroscopy_set(dimer, specs, verbose=False, **kw)
            sec, samples, sigs = timed(lambda: qspectroscopy_set(dimer, specs, verbose=False, **kw))
            points = len(specs) * 64 * 64
            out['C3_dimer_64x1x64_6pathways'] = dict(grid_points=points, seconds=sec, grid_points_per_s=points / sec,
                                                     api='qspectroscopy_set', samples_s=samples,
                                                     checksum=float(sum(np.abs(s_).sum() for s_ in sigs)))
        except Exception as exc:                                              # keep the headline line alive
            out['C3_dimer_64x1x64_6pathways'] = dict(error=repr(exc))
        try:
            chain = c4_member()
            specs = c4_specs()
            kw = dict(dt=DT, coll_rates=[0.2])
            t_first = time.perf_counter()
            qspectroscopy_set(chain, specs, verbose=False, **kw)
            first = time.perf_counter() - t_first
            for _ in range(2):
                qspectroscopy_set(chain, specs, verbose=False, **kw)
            sec, samples, sigs = timed(lambda: qspectroscopy_set(chain, specs, verbose=False, **kw))
            entry = dict(grid_points=POINTS_PER_MEMBER, seconds=sec, grid_points_per_s=POINTS_PER_MEMBER / sec,
                         api='qspectroscopy_set (host objects -> NumPy signals); ONE system is latency-bound (64 basis rows on 148 SMs): '
                             'under torchrun every rank evaluates it whole (replicas), more GPUs serve more realisations', samples_s=samples, first_call_s=first,
                         checksum=float(sum(np.abs(s_).sum() for s_ in sigs)))

            def pipeline():
                signals = qspectroscopy_set(chain, specs, verbose=False, device_output=True, **kw)
                acc = torch.zeros((), dtype=torch.float64, device='cuda')
                for sp, sig in zip(specs, signals):
                    acc += postprocessing_all_T2(sp, sig, RF_freq=1.5, damping_rate=0.005, pad_extension=3)[1].abs().sum()
                return float(acc)

            pipeline()
            sec, samples, total = timed(pipeline)
            entry['with_all_48_spectra'] = dict(seconds=sec, samples_s=samples, spectra=3 * N_T2, checksum=total,
                                                grid_points_per_s=POINTS_PER_MEMBER / sec,
                                                api='qspectroscopy_set(device_output=True) + postprocessing_all_T2 '
                                                    '(zero-padded 384 x 384 2-D spectra, cuFFT)')
            out['C4_one_system_128x16x128'] = entry
        except Exception as exc:
            out['C4_one_system_128x16x128'] = dict(error=repr(exc))
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def c4_parity(signals_dev, members):
    """The timed evaluation's signals against the C oracle on whole t3 rows: member 0 (the BASELINE system) on the far
    corners and random interior rows, member 1 (a disordered realisation) on two rows.  Runs the six oracle calls
    concurrently (the library releases the GIL).  -> dict."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import fast
    rng = np.random.default_rng(7)
    rows0 = [(127, 15), (127, 0), (0, 15), (0, 0)] + [(int(a), int(b)) for a, b in zip(rng.integers(1, 127, 4), rng.integers(1, 15, 4))]
    rows1 = [(127, 15), (int(rng.integers(1, 127)), int(rng.integers(1, 15)))]
    specs = c4_specs()
    jobs = []
    for d, spec in enumerate(specs):
        jobs.append((d, 0, rows0))
        if len(members) > 1:
            jobs.append((d, 1, rows1))

    def run(job):
        d, m, rows = job
        ref, _ = fast.response_points(members[m], specs[d], rows, DT, coll_rates=[0.2], kraus=True, n_threads=len(rows))
        return ref

    t0 = time.perf_counter()
    with ThreadPoolExecutor(len(jobs)) as pool:
        refs = list(pool.map(run, jobs))
    sec = time.perf_counter() - t0
    worst_rel, worst_abs, n_points, worst_local = 0.0, 0.0, 0, 0.0
    for (d, m, rows), ref in zip(jobs, refs):
        sig = signals_dev[d][m] if signals_dev[d].dim() == 4 else signals_dev[d]
        scale = float(sig.abs().max())
        got = np.array([sig[i, j].cpu().numpy() for i, j in rows])
        err = np.abs(got - ref)
        worst_abs = max(worst_abs, float(err.max()))
        worst_rel = max(worst_rel, float(err.max()) / scale)
        worst_local = max(worst_local, float((err.max(axis=1) / np.abs(ref).max(axis=1)).max()))
        n_points += ref.size
    return dict(max_rel_err=worst_rel, max_abs_err=worst_abs, max_err_relative_to_each_row=worst_local, points=n_points,
                rows_member0=rows0, rows_member1=rows1 if len(members) > 1 else [],
                includes=['(127,15,127)', '(127,0,0)', '(0,15,127)', '(0,0,0)'], oracle_seconds=sec,
                note='whole t3 rows (128 points each) per diagram; relative = max |gpu - oracle| / max |signal of the diagram|')


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank), timeout=datetime.timedelta(seconds=180))
    from qsextra_b200.engine import response as R
    from qsextra_b200.engine.runtime import get_engine
    from qsextra_b200.spectroscopy import qspectroscopy_ensemble
    from qsextra_b200.spectroscopy.simulation.qspectroscopy import ensemble_program

    eng = get_engine()
    M = args.members
    members = c4_members(M * world)                             # global ensemble; ranks own contiguous slices of it
    specs = c4_specs()
    kw = dict(dt=DT, coll_rates=[0.2])
    # every rank lowers its own members only; the tree runs unsharded on them and one all-gather per diagram follows
    share = (rank * M, (rank + 1) * M) if world > 1 else None
    global_sets = M * world if world > 1 else None
    program, mu, sequences, counts = ensemble_program(members, specs, members=share, **kw)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)

    def hot_path():
        """All members' signals on the device: the response tree (CUDA graph from the third call on) + the gather."""
        return R.response_function_set(program, mu, sequences, counts, engine=eng, device_output=True, global_sets=global_sets)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.wait_for_first_sample()
    # at least W (>= 3) steps and half a second under load before timing; the SAME number of evaluations on every rank
    # (each holds a collective): rank 0 decides how many extra ones fill the half second
    signals = None
    t_warm = time.time()
    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        signals = hot_path()
    torch.cuda.synchronize()
    per_step = (time.time() - t_warm) / max(args.warmup, 3)
    extra = torch.tensor([int(np.ceil(max(0.0, 0.5 - (time.time() - t_warm)) / max(per_step, 1e-3)))], device=eng.device)
    if world > 1:
        dist.broadcast(extra, 0)
    for _ in range(int(extra.item())):
        flush.fill_(1)
        signals = hot_path()
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = eng.launches
    barrier()
    wall0 = time.perf_counter()
    t_region0 = time.time()
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        starts[i].record()
        signals = hot_path()
        stops[i].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launches - launches0
    clocks = sampler.stop(t_region0, time.time())
    step_ms = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    t_dev = torch.tensor([sum(step_ms)], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())
    points_per_step = POINTS_PER_MEMBER * M * world
    value = args.steps * points_per_step / (total_ms * 1e-3)
    uses_graph = any(p is not None for p in eng.plans.values())

    # ---- one traced evaluation: device time of every launch (CUDA events on the launching streams), the gather ----
    eng.trace = []
    t_ref = torch.cuda.Event(enable_timing=True)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_ref.record()
    local_rows, infos = R._tree(eng, program, mu, sequences, counts, global_sets)
    g0.record()
    traced = [R._finish(eng, o, info, True) for o, info in zip(local_rows, infos)]
    g1.record()
    barrier()
    trace, eng.trace = eng.trace, None
    launch_rows, kernel_sums = trace_table(trace, t_ref)
    traced_ms = t_ref.elapsed_time(g1)
    gather_ms = g0.elapsed_time(g1)
    for _ in range(5 if world > 1 else 0):             # the gather alone, ranks aligned: the smallest of five
        barrier()
        g0.record()
        again = [R._finish(eng, o, info, True) for o, info in zip(local_rows, infos)]
        g1.record()
        torch.cuda.synchronize()
        gather_ms = min(gather_ms, g0.elapsed_time(g1))
        del again
    t_g = torch.tensor([gather_ms], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t_g, op=dist.ReduceOp.MAX)
    gather_ms = float(t_g.item())
    del traced
    # ... and one evaluation with every launch on ONE stream: the duration of each kernel running alone (in the step the
    # kernels of different streams share the SMs, which stretches each of them)
    os.environ['QSX_ONE_STREAM'] = '1'
    eng.trace = []
    t_alone = torch.cuda.Event(enable_timing=True)
    barrier()
    t_alone.record()
    alone_rows, alone_infos = R._tree(eng, program, mu, sequences, counts, global_sets)
    barrier()
    os.environ.pop('QSX_ONE_STREAM')
    alone_trace, eng.trace = eng.trace, None
    _, alone_sums = trace_table(alone_trace, t_alone)
    del alone_rows, alone_infos

    # ---- end to end through the public API: host objects in (new realisations every step), pinned host signals out ---
    lo, hi = rank * M, (rank + 1) * M
    pinned = [torch.empty((M, N_T1, N_T2, N_T3), dtype=torch.complex128).pin_memory() for _ in DIAGRAMS]

    def e2e_once(seed):
        fresh = c4_members(M * world, seed=seed)
        # every rank reads back its own members (together the complete result is in host memory once); the copy of a
        # diagram starts when the diagram is complete, beside the diagrams that still run; returns after the copies
        qspectroscopy_ensemble(fresh, specs, shots=0, verbose=False, device_output=True, host_out=pinned, **kw)

    e2e_once(99)
    e2e_once(100)
    h2d0 = eng.h2d_bytes
    e2e_iters = max(1, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_iters):
        e2e_once(101 + k)
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = e2e_iters * points_per_step / float(t_e2e.item())
    h2d = (eng.h2d_bytes - h2d0) // e2e_iters

    fp64_peak = eng.fp64_peak(5) if rank == 0 else None
    sm_clock = (clocks.get('sm_mhz') or 1965.0) * 1e6
    fp64_peak_all = torch.tensor([fp64_peak or 0.0], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.broadcast(fp64_peak_all, 0)
    extras_c2 = extras_spec = None
    c2_state = None
    if not args.no_extras:
        extras_c2, c2_state = collision_steps_block(eng, world, rank, barrier, float(fp64_peak_all.item()), sm_clock)
        extras_spec = spectroscopy_extras(world, rank, barrier)

    if rank == 0:
        peaks = measured_peaks()
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        # dominant kernel of the step = largest device time in the traced evaluation
        dom_name = max(kernel_sums, key=lambda k: kernel_sums[k]['ms'])
        dom = kernel_sums[dom_name]
        dom_rows = [r for r in launch_rows if r['kernel'] == dom_name]
        dom_tf = dom['executed_flops'] / (dom['ms'] * 1e-3) / 1e12 if dom['ms'] else 0.0
        total_flops = sum(s['executed_flops'] for s in kernel_sums.values())
        busy_ms = sum(s['ms'] for s in kernel_sums.values())
        mean_step_ms = float(np.mean(step_ms))
        capture = ncu_capture(dom_name)
        dom_instances = sum(r['instances'] for r in dom_rows) / max(1, dom['launches'])
        traffic = traffic_source = data_pipe = None
        if capture.get('dram_bytes_per_instance'):
            traffic, traffic_source = int(capture['dram_bytes_per_instance'] * dom_instances), capture.get('source')
        if capture.get('lsu_wavefronts_per_instance_step') and dom['ms']:
            # the pipe that bounds this kernel: one wavefront of warp shuffles / shared-memory accesses per SM and clock
            per = capture['lsu_wavefronts_per_instance_step']
            rate = per * sum(r['instances'] * r['steps'] for r in dom_rows if r.get('steps') is not None) / (dom['ms'] * 1e-3)
            peak_rate = eng.sm_count * sm_clock
            data_pipe = dict(bound='shared-memory / shuffle data pipe (l1tex__data_pipe_lsu_wavefronts)', achieved=rate / 1e9,
                             peak=peak_rate / 1e9, unit='Gwavefronts/s', frac=rate / peak_rate,
                             wavefronts_per_instance_step=per,
                             ncu_pct_of_peak_sustained_elapsed=capture.get('lsu_data_pipe_pct_of_elapsed_cycles'),
                             note='wavefronts per instance-step from the committed ncu capture of this kernel x the instance-steps '
                                  'of this run / its device time here; peak = 1 wavefront per SM and clock at the sampled SM '
                                  'clock.  This, not FP64 issue, is what the kernel runs against (profiles/r2_c4_rrc3_kernel.md)')
        alone = None
        if alone_sums.get(dom_name, {}).get('ms'):
            a_ms = alone_sums[dom_name]['ms']
            alone = dict(kernel_ms_per_step=a_ms, achieved=dom['executed_flops'] / (a_ms * 1e-3) / 1e12,
                         frac=dom['executed_flops'] / (a_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                         data_pipe_frac=None if data_pipe is None else data_pipe['frac'] * dom['ms'] / a_ms,
                         note='the same launches with the whole evaluation on one stream (no other kernel shares the SMs): '
                              'what ncu measures; the figures above are taken inside the multi-stream step')
        blk = program.block(1, 1)
        f_step = survey_flops_per_step(4, 1, blk.size, 3)
        inst_steps = sum(r['instances'] * r['steps'] for r in dom_rows if r.get('steps') is not None)
        roofline = dict(
            bound='fp64', kernel=dom_name + ' (CTA-wide register-resident kernel: t2 level of se / esa, all members)',
            achieved=dom_tf, peak=fp64_peak, unit='TFLOP/s', frac=dom_tf / fp64_peak if fp64_peak else None,
            traffic=traffic, traffic_source=traffic_source, data_pipe=data_pipe, alone=alone,
            what='executed FP64 flops of the kernel (counted from its operation list, RRCProgram.executed_flops_per_step) '
                 '/ its launch durations (CUDA events on the launching stream, traced evaluation of the same step)',
            peak_source='qsx_fp64_peak DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure)',
            launches_per_step=dom['launches'], kernel_ms_per_step=dom['ms'], share_of_step=dom['ms'] / mean_step_ms,
            instance_steps_per_step=inst_steps,
            fp64_pipe_utilisation_estimate=dom['fp64_instructions'] / (dom['ms'] * 1e-3) / (64 * eng.sm_count * sm_clock) if dom['ms'] else None,
            survey_formula=dict(flops_per_collision_step=f_step,
                                achieved=f_step * inst_steps / (dom['ms'] * 1e-3) / 1e12 if dom['ms'] else None,
                                note='SURVEY 8(d) counts the un-fused gate list on the (1,1) block; the kernel executes '
                                     'fewer flops by exact algebra, so this figure can exceed the peak'),
            whole_step=dict(executed_flops=total_flops, ms=mean_step_ms, achieved=total_flops / (mean_step_ms * 1e-3) / 1e12,
                            frac=total_flops / (mean_step_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                            kernel_busy_ms_summed_over_streams=busy_ms, traced_step_ms=traced_ms,
                            note='all launches of one evaluation (evolutions, dipoles, read-out products)'),
            kernels={k: dict(launches=v['launches'], ms=v['ms'],
                             tflops=v['executed_flops'] / (v['ms'] * 1e-3) / 1e12 if v['ms'] and v['executed_flops'] else None)
                     for k, v in sorted(kernel_sums.items(), key=lambda kv: -kv[1]['ms'])},
            hbm=dict(peak=hbm_peak, unit='GB/s', peak_source='MEASURED_PEAKS.json' if peaks else 'fallback',
                     note='the states stay on chip during a propagation; HBM carries the snapshots between levels'))
        cpu_baseline = parity = c2_err = None
        if world == 1:                                  # reported on rank 0 at N = 1 only
            rate, used, sec, what = cpu_sample()
            cpu_baseline = dict(value=rate, unit=UNIT, cores=used, kind='port', sample=what)
            parity = c4_parity(signals, members)
            if c2_state is not None:
                c2_err = c2_parity(c2_state[0], c2_state[3])
                extras_c2['parity_max_abs_err_vs_oracle'] = c2_err
                extras_c2['parity_sets_checked'] = 256
        if args.trace_out:
            json.dump(dict(launches=launch_rows, traced_step_ms=traced_ms, gather_ms=gather_ms), open(args.trace_out, 'w'), indent=1)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f64', data='synthetic', config=workload_config(world, M),
                    clocks=clocks, gpu_launches=launches, cuda_graph=uses_graph,
                    collective_ms=gather_ms if world > 1 else 0.0,
                    collective='all_gather_into_tensor of the rank-local signal rows (one per diagram, equal shares gathered '
                               'straight into the result) + reshaping, timed alone on the rows of the traced evaluation (smallest '
                               'of five, max over ranks); it is inside every timed step',
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d),
                             d2h_bytes_per_step=int(sum(p.numel() for p in pinned) * 16),
                             api='qsextra_b200.spectroscopy.qspectroscopy_ensemble(new ChromophoreSystem realisations every '
                                 'step, FeynmanDiagram x 3, host_out = pinned tensors) -> signals in pinned host memory (each '
                                 'rank its own members)'),
                    roofline=roofline, cpu_baseline=cpu_baseline,
                    parity_max_rel_err_vs_oracle=None if parity is None else parity['max_rel_err'], parity=parity,
                    wall_s_timed_region=wall, collision_steps=extras_c2, spectroscopy=extras_spec)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--members', type=int, default=8, help='disorder realisations of the C4 aggregate per GPU')
    ap.add_argument('--no-extras', action='store_true', help='skip the C2 / C3 / single-system extras')
    ap.add_argument('--no-spectroscopy', action='store_true', help=argparse.SUPPRESS)     # (older name of --no-extras)
    ap.add_argument('--trace-out', default=None, help='write the per-launch device times of one evaluation to this file')
    args = ap.parse_args()
    args.no_extras = args.no_extras or args.no_spectroscopy
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
