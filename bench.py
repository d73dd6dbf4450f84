#!/usr/bin/env python
"""bench.py -- headline measurement of the hot path on B200 (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1], "C2"): non-Markovian dimers, one 2-level pseudomode per site, 1000 collision
steps, a batch of 4096 parameter sets PER GPU (weak scaling), populations written after every step
([4096, 1001, 2] float64).  One bench "step" = one pass of the hot path over that batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

metric = collision steps per second (whole job).  `value` is timed with CUDA events around the K launches with
all inputs resident in HBM (L2 flushed between launches, outside the event pairs); `e2e` goes through the public
API `qevolve_batch` from host arrays to a pinned host result.  `--impl reference` times the CPU oracle port
(oracle/csrc/qsx_oracle.c, all host cores) on a bounded sample of the same workload.
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
N_SETS = 4096
N_STEPS = 1000
METRIC = 'collision_steps_per_sec'
UNIT = 'steps/s'


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


def algorithmic_flops_per_step(n_sites=2, n_modes=1, stored_elements=64):
    """F_step = E (6 n_Z + 12 n_R + 2.5 n_AD) for the d=2 pseudomode step, E = stored elements of
    the block (1-exciton sector of the dimer: 8 x 8)."""
    n_z = n_sites + n_sites * n_modes
    n_r = 2 * (n_sites * (n_sites - 1) // 2) + 2 * n_sites * n_modes
    n_ad = n_sites * n_modes
    return stored_elements * (6 * n_z + 12 * n_r + 2.5 * n_ad)


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port)
# ------------------------------------------------------------------------------------------------
def cpu_systems(params, lo, hi):
    import warnings
    from qsextra_b200 import ChromophoreSystem
    out, rates = [], []
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for b in range(lo, hi):
            s = ChromophoreSystem(energies=params['energies'][b].tolist(), couplings=float(params['couplings'][b, 0, 1]),
                                  dipole_moments=[1., 1.], frequencies_pseudomode=float(params['omega'][b, 0]),
                                  levels_pseudomode=2, couplings_ep=float(params['coupl_ep'][b, 0]))
            s.set_state('localized excitation', 0)
            out.append(s)
            rates.append([float(params['rates'][b, 0])])
    return out, np.array(rates)


def cpu_sample(params, n_sample, n_steps=N_STEPS, keep=None):
    """Time the C oracle on the first n_sample parameter sets; -> (steps/s, threads, seconds).  ``keep`` (a list)
    receives the populations it computed, so that the GPU result of the same sets can be compared with them."""
    from oracle import fast
    systems, rates = cpu_systems(params, 0, n_sample)
    time_grid = np.arange(n_steps + 1) * DT
    fast.populations_batch(systems[:2], time_grid[:3], dt=DT, coll_rates=rates[:2])      # load + warm
    t0 = time.perf_counter()
    pops, threads = fast.populations_batch(systems, time_grid, dt=DT, coll_rates=rates, n_threads=os.cpu_count() or 1)
    sec = time.perf_counter() - t0
    if keep is not None:
        keep.append(pops)
    return n_sample * n_steps / sec, threads, sec


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    params = c2_parameters(N_SETS)
    n_sample = 128
    cpu_sample(params, 8, 50)
    for _ in range(args.warmup):
        cpu_sample(params, 16, 100)
    t0 = time.perf_counter()
    threads = 1
    for _ in range(args.steps):
        _, threads, _ = cpu_sample(params, n_sample)
    sec = time.perf_counter() - t0
    value = args.steps * n_sample * N_STEPS / sec
    sample = '{} of {} parameter sets x {} steps per bench step (C oracle port, full-space literal gates, OpenMP)'.format(
        n_sample, N_SETS, N_STEPS)
    line = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * sec / args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f64', data='synthetic',
                config=workload_config(args.gpus),
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind='port', sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def workload_config(n_gpus):
    return dict(workload='C2: qcomo qevolve, non-Markovian dimer, 1 two-level pseudomode/site, {} collision steps, '
                         '{} parameter sets per GPU, populations after every step'.format(N_STEPS, N_SETS),
                parameter_sets_per_gpu=N_SETS, collision_steps=N_STEPS, global_parameter_sets=N_SETS * n_gpus,
                sharding='parameter sets over ranks, no data-path collective',
                cache='L2 flushed (256 MiB write) between timed launches, outside the event pairs')


# ------------------------------------------------------------------------------------------------
# GPU arm
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


def spectroscopy_extras(world, rank, barrier):
    """BASELINE metric part 1, '2DES grid points/sec': full C3 (dimer, 64 x 64, t2 = 0, the three rephasing and the
    three non-rephasing pathways) and full C4 (4-site chain with
    pseudomodes, 128 x 128 x 16) grids, three rephasing diagrams each, through the public qspectroscopy API (delay-grid
    chains sharded over the ranks, one all-gather).  Wall-clock with synchronisation on both sides, max over ranks by the
    barrier.  Reported next to the headline line; not part of the timed K steps."""
    import warnings
    import torch
    from qsextra_b200 import ChromophoreSystem, ExcitonicSystem
    from qsextra_b200.spectroscopy import FeynmanDiagram, Spectroscopy, qspectroscopy, qspectroscopy_set
    gamma = 59.08e-3
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        dimer = ExcitonicSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 1.])
        J = np.zeros((4, 4))
        for i in range(3):
            J[i, i + 1] = J[i + 1, i] = -0.01
        chain = ChromophoreSystem(energies=[1.55, 1.50, 1.46, 1.52], couplings=J, dipole_moments=[1., 1., 1., 1.],
                                  frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=float(np.sqrt(gamma * 0.1 / 2)))
        configs = {'C3_dimer_64x1x64_6pathways': (dimer, dict(dt=DT, coll_rates=gamma / 4),
                                        [(30 * DT * np.arange(64)).tolist(), [0.], (30 * DT * np.arange(64)).tolist()]),
                   'C4_chain4_pseudomodes_128x16x128': (chain, dict(dt=DT, coll_rates=[0.2]),
                                                        [(15 * DT * np.arange(128)).tolist(), (100 * DT * np.arange(16)).tolist(),
                                                         (15 * DT * np.arange(128)).tolist()])}
        for name, (system, kw, delays) in configs.items():
            try:
                specs = [FeynmanDiagram(lab, delays) for lab in ('gsb', 'se', 'esa')]
                if name.startswith('C3'):                       # config 2 asks for the non-rephasing pathways as well
                    for sides, directions in (('kkkk', 'ioio'), ('kbbk', 'iioo'), ('kbkk', 'iiio')):
                        extra = Spectroscopy(delays)
                        extra.side_seq, extra.direction_seq = sides, directions
                        specs.append(extra)
                for sp in specs:
                    qspectroscopy(system, sp, shots=0, verbose=False, **kw)              # warm-up (module loads, caches)
                def timed(run, repeats=3):
                    """Median wall-clock of ``repeats`` evaluations (each between barriers) and the last result."""
                    samples, result = [], None
                    for _ in range(repeats):
                        barrier()
                        t_start = time.perf_counter()
                        result = run()
                        barrier()
                        samples.append(time.perf_counter() - t_start)
                    return sorted(samples)[len(samples) // 2], samples, result

                sec, samples, sigs = timed(lambda: [qspectroscopy(system, sp, shots=0, verbose=False, **kw) for sp in specs])
                points = len(specs) * int(np.prod([len(d) for d in delays]))
                out[name] = dict(grid_points=points, seconds=sec, grid_points_per_s=points / sec, warmup_runs=1,
                                 samples_s=samples, checksum=float(sum(np.abs(s_).sum() for s_ in sigs)))
                # the same diagrams through qspectroscopy_set: common delay levels once, every read-out propagation
                # enqueued up front (nothing of the runs above is reused: the sharing lives inside the call)
                sec, samples, sigs = timed(lambda: qspectroscopy_set(system, specs, verbose=False, **kw))
                out[name]['as_one_set'] = dict(seconds=sec, grid_points_per_s=points / sec, api='qspectroscopy_set',
                                               samples_s=samples,
                                               checksum=float(sum(np.abs(s_).sum() for s_ in sigs)))
                if name.startswith('C4'):
                    # the full 2-D spectra (every t2 slice, pad_extension 3) with the signal kept on the device from
                    # the simulation to the transformed spectrum; only the spectra's checksum comes back to the host
                    from qsextra_b200.spectroscopy.postprocessing import postprocessing
                    import torch as _torch
                    warm = qspectroscopy(system, specs[0], shots=0, verbose=False, device_output=True, **kw)
                    float(postprocessing(specs[0], warm, pad_extension=3)[1].abs().sum())   # cuFFT plans, lazy module loads
                    barrier()
                    t0 = time.perf_counter()
                    total = _torch.zeros((), dtype=_torch.float64, device='cuda')
                    for sp in specs:
                        sig = qspectroscopy(system, sp, shots=0, verbose=False, device_output=True, **kw)
                        for idx in range(len(delays[1])):
                            _, spectrum = postprocessing(sp, sig, RF_freq=1.5, damping_rate=0.005, pad_extension=3, T2_index=idx)
                            total += spectrum.abs().sum()
                    total = float(total)                                     # one device-to-host read at the end
                    barrier()
                    sec = time.perf_counter() - t0
                    out[name]['with_spectra'] = dict(seconds=sec, spectra=3 * len(delays[1]), shape=list(spectrum.shape),
                                                     grid_points_per_s=points / sec, checksum=float(total))
                    # the whole pipeline with the batched entry points: one set call, then all t2 slices per diagram
                    from qsextra_b200.spectroscopy.postprocessing import postprocessing_all_T2

                    def pipeline():
                        signals = qspectroscopy_set(system, specs, verbose=False, device_output=True, **kw)
                        acc = _torch.zeros((), dtype=_torch.float64, device='cuda')
                        for sp, sig in zip(specs, signals):
                            acc += postprocessing_all_T2(sp, sig, RF_freq=1.5, damping_rate=0.005, pad_extension=3)[1].abs().sum()
                        return float(acc)

                    pipeline()
                    sec, samples, total = timed(pipeline)
                    out[name]['set_with_spectra'] = dict(seconds=sec, samples_s=samples, spectra=3 * len(delays[1]),
                                                         grid_points_per_s=points / sec, checksum=total,
                                                         api='qspectroscopy_set + postprocessing_all_T2')
            except Exception as exc:                                              # keep the headline line alive
                out[name] = dict(error=repr(exc))
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ['NCCL_DEBUG'] = 'WARN'          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from qsextra_b200.engine.program import StepProgram, SystemBatch
    from qsextra_b200.engine.runtime import get_engine
    from qsextra_b200.engine.sectors import population_observables
    from qsextra_b200.qcomo import qevolve_batch

    eng = get_engine()
    params = c2_parameters(N_SETS * world)                      # global batch; this rank owns a contiguous slice
    lo, hi = rank * N_SETS, (rank + 1) * N_SETS
    batch = SystemBatch(params['energies'][lo:hi], params['couplings'][lo:hi], params['dipole_moments'][lo:hi],
                        params['omega'][lo:hi], params['coupl_ep'][lo:hi], (2,))
    program = StepProgram(batch, DT, 1, 1, params['rates'][lo:hi])
    block = program.block(1, 1)
    bound = program.bind(block)
    dev = eng.upload_program(bound, program)
    sigma0 = torch.zeros((1, block.size), dtype=torch.complex128, device=eng.device)
    site0 = (0 << block.n_bits) * block.dim_b + (0 << block.n_bits)    # |cfg 0 (site 0 excited), modes 0>
    sigma0[0, site0] = 1.0
    inst_set = torch.arange(N_SETS, dtype=torch.int32, device=eng.device)
    inst_src = torch.zeros(N_SETS, dtype=torch.int32, device=eng.device)
    emit = eng.to_device(np.arange(N_STEPS + 1), torch.int32)
    obs = eng.upload_observables(*population_observables(block), kind='populations')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)

    def hot_path():
        return eng.evolve(dev, sigma0, N_SETS, N_STEPS, emit, obs=obs, obs_real=True, inst_set=inst_set,
                          inst_src=inst_src)[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    sampler = ClockSampler(local_rank)
    sampler.wait_for_first_sample()
    warm_until = time.time() + 0.5                  # at least W steps and half a second under load before timing
    n_warm = 0
    while n_warm < max(args.warmup, 3) or time.time() < warm_until:
        flush.fill_(1)
        out = hot_path()
        n_warm += 1
        if n_warm % 16 == 0:
            torch.cuda.synchronize()
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
        out = hot_path()
        stops[i].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launches - launches0
    clocks = sampler.stop(t_region0, time.time())
    kernel_ms = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    t_dev = torch.tensor([sum(kernel_ms)], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())
    steps_done = args.steps * N_SETS * N_STEPS * world
    value = steps_done / (total_ms * 1e-3)

    timed_output = out.cpu().numpy() if rank == 0 else None            # compared with the cpu_baseline run below

    # ---- end to end through the public API: host arrays in, pinned host populations out ----------------------
    pinned = torch.empty((N_SETS, N_STEPS + 1, 2), dtype=torch.float64).pin_memory()
    host_batch = SystemBatch(np.array(batch.energies), np.array(batch.couplings), np.array(batch.dipole_moments),
                             np.array(batch.omega), np.array(batch.coupl_ep), (2,))
    time_grid = np.arange(N_STEPS + 1) * DT
    world_was = world

    def e2e_once():
        # per-rank call on this rank's own parameter sets (no gather: weak scaling, independent units)
        res = qevolve_batch(host_batch, time_grid, dt=DT, coll_rates=params['rates'][lo:hi], device_output=True,
                            shard=False)
        pinned.copy_(res, non_blocking=True)
        torch.cuda.synchronize()

    e2e_once()
    h2d0 = eng.h2d_bytes
    barrier()
    t0 = time.perf_counter()
    e2e_iters = max(1, min(args.steps, 5))
    for _ in range(e2e_iters):
        e2e_once()
    barrier()
    e2e_sec = time.perf_counter() - t0
    h2d = (eng.h2d_bytes - h2d0) // e2e_iters
    t_e2e = torch.tensor([e2e_sec], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = e2e_iters * N_SETS * N_STEPS * world_was / float(t_e2e.item())

    spectro = spectroscopy_extras(world, rank, barrier) if not args.no_spectroscopy else None

    if rank == 0:
        fp64_peak = eng.fp64_peak(5)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
        except OSError:
            pass
        f_step = algorithmic_flops_per_step(2, 1, block.size)
        executed = None
        if 'rr' in dev:
            ex_flops, ex_instr = dev['rr'].executed_flops_per_step()
            ex_tf = ex_flops * N_SETS * N_STEPS / (float(np.mean(kernel_ms)) * 1e-3) / 1e12
            # FP64 pipe: 64 lanes per clock and SM; every DFMA / DMUL / DADD occupies one lane-slot
            sm_clock = (clocks.get('sm_mhz') or 1965.0) * 1e6
            pipe = ex_instr * N_SETS * N_STEPS / (float(np.mean(kernel_ms)) * 1e-3) / (64 * eng.sm_count * sm_clock)
            executed = dict(flops_per_collision_step=ex_flops, fp64_instructions_per_collision_step=ex_instr,
                            achieved=ex_tf, unit='TFLOP/s', frac_of_dfma_peak=ex_tf / fp64_peak if fp64_peak else None,
                            fp64_pipe_utilisation_estimate=pipe,
                            note='flops the register-resident kernel executes, counted from its operation list '
                                 '(read-outs excluded); the SURVEY formula above counts the un-fused gate list')
        # The BASELINE batch (4096 sets = 512 warps) leaves the GPU latency-bound; the same kernel on a 16 x larger batch
        # of the same distribution shows its own rate (reported next to, never instead of, the BASELINE configuration).
        saturated = None
        if 'rr' in dev:
            big_n = 16 * N_SETS
            bp = c2_parameters(big_n)
            bb = SystemBatch(bp['energies'], bp['couplings'], bp['dipole_moments'], bp['omega'], bp['coupl_ep'], (2,))
            bprog = StepProgram(bb, DT, 1, 1, bp['rates'])
            bdev = eng.upload_program(bprog.bind(block, fused=False), bprog)
            b_set = torch.arange(big_n, dtype=torch.int32, device=eng.device)
            b_src = torch.zeros(big_n, dtype=torch.int32, device=eng.device)
            times = []
            for rep in range(4):
                flush.fill_(rep)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.evolve(bdev, sigma0, big_n, N_STEPS, emit, obs=obs, obs_real=True, inst_set=b_set, inst_src=b_src)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            big_ms = sorted(times[1:])[1]
            big_rate = big_n * N_STEPS / (big_ms * 1e-3)
            saturated = dict(parameter_sets=big_n, kernel_ms_per_launch=big_ms, steps_per_s=big_rate,
                             achieved=f_step * big_rate / 1e12, frac=f_step * big_rate / 1e12 / fp64_peak if fp64_peak else None,
                             executed_achieved=ex_flops * big_rate / 1e12,
                             executed_frac_of_dfma_peak=ex_flops * big_rate / 1e12 / fp64_peak if fp64_peak else None,
                             fp64_pipe_utilisation_estimate=ex_instr * big_rate / (64 * eng.sm_count * sm_clock),
                             note='same kernel, 16 x the BASELINE batch per GPU: enough warps to hide the latencies')
            del bdev, b_set, b_src
        per_launch_flops = f_step * N_SETS * N_STEPS
        per_launch_ms = float(np.mean(kernel_ms))
        achieved = per_launch_flops / (per_launch_ms * 1e-3) / 1e12
        out_bytes = N_SETS * (N_STEPS + 1) * 2 * 8
        in_bytes = bound.params.nbytes + block.size * 16
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        cpu_baseline = parity = None
        if world == 1:                                  # reported on rank 0 at N = 1 only
            kept = []
            cpu_value, cpu_threads, cpu_sec = cpu_sample(params, N_SETS, keep=kept)
            parity = float(np.abs(timed_output - kept[0]).max())    # the timed kernel's output vs the CPU baseline's
            cpu_baseline = dict(value=cpu_value, unit=UNIT, cores=cpu_threads, kind='port',
                                sample='all %d parameter sets x %d steps (the whole workload of one bench step), C oracle '
                                       'port (full-space literal gates, OpenMP), %.1f s' % (N_SETS, N_STEPS, cpu_sec))
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f64', data='synthetic', config=workload_config(world),
                    clocks=clocks, gpu_launches=launches,
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(pinned.numel() * 8),
                             api='qsextra_b200.qcomo.qevolve_batch(SystemBatch of host arrays) -> pinned host tensor'),
                    roofline=dict(bound='fp64', kernel='rr::evolve_rr_static_kernel<0, 1> (register-resident, compile-time program)', achieved=achieved, peak=fp64_peak,
                                  unit='TFLOP/s', frac=achieved / fp64_peak if fp64_peak else None,
                                  traffic=18869000, traffic_source='dram__bytes_read.sum + dram__bytes_write.sum of one launch '
                                  'in profiles/r1_c2_static_rr_kernel.md (ncu --set full); most of the 65.6 MB of read-outs is '
                                  'still in L2 when the kernel ends',
                                  peak_source='qsx_fp64_peak DFMA micro-benchmark measured in this run '
                                              '(MEASURED_PEAKS.json has no FP64 figure)',
                                  flops_per_collision_step=f_step, flops_per_launch=per_launch_flops,
                                  kernel_ms_per_launch=per_launch_ms, executed=executed, saturated=saturated,
                                  hbm=dict(achieved=(in_bytes + out_bytes) / (per_launch_ms * 1e-3) / 1e9, peak=hbm_peak,
                                           unit='GB/s', algorithmic_bytes_per_launch=in_bytes + out_bytes,
                                           peak_source='MEASURED_PEAKS.json' if peaks else 'fallback')),
                    cpu_baseline=cpu_baseline,
                    parity_max_abs_err_vs_oracle=parity, wall_s_timed_region=wall, spectroscopy=spectro)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-spectroscopy', action='store_true', help='skip the 2DES extras (C3 / C4 grid points per second)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
