"""Throughput of the C2 kernel (populations after every step) as a function of the number of parameter sets per GPU:
the BASELINE batch (4096 sets = 512 warps) leaves the GPU latency-bound; larger batches show the kernel's own rate."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import c2_parameters, DT, N_STEPS
from qsextra_b200.engine.program import SystemBatch, StepProgram
from qsextra_b200.engine.runtime import get_engine
from qsextra_b200.engine.sectors import population_observables
eng = get_engine()
peak = eng.fp64_peak()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for n_sets in [int(a) for a in sys.argv[1:]] or [4096, 16384, 65536, 262144]:
    p = c2_parameters(n_sets)
    batch = SystemBatch(p['energies'], p['couplings'], p['dipole_moments'], p['omega'], p['coupl_ep'], (2,))
    prog = StepProgram(batch, DT, 1, 1, p['rates'])
    block = prog.block(1, 1)
    dev = eng.upload_program(prog.bind(block, fused=False), prog)
    obs = eng.upload_observables(*population_observables(block), kind='populations')
    sigma0 = np.zeros((1, block.size), dtype=complex); sigma0[0, (1 << block.n_bits) * block.dim_b + (1 << block.n_bits)] = 1.0
    sig = eng.to_device(sigma0)
    emit = eng.to_device(np.arange(N_STEPS + 1), torch.int32)
    inst_set = torch.arange(n_sets, dtype=torch.int32, device='cuda')
    inst_src = torch.zeros(n_sets, dtype=torch.int32, device='cuda')
    times = []
    for rep in range(6):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, _ = eng.evolve(dev, sig, n_sets, N_STEPS, emit, obs=obs, obs_real=True, inst_set=inst_set, inst_src=inst_src)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times[1:])[len(times[1:]) // 2]
    ex_flops, ex_instr = dev['rr'].executed_flops_per_step()
    rate = n_sets * N_STEPS / (ms * 1e-3)
    print(json.dumps(dict(parameter_sets=n_sets, warps=n_sets * 4 // 32, ms=round(ms, 3), steps_per_s=rate,
                          executed_tflops=ex_flops * rate / 1e12, frac_of_dfma_peak=ex_flops * rate / 1e12 / peak,
                          fp64_pipe_estimate=ex_instr * rate / (64 * eng.sm_count * 1.965e9),
                          trace=float(out[-1, -1].sum()))))
