"""Kernel-only timing sweep of the C2 workload over team size / pseudomode tile width (run under gpurun)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import c2_parameters, DT, N_SETS, N_STEPS
from qsextra_b200.engine.program import StepProgram, SystemBatch
from qsextra_b200.engine.runtime import get_engine
from qsextra_b200.engine.sectors import population_observables

eng = get_engine()
params = c2_parameters(N_SETS)
batch = SystemBatch(params['energies'], params['couplings'], params['dipole_moments'], params['omega'], params['coupl_ep'], (2,))
program = StepProgram(batch, DT, 1, 1, params['rates'])
block = program.block(1, 1)
sigma0 = torch.zeros((1, block.size), dtype=torch.complex128, device=eng.device); sigma0[0, 0] = 1.0
inst_set = torch.arange(N_SETS, dtype=torch.int32, device=eng.device)
inst_src = torch.zeros(N_SETS, dtype=torch.int32, device=eng.device)
obs = eng.upload_observables(*population_observables(block), kind='populations')
configs = [(rb, None, 'rr') for rb in (2, 3, 4)]
ref = None
for team, kb, kind in configs:
    if kind == 'rr':
        bound = program.bind(block, fused=False)
        dev = eng.upload_program(bound, program, reg_bits=team)
        assert 'rr' in dev
    else:
        bound = program.bind(block, fused=(kind == 'passes'), team=team, kb=kb)
        dev = eng.upload_program(bound)
    for emit_every in (True, False):
        emit = eng.to_device(np.arange(N_STEPS + 1) if emit_every else np.array([N_STEPS]), torch.int32)
        run = lambda: eng.evolve(dev, sigma0, N_SETS, N_STEPS, emit, obs=obs, obs_real=True, inst_set=inst_set, inst_src=inst_src)[0]
        out = run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        last = out[:, -1].cpu().numpy()
        if ref is None: ref = last
        print(json.dumps(dict(kind=kind, team=team, kb=kb, emit_every_step=emit_every, ms=round(ms, 3),
                              steps_per_s=N_SETS * N_STEPS / ms * 1e3, n_passes=0 if bound.passes is None or kind == 'ops' else len(bound.passes),
                              err=float(np.abs(last - ref).max()))))
