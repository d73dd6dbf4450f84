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
ptr, idx, w = population_observables(block)
dev = eng.upload_program(program.bind(block, fused=False), program, reg_bits=4)
for nobs in (2, 1):
    obs = eng.upload_observables(ptr[:nobs + 1], idx[:ptr[nobs]], w[:ptr[nobs]])
    for every in (1, 2, 4, 8, 1000):
        emit = eng.to_device(np.arange(every, N_STEPS + 1, every), torch.int32)
        run = lambda: eng.evolve(dev, sigma0, N_SETS, N_STEPS, emit, obs=obs, obs_real=True, inst_set=inst_set, inst_src=inst_src)[0]
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        print(nobs, 'obs, emit every', every, 'ms', round(e0.elapsed_time(e1) / 5, 3))
