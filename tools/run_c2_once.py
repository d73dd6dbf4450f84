"""Run the C2 kernel a few times in one configuration (for ncu): python tools/run_c2_once.py rr 4 | passes 8 1
(QSX_C2_SETS=65536 runs a batch other than BASELINE's 4096 parameter sets)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import c2_parameters, DT, N_SETS, N_STEPS
N_SETS = int(os.environ.get('QSX_C2_SETS', N_SETS))
from qsextra_b200.engine.program import StepProgram, SystemBatch
from qsextra_b200.engine.runtime import get_engine
from qsextra_b200.engine.sectors import population_observables

kind = sys.argv[1]
eng = get_engine()
params = c2_parameters(N_SETS)
batch = SystemBatch(params['energies'], params['couplings'], params['dipole_moments'], params['omega'], params['coupl_ep'], (2,))
program = StepProgram(batch, DT, 1, 1, params['rates'])
block = program.block(1, 1)
sigma0 = torch.zeros((1, block.size), dtype=torch.complex128, device=eng.device); sigma0[0, 0] = 1.0
inst_set = torch.arange(N_SETS, dtype=torch.int32, device=eng.device)
inst_src = torch.zeros(N_SETS, dtype=torch.int32, device=eng.device)
obs = eng.upload_observables(*population_observables(block), kind='populations')
if kind == 'rr':
    dev = eng.upload_program(program.bind(block, fused=False), program, reg_bits=int(sys.argv[2]))
else:
    dev = eng.upload_program(program.bind(block, team=int(sys.argv[2]), kb=int(sys.argv[3])))
emit = eng.to_device(np.arange(N_STEPS + 1), torch.int32)
for _ in range(4):
    out = eng.evolve(dev, sigma0, N_SETS, N_STEPS, emit, obs=obs, obs_real=True, inst_set=inst_set, inst_src=inst_src)[0]
torch.cuda.synchronize()
print(float(out[:, -1].sum()))
