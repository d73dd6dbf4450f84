import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import c2_parameters, DT, N_SETS, N_STEPS
from qsextra_b200.engine.program import SystemBatch
from qsextra_b200.qcomo import qevolve_batch
params = c2_parameters(N_SETS)
batch = SystemBatch(params['energies'], params['couplings'], params['dipole_moments'], params['omega'], params['coupl_ep'], (2,))
tg = np.arange(N_STEPS + 1) * DT
pinned = torch.empty((N_SETS, N_STEPS + 1, 2), dtype=torch.float64).pin_memory()
def once():
    res = qevolve_batch(batch, tg, dt=DT, coll_rates=params['rates'], device_output=True, shard=False)
    pinned.copy_(res, non_blocking=True); torch.cuda.synchronize()
once(); once()
t0 = time.perf_counter()
for _ in range(5): once()
print('e2e ms', (time.perf_counter() - t0) / 5 * 1e3)
cProfile.run('once()', '/tmp/e2e.prof')
pstats.Stats('/tmp/e2e.prof').sort_stats('cumtime').print_stats(45)
