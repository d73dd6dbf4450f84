"""Wall-clock timeline of one end-to-end C2 call (host phases, then the wait for kernel + device-to-host copy)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import c2_parameters, DT, N_SETS, N_STEPS
from qsextra_b200.engine.program import SystemBatch, StepProgram
from qsextra_b200.engine.runtime import get_engine
from qsextra_b200.engine.sectors import population_observables
from qsextra_b200.qcomo import qevolve_batch
import qsextra_b200.qcomo.evolution.qevolve as qe

params = c2_parameters(N_SETS)
batch = SystemBatch(params['energies'], params['couplings'], params['dipole_moments'], params['omega'], params['coupl_ep'], (2,))
tg = np.arange(N_STEPS + 1) * DT
pinned = torch.empty((N_SETS, N_STEPS + 1, 2), dtype=torch.float64).pin_memory()
eng = get_engine()
marks = []
def mark(label):
    marks.append((label, time.perf_counter()))

def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        mark('> ' + label); r = f(*a, **k); mark('< ' + label); return r
    setattr(obj, name, g)

wrap(StepProgram, 'bind', 'bind')
wrap(type(eng), 'upload_program', 'upload_program')
wrap(type(eng), 'upload_observables', 'upload_observables')
wrap(type(eng), 'evolve', 'evolve')
def once():
    marks.clear(); mark('start')
    res = qevolve_batch(batch, tg, dt=DT, coll_rates=params['rates'], device_output=True, shard=False)
    mark('api returned')
    pinned.copy_(res, non_blocking=True); mark('copy queued')
    torch.cuda.synchronize(); mark('synced')
for _ in range(5): once()
import collections
acc = collections.OrderedDict()
N = 40
for _ in range(N):
    once()
    t0 = marks[0][1]
    for label, t in marks:
        acc[label] = acc.get(label, 0.0) + (t - t0)
for label, t in acc.items():
    print('%8.3f ms  %s   (mean of %d calls)' % (t / N * 1e3, label, N))
# device-side: kernel alone and copy alone
res = qevolve_batch(batch, tg, dt=DT, coll_rates=params['rates'], device_output=True, shard=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pinned.copy_(res, non_blocking=True); e1.record(); torch.cuda.synchronize()
print('D2H %.3f ms for %.1f MB -> %.1f GB/s' % (e0.elapsed_time(e1), res.numel() * 8 / 1e6, res.numel() * 8 / 1e6 / e0.elapsed_time(e1)))
