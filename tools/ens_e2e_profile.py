"""Host-side timeline of the bench's end-to-end step (new realisations -> pinned host signals): when every propagation
is enqueued relative to the start of the call, when the tree returns, when the device is done, when the copies are."""
import os, sys, time, cProfile, pstats, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np, torch
import bench
from qsextra_b200.engine.runtime import get_engine
from qsextra_b200.spectroscopy import qspectroscopy_ensemble
M = int(os.environ.get('QSX_MEMBERS', '8'))
eng = get_engine()
specs = bench.c4_specs()
kw = dict(dt=bench.DT, coll_rates=[0.2])
pinned = [torch.empty((M, bench.N_T1, bench.N_T2, bench.N_T3), dtype=torch.complex128).pin_memory() for _ in bench.DIAGRAMS]
marks = []
t_start = [0.0]
for name in ('_evolve_rrc', '_evolve_rr', 'apply_dipole', 'readout_gemm'):
    original = getattr(eng, name)
    def wrapped(*a, _o=original, _n=name, **k):
        t = time.perf_counter() - t_start[0]
        out = _o(*a, **k)
        marks.append((_n, round(t * 1e3, 2), round((time.perf_counter() - t_start[0]) * 1e3, 2)))
        return out
    setattr(eng, name, wrapped)


def once(seed, log=False):
    marks.clear()
    torch.cuda.synchronize()
    t_start[0] = t0 = time.perf_counter()
    fresh = bench.c4_members(M, seed=seed)
    t1 = time.perf_counter()
    qspectroscopy_ensemble(fresh, specs, shots=0, verbose=False, device_output=True, host_out=pinned, **kw)
    t2 = time.perf_counter()
    if log:
        print('members %.2f | api returns with the signals in pinned host memory %.2f ms' % tuple((t - t0) * 1e3 for t in (t1, t2)))
        for m in marks:
            print('   ', m)


for s in (100, 101):
    once(s)
for s in (102, 103, 104):
    once(s, log=True)
cProfile.run('once(105)', '/tmp/ens.prof')
pstats.Stats('/tmp/ens.prof').sort_stats('tottime').print_stats(30)
pstats.Stats('/tmp/ens.prof').sort_stats('cumtime').print_stats('qsextra_b200|bench', 45)
