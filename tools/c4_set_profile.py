"""Host-side profile of qspectroscopy_set on the C4 grid (cProfile, cumulative)."""
import sys, os, time, warnings, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np, torch
from tools.spectro_bench import systems
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy_set
system, kw, delays = systems('c4')
delays = [np.asarray(d).tolist() for d in delays]
specs = [FeynmanDiagram(l, delays) for l in ('gsb', 'se', 'esa')]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = qspectroscopy_set(system, specs, verbose=False, **kw)
    torch.cuda.synchronize(); print('rep %d: %.2f ms' % (rep, (time.perf_counter() - t0) * 1e3))
torch.cuda.synchronize(); t0 = time.perf_counter()
out = qspectroscopy_set(system, specs, verbose=False, device_output=True, **kw)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('device_output: host returned after %.2f ms, GPU done after %.2f ms' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
cProfile.run('qspectroscopy_set(system, specs, verbose=False, device_output=True, **kw)', '/tmp/s.prof')
pstats.Stats('/tmp/s.prof').sort_stats('tottime').print_stats(16)
