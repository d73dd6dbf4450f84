"""One warm-up and one measured evaluation of the full C4 grid (three diagrams); for profiler launch lists."""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np, torch
from tools.spectro_bench import systems
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy
system, kw, delays = systems('c4')
delays = [np.asarray(d).tolist() for d in delays]
specs = [FeynmanDiagram(l, delays) for l in ('gsb', 'se', 'esa')]
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sigs = [qspectroscopy(system, sp, shots=0, verbose=False, **kw) for sp in specs]
    torch.cuda.synchronize()
    print('run %d: %.1f ms, checksum %.9f' % (rep, (time.perf_counter() - t0) * 1e3, sum(float(np.abs(s).sum()) for s in sigs)))
