"""Full C4 grid through qspectroscopy_set a few times (wall clock per call; the third call on replays the CUDA graph)."""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np, torch
from tools.spectro_bench import systems
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy_set
system, kw, delays = systems('c4')
delays = [np.asarray(d).tolist() for d in delays]
specs = [FeynmanDiagram(l, delays) for l in ('gsb', 'se', 'esa')]
reps = int(os.environ.get('QSX_REPS', '6'))
for rep in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sigs = qspectroscopy_set(system, specs, verbose=False, **kw)
    torch.cuda.synchronize()
    print('run %d: %.2f ms, checksum %.9f' % (rep, (time.perf_counter() - t0) * 1e3, sum(float(np.abs(s).sum()) for s in sigs)))
