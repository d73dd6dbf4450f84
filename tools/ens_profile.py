"""cProfile of the end-to-end ensemble call of bench.py (new realisations every call: eager tree, no CUDA graph)."""
import cProfile, os, pstats, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import torch
import bench
from qsextra_b200.spectroscopy import qspectroscopy_ensemble
specs = bench.c4_specs()
kw = dict(dt=bench.DT, coll_rates=[0.2])
def once(seed):
    fresh = bench.c4_members(8, seed=seed)
    sigs = qspectroscopy_ensemble(fresh, specs, shots=0, verbose=False, device_output=True, **kw)
    torch.cuda.synchronize()
    return sigs
for k in range(3):
    once(200 + k)
t0 = time.perf_counter()
for k in range(5):
    once(300 + k)
print('wall per call: %.2f ms' % ((time.perf_counter() - t0) / 5 * 1e3))
pr = cProfile.Profile()
pr.enable()
for k in range(5):
    once(400 + k)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
