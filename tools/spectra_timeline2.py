import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np, torch
from tools.spectro_bench import systems
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy
system, kw, delays = systems('c4')
delays = [np.asarray(d).tolist() for d in delays]
specs = [FeynmanDiagram(l, delays) for l in ('gsb', 'se', 'esa')]
def run(device_output, sync_each):
    torch.cuda.synchronize(); t0 = time.perf_counter(); per = []
    for sp in specs:
        t1 = time.perf_counter()
        s = qspectroscopy(system, sp, shots=0, verbose=False, device_output=device_output, **kw)
        if sync_each: torch.cuda.synchronize()
        per.append((time.perf_counter() - t1) * 1e3)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, per
for mode in [(False, False), (True, True), (True, False), (False, True)]:
    for rep in range(3):
        tot, per = run(*mode)
        print('device_output=%s sync_each=%s rep %d: total %.1f ms' % (mode[0], mode[1], rep, tot), ['%.1f' % p for p in per])
