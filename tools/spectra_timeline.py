"""C4: simulation with the signal kept on the device, then every 2-D spectrum; times of both stages."""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np, torch
from tools.spectro_bench import systems
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy
from qsextra_b200.spectroscopy.postprocessing import postprocessing
system, kw, delays = systems('c4')
delays = [np.asarray(d).tolist() for d in delays]
specs = [FeynmanDiagram(l, delays) for l in ('gsb', 'se', 'esa')]
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sigs = [qspectroscopy(system, sp, shots=0, verbose=False, device_output=True, **kw) for sp in specs]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    total = torch.zeros((), dtype=torch.float64, device='cuda')
    for sp, sig in zip(specs, sigs):
        for idx in range(len(delays[1])):
            _, spectrum = postprocessing(sp, sig, RF_freq=1.5, damping_rate=0.005, pad_extension=3, T2_index=idx)
            total += spectrum.abs().sum()
    total = float(total)
    t2 = time.perf_counter()
    st = torch.cuda.memory_stats()
    print('rep %d: simulation %.1f ms, 48 spectra %.1f ms' % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3),
          'cudaMalloc calls', st['num_device_alloc'], 'cudaFree calls', st['num_device_free'], 'reserved MB', st['reserved_bytes.all.current'] >> 20)
import cProfile, pstats
def post():
    for idx in range(16):
        postprocessing(specs[0], sigs[0], RF_freq=1.5, damping_rate=0.005, pad_extension=3, T2_index=idx)
    torch.cuda.synchronize()
cProfile.run('post()', '/tmp/q.prof'); pstats.Stats('/tmp/q.prof').sort_stats('tottime').print_stats(8)
print('--- interleaved: simulate one diagram, transform its 16 slices, next diagram')
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    total = torch.zeros((), dtype=torch.float64, device='cuda')
    marks = []
    for sp in specs:
        ta = time.perf_counter()
        sig = qspectroscopy(system, sp, shots=0, verbose=False, device_output=True, **kw)
        tb = time.perf_counter()
        for idx in range(len(delays[1])):
            _, spectrum = postprocessing(sp, sig, RF_freq=1.5, damping_rate=0.005, pad_extension=3, T2_index=idx)
            total += spectrum.abs().sum()
        marks.append(((tb - ta) * 1e3, (time.perf_counter() - tb) * 1e3))
    total = float(total)
    print('rep %d: %.1f ms' % (rep, (time.perf_counter() - t0) * 1e3), ['sim %.1f post %.1f' % m for m in marks])
