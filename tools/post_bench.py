"""Time the device post-processing of one C4-sized signal (128 x 16 x 128, pad_extension 3), call by call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qsextra_b200.spectroscopy import FeynmanDiagram
from qsextra_b200.spectroscopy.postprocessing import postprocessing
t = (np.arange(128) * 1.5).tolist()
spec = FeynmanDiagram('gsb', [t, (np.arange(16) * 10.).tolist(), t])
sig = torch.randn(128, 16, 128, dtype=torch.complex128, device='cuda')
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for idx in range(16):
        w, s = postprocessing(spec, sig, RF_freq=1.5, damping_rate=0.005, pad_extension=3, T2_index=idx)
    torch.cuda.synchronize()
    print('rep %d: %.3f ms per spectrum' % (rep, (time.perf_counter() - t0) / 16 * 1e3))
import cProfile, pstats
cProfile.run('for i in range(16): postprocessing(spec, sig, RF_freq=1.5, damping_rate=0.005, pad_extension=3, T2_index=i)\ntorch.cuda.synchronize()', '/tmp/p.prof')
pstats.Stats('/tmp/p.prof').sort_stats('tottime').print_stats(14)
