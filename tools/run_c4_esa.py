"""Run the C4 ESA diagram once (for ncu): python tools/run_c4_esa.py [n1 n2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.simplefilter('ignore')
import numpy as np, torch
from tools.spectro_bench import systems
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy
system, kw, delays = systems('c4')
if len(sys.argv) > 2:
    n1, n2 = int(sys.argv[1]), int(sys.argv[2])
    delays = [delays[0][:n1], delays[1][:n2], delays[2]]
spec = FeynmanDiagram(sys.argv[3] if len(sys.argv) > 3 else 'esa', [np.asarray(d).tolist() for d in delays])
torch.cuda.synchronize(); t0 = time.perf_counter()
sig = qspectroscopy(system, spec, shots=0, verbose=False, **kw)
torch.cuda.synchronize(); print('ms', (time.perf_counter() - t0) * 1e3, float(np.abs(sig).sum()))
