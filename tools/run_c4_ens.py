"""The bench workload (C4 grid, MEMBERS disorder realisations, three diagrams) evaluated eagerly a few times: for ncu
captures of its kernels (QSX_MEMBERS, QSX_REPS)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
os.environ.setdefault('QSX_DISABLE_GRAPH', '1')
import numpy as np, torch
import bench
from qsextra_b200.engine import response as R
from qsextra_b200.engine.runtime import get_engine
from qsextra_b200.spectroscopy.simulation.qspectroscopy import ensemble_program
M = int(os.environ.get('QSX_MEMBERS', '8'))
eng = get_engine()
program, mu, sequences, counts = ensemble_program(bench.c4_members(M), bench.c4_specs(), dt=bench.DT, coll_rates=[0.2])
for rep in range(int(os.environ.get('QSX_REPS', '2'))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sigs = R.response_function_set(program, mu, sequences, counts, engine=eng, device_output=True)
    torch.cuda.synchronize()
    print('run %d: %.2f ms, checksum %.9f' % (rep, (time.perf_counter() - t0) * 1e3, sum(float(s.abs().sum()) for s in sigs)))
