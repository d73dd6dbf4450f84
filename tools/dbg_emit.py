import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.simplefilter('ignore')
import numpy as np, torch
from qsextra_b200 import ExcitonicSystem
from qsextra_b200.engine.program import StepProgram, SystemBatch
from qsextra_b200.engine.runtime import get_engine
eng = get_engine()
d = ExcitonicSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 0.8])
prog = StepProgram(SystemBatch.from_systems([d]), 0.15, 1, 1, np.array([0.015]))
mu = np.array([1., 0.8])
for ka, kb in ((2, 1), (1, 0), (1, 2), (0, 1)):
    blk = prog.block(ka, kb)
    dev = eng.upload_program(prog.bind(blk), prog)
    x = np.arange(1, blk.size + 1) * (1 + 0.5j)
    sig = eng.to_device(x[None].astype(complex))
    ptr, idx, w = blk.emission_observable(mu)
    want = np.sum(w * x[idx])
    for kind in ('emission', None):
        obs = eng.upload_observables(ptr, idx, w, kind=kind, mu=mu)
        out, _ = eng.evolve(dev, sig, 1, 0, np.array([0]), obs=obs)
        print((ka, kb), kind, 'rr' in dev, out.cpu().numpy().ravel(), 'want', want)
