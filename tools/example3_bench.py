"""Example 3 of the reference ("More is better"): a dimer with d-level pseudomodes, d = 2, 4, 16 -- time of qevolve."""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np, torch
from qsextra_b200 import ChromophoreSystem
from qsextra_b200.qcomo import qevolve
from qsextra_b200.qcomo.evolution.qevolve import _program_for
from qsextra_b200.engine.runtime import get_engine
eng = get_engine()
dt = 0.05
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for d in (2, 4, 8, 16):
    s = ChromophoreSystem(energies=[1.0, 1.2], couplings=0.1, dipole_moments=[1., 1.], frequencies_pseudomode=[0.5],
                          levels_pseudomode=[d], couplings_ep=[0.3])
    s.set_state('localized excitation', 0)
    t = np.arange(n_steps + 1) * dt
    prog = _program_for(s, dt, dt, 1, 1, [0.4])
    blk = prog.block(1, 1)
    bound = prog.bind(blk, fused=False)
    for rep in range(2):
        torch.cuda.synchronize(); l0 = eng.launches; t0 = time.perf_counter()
        res = qevolve(s, t, shots=0, dt=dt, coll_rates=[0.4], verbose=False)
        torch.cuda.synchronize(); sec = time.perf_counter() - t0
    print('d = %2d: block %d x %d (%d bits), %d operations per step, %d steps in %.3f s (%.1f us per step), trace %.12f' % (
        d, blk.dim_a, blk.dim_b, blk.n_bits, len(bound.ops), n_steps, sec, sec / n_steps * 1e6, res.populations[-1].sum()))
