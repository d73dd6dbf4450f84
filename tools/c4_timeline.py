"""Per-call timeline of the C4 response tree (each engine call synchronised and timed; host gaps shown)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings
import numpy as np, torch
warnings.simplefilter('ignore')
from tools.spectro_bench import systems
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy
from qsextra_b200.engine.runtime import get_engine
from qsextra_b200.engine.program import StepProgram

eng = get_engine()
rows = []
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); rows.append((name, (time.perf_counter() - t0) * 1e3, info(name, a, k)))
        return r
    setattr(obj, name, g)
def info(name, a, k):
    if name == 'evolve':
        prog = a[1]
        return 'block %s inst %d steps %d emits %d %s' % ((prog['block'].block.ka, prog['block'].block.kb), a[3], a[4], len(a[5]),
                                                           'rrc' if prog.get('rrc') is not None else 'rr' if 'rr' in prog else 'pass')
    return ''
for name in ('evolve', 'apply_dipole', 'upload_program', 'upload_observables', 'readout_gemm'):
    wrap(type(eng), name)
wrap(StepProgram, 'bind')
system, kw, delays = systems(sys.argv[1] if len(sys.argv) > 1 else 'c4')
for label in ('gsb', 'se', 'esa'):
    spec = FeynmanDiagram(label, [np.asarray(d).tolist() for d in delays])
    for rep in range(2):
        rows.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        qspectroscopy(system, spec, shots=0, verbose=False, **kw)
        torch.cuda.synchronize(); total = (time.perf_counter() - t0) * 1e3
    print('== %s total %.2f ms (second run), in engine calls %.2f ms' % (label, total, sum(r[1] for r in rows)))
    for name, ms, what in rows:
        print('   %-18s %8.3f ms  %s' % (name, ms, what))
