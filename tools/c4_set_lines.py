"""Line-level wall-clock profile of the response tree (host side) for qspectroscopy_set on the C4 grid."""
import sys, os, time, warnings, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np, torch
from tools.spectro_bench import systems
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy_set
from qsextra_b200.engine import response as RESP, runtime as RT
system, kw, delays = systems('c4')
delays = [np.asarray(d).tolist() for d in delays]
specs = [FeynmanDiagram(l, delays) for l in ('gsb', 'se', 'esa')]
for _ in range(3):
    qspectroscopy_set(system, specs, verbose=False, device_output=True, **kw)
torch.cuda.synchronize()
targets = {RESP.response_function.__code__, RESP.response_function_set.__code__, RESP._propagate_readout.__code__,
           RT.Engine.upload_program.__code__, RT.Engine._evolve_rrc.__code__, RT.Engine.evolve.__code__}
acc = collections.defaultdict(float); last = [None, 0.0]
def tracer(frame, event, arg):
    if frame.f_code not in targets: return None
    def local(frame, event, arg):
        now = time.perf_counter()
        if last[0] is not None: acc[last[0]] += now - last[1]
        last[0] = (frame.f_code.co_name, frame.f_lineno); last[1] = time.perf_counter()
        return local
    return local
sys.settrace(tracer)
qspectroscopy_set(system, specs, verbose=False, device_output=True, **kw)
sys.settrace(None)
torch.cuda.synchronize()
for (fn, ln), t in sorted(acc.items(), key=lambda kv: -kv[1])[:22]:
    print('%-22s line %4d  %.3f ms' % (fn, ln, t * 1e3))
