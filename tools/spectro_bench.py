"""Time the response-function tree on the spectroscopy configs of BASELINE.json (run under gpurun).

    python tools/spectro_bench.py c3 | c3modes | c4small | c4
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings
import numpy as np, torch
warnings.simplefilter('ignore')
from qsextra_b200 import ChromophoreSystem, ExcitonicSystem
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy
from qsextra_b200.engine.runtime import get_engine

HBAR = 0.658212
DT = 0.1 / HBAR
GAMMA = 59.08e-3


def systems(which):
    if which == 'c3':
        s = ExcitonicSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 1.])
        return s, dict(dt=DT, coll_rates=GAMMA / 4), [30 * DT * np.arange(64), [0.], 30 * DT * np.arange(64)]
    if which == 'c3modes':
        s = ChromophoreSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 1.], frequencies_pseudomode=0.02,
                              levels_pseudomode=2, couplings_ep=float(np.sqrt(GAMMA * 0.1 / 2)))
        return s, dict(dt=DT, coll_rates=[0.2]), [30 * DT * np.arange(64), [0.], 30 * DT * np.arange(64)]
    n1, n2 = (128, 16) if which == 'c4' else (16, 4)
    J = np.zeros((4, 4))
    for i in range(3):
        J[i, i + 1] = J[i + 1, i] = -0.01
    s = ChromophoreSystem(energies=[1.55, 1.50, 1.46, 1.52], couplings=J, dipole_moments=[1., 1., 1., 1.],
                          frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=float(np.sqrt(GAMMA * 0.1 / 2)))
    return s, dict(dt=DT, coll_rates=[0.2]), [15 * DT * np.arange(n1), 100 * DT * np.arange(n2), 15 * DT * np.arange(n1)]


def main():
    which = sys.argv[1]
    eng = get_engine()
    system, kw, delays = systems(which)
    out = {}
    total_ms = 0.0
    for label in ('gsb', 'se', 'esa'):
        spec = FeynmanDiagram(label, [np.asarray(d).tolist() for d in delays])
        if which != 'c4':
            qspectroscopy(system, spec, shots=0, verbose=False, **kw)       # warm-up (skipped for the big config)
        torch.cuda.synchronize()
        l0 = eng.launches
        t0 = time.perf_counter()
        sig = qspectroscopy(system, spec, shots=0, verbose=False, **kw)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        total_ms += ms
        out[label] = dict(ms=round(ms, 2), launches=eng.launches - l0, shape=list(sig.shape), checksum=[float(sig.real.sum()), float(sig.imag.sum())],
                          s000=[float(sig[0, 0, 0].real), float(sig[0, 0, 0].imag)])
    points = 3 * int(np.prod([len(d) for d in delays]))
    out['grid_points'] = points
    out['total_ms'] = round(total_ms, 2)
    out['grid_points_per_s'] = points / (total_ms * 1e-3)
    print(json.dumps(dict(config=which, **out)))


if __name__ == '__main__':
    main()
