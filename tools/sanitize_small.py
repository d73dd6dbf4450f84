"""Tiny run through every kernel family (for compute-sanitizer): RR static + generic, RRC, pass program, op list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.simplefilter('ignore')
import numpy as np
from qsextra_b200 import ChromophoreSystem, ExcitonicSystem
from qsextra_b200.qcomo import qevolve
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy
DT = 0.15


def chain(n, modes=True, d=2):
    J = np.zeros((n, n))
    for i in range(n - 1):
        J[i, i + 1] = J[i + 1, i] = -0.1
    kw = dict(energies=(1.5 + 0.01 * np.arange(n)).tolist(), couplings=J, dipole_moments=[1.] * n)
    if modes:
        return ChromophoreSystem(**kw, frequencies_pseudomode=0.2, levels_pseudomode=d, couplings_ep=0.3)
    return ExcitonicSystem(**kw)


t = [0., 2 * DT, 4 * DT]
out = []
for system, kw in ((chain(2), dict(coll_rates=[0.2])),            # RR static
                   (chain(3), dict(coll_rates=[0.2])),            # RRC (64 threads) + pass program
                   (chain(3, modes=False), dict(coll_rates=0.1)),  # pass program / op list
                   (chain(2, d=3), dict(coll_rates=[0.2]))):      # op list with ancilla + reset
    for lab in ('gsb', 'esa'):
        sig = qspectroscopy(system, FeynmanDiagram(lab, [t, [0., DT], t]), shots=0, verbose=False, dt=DT, **kw)
        out.append(float(np.abs(sig).sum()))
    system.set_state('localized excitation', 0)
    out.append(float(qevolve(system, t, shots=0, dt=DT, verbose=False, **kw).populations.sum()))
print('ok', out)
