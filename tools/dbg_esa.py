import sys, os
sys.path.insert(0, '/root/repo')
import warnings; warnings.simplefilter('ignore')
import numpy as np
from qsextra_b200 import ExcitonicSystem
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy
from oracle import reference_path as R
s = ExcitonicSystem(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 0.8])
DT = 0.1/0.658212
kw = dict(dt=DT, coll_rates=59.08e-3/4)
spec = FeynmanDiagram('esa', [[0., 3*DT], [0.], [0., 3*DT]])
print(os.environ.get('QSX_DISABLE_RR'), os.environ.get('QSX_RR_NO_STATIC'))
print(qspectroscopy(s, spec, verbose=False, **kw).ravel())
print(R.qspectroscopy_operator(s, spec, **kw).ravel())
