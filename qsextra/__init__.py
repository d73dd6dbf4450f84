"""Drop-in alias: the reference's import paths (``qsextra``, ``qsextra.qcomo``, ``qsextra.spectroscopy``,
``qsextra.spectroscopy.postprocessing``, ``qsextra.tools`` ...) resolve to ``qsextra_b200``.

    from qsextra import ExcitonicSystem, ChromophoreSystem            # reference qsextra/__init__.py:1
    from qsextra.qcomo import qevolve, clevolve                       # qcomo/__init__.py:1
    from qsextra.spectroscopy import FeynmanDiagram, qspectroscopy    # spectroscopy/__init__.py:1-2
"""
import importlib
import sys

import qsextra_b200
from qsextra_b200 import ExcitonicSystem, ChromophoreSystem  # noqa: F401

for _name in ('system', 'system.system', 'tools', 'tools.tools', 'tools.oracle', 'qcomo', 'qcomo.evolution',
              'qcomo.evolution.qevolve', 'qcomo.evolution.clevolve', 'spectroscopy', 'spectroscopy.spectroscopy',
              'spectroscopy.postprocessing', 'spectroscopy.simulation', 'spectroscopy.simulation.qspectroscopy',
              'spectroscopy.simulation.clspectroscopy'):
    _mod = importlib.import_module('qsextra_b200.' + _name)
    sys.modules['qsextra.' + _name] = _mod
    if '.' not in _name:
        globals()[_name] = _mod

__version__ = qsextra_b200.__version__
