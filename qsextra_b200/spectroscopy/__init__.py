from .spectroscopy import Spectroscopy, FeynmanDiagram  # noqa: F401  (reference: spectroscopy/__init__.py:1-2)
from .simulation import clspectroscopy, qspectroscopy, qspectroscopy_ensemble, qspectroscopy_set  # noqa: F401
