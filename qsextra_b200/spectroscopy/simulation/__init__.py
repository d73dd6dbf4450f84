from .qspectroscopy import qspectroscopy, qspectroscopy_ensemble, qspectroscopy_set  # noqa: F401  (reference: spectroscopy/simulation/__init__.py:1-2)
from .clspectroscopy import clspectroscopy  # noqa: F401
