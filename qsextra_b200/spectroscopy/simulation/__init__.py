from .qspectroscopy import qspectroscopy  # noqa: F401  (reference: spectroscopy/simulation/__init__.py:1-2)
from .clspectroscopy import clspectroscopy  # noqa: F401
