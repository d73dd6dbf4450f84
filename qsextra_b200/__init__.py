"""qsextra_b200 -- B200-native engine behind the qsextra system / qcomo / spectroscopy API.

``from qsextra_b200 import ExcitonicSystem, ChromophoreSystem`` mirrors ``qsextra/__init__.py:1`` of the reference;
the top-level ``qsextra`` package of this repository re-exports everything under the reference's import paths.
"""
from .system import ExcitonicSystem, ChromophoreSystem  # noqa: F401

__version__ = '0.1.0'
