from .system import ExcitonicSystem, ChromophoreSystem  # noqa: F401  (reference: qsextra/system/__init__.py:1)
