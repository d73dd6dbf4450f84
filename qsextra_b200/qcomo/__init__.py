from .evolution import qevolve, qevolve_batch, clevolve  # noqa: F401  (reference: qcomo/__init__.py:1)
