from .qevolve import qevolve, qevolve_batch  # noqa: F401  (reference: qcomo/evolution/__init__.py:1-2)
from .clevolve import clevolve  # noqa: F401
