from .tools import *  # noqa: F401,F403  (reference: qsextra/tools/__init__.py:1)
from .tools import H_BAR_EV_FS  # noqa: F401
