from .result import Result  # noqa: F401
