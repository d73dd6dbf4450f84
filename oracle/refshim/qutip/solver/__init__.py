from qutip import _SolverResult as Result  # noqa: F401
