"""``clevolve`` -- continuous-time (Lindblad) comparison path of the reference
(``qsextra/qcomo/evolution/clevolve.py:110-192``, QuTiP ``sesolve``/``mesolve``).

Out of scope of the hot path (SURVEY.md section 8, row F3 "next"): the import path and signature are kept so that
scripts importing it keep loading, but calling it raises until the GPU Lindblad integrator lands.
"""


def clevolve(system, time, rates=None, measure_populations=True, state_overwrite=None, verbose=True):
    raise NotImplementedError('clevolve (QuTiP master-equation path) is outside the B200 hot path; '
                              'use qevolve with a small dt for the collision-model dynamics.')
