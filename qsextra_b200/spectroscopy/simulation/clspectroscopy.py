"""``clspectroscopy`` -- Liouville-space comparison path of the reference
(``qsextra/spectroscopy/simulation/clspectroscopy.py:134-181``, built on QuTiP).

Out of scope of the hot path (SURVEY.md section 8, row F3 "next"); kept as an importable name.
"""


def clspectroscopy(system, spectroscopy, verbose=True, **clevolve_kwds):
    raise NotImplementedError('clspectroscopy (QuTiP master-equation path) is outside the B200 hot path; '
                              'use qspectroscopy with a small dt.')
