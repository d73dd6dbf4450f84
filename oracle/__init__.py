"""TEST INFRASTRUCTURE -- CPU oracle of the qsextra hot path.  Never imported by the product.

Contents
--------
circuits.py        restatement of the reference's circuit construction (qevolve.py:15-428, qspectroscopy.py:18-178)
                   as flat gate lists; lists the six Qiskit behaviours it relies on that cannot be verified offline
densesim.py        dense density-matrix simulator (exact expectation values of those gate lists)
reference_path.py  qevolve_exact, qspectroscopy_literal (circuit by circuit), qspectroscopy_operator (collective form)
csrc/ + fast.py    the same literal gate application in plain C + OpenMP (CPU baseline / ``bench.py --impl reference``)
postprocessing.py  NumPy/SciPy restatement of the reference's post-processing (spectroscopy/postprocessing.py:9-122)
lindblad.py        dense full-space master equation with the reference's operators (clevolve.py, clspectroscopy.py),
                   exact propagator exp(Lt) instead of QuTiP's ODE solver
refshim/           stand-ins for qiskit / qiskit_aer / qutip, used ONLY by tests/golden/make_golden.py to run the
                   UNMODIFIED reference code and write tests/golden/*.npz

Pinning
-------
The reference has no tests and its dependencies (qiskit==1.0.0, qiskit-aer==0.13.3, qutip>=4.7.5) cannot be installed
here, so parity against real Qiskit/Aer execution is UNPINNED.  What pins this oracle: every fixture in tests/golden/
(written by the reference's own code running over refshim) is reproduced to <= 1e-13, and the notebook known-answers of
SURVEY.md section 4 hold (tests/test_oracle_golden.py, tests/test_api.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
