"""TEST INFRASTRUCTURE (oracle) -- ctypes front-end of ``oracle/csrc/qsx_oracle.c`` (literal full-space gate
application in C, OpenMP over parameter sets).  Gate lists come from ``oracle/circuits.py``; this module only
converts labels to masks and calls C.  Used by tests/ and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import circuits as CIR
from .densesim import label_to_masks

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_build', 'libqsx_oracle.so')
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError('oracle C library missing: run `make -C oracle/csrc`')
        _lib = C.CDLL(_LIB_PATH)
        _lib.qsxo_populations.restype = C.c_int
    return _lib


def encode_ops(ops):
    """[('rot', label, theta) | ('reset', q)] -> (kind, x, z, ny, theta) arrays."""
    kind, x, z, ny, theta = [], [], [], [], []
    for op in ops:
        if op[0] == 'rot':
            mx, mz, n_y = label_to_masks(op[1])
            kind.append(0), x.append(mx), z.append(mz), ny.append(n_y), theta.append(op[2])
        elif op[0] == 'reset':
            kind.append(1), x.append(op[1]), z.append(0), ny.append(0), theta.append(0.0)
        else:
            raise ValueError(op[0])
    i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    return i32(kind), i32(x), i32(z), i32(ny), np.ascontiguousarray(theta, dtype=np.float64)


def populations_batch(systems, time, dt=None, Trotter_number=1, Trotter_order=1, coll_rates=None, n_threads=0):
    """Populations [B, T, N] of ``qevolve`` for a list of systems with identical structure and initial state type
    (coll_rates: one entry per system, or one shared value).  Returns (populations, threads used)."""
    systems = list(systems)
    B = len(systems)
    per_system_rates = coll_rates is not None and np.ndim(coll_rates) >= 1 and len(coll_rates) == B and \
        (not CIR.is_chromophore(systems[0]) or np.ndim(coll_rates) == 2)
    thetas, structure = [], None
    for b, s in enumerate(systems):
        rates = CIR.normalise_coll_rates(s, (coll_rates[b] if per_system_rates else coll_rates))
        if isinstance(rates, np.ndarray):
            rates = rates.tolist()
        dt_, dt_trotter, tn, counts = CIR.resolve_time_grid(time, dt, Trotter_number)
        layout = CIR.Layout(s, rates is not None)
        kind, x, z, ny, theta = encode_ops(CIR.step_ops(s, layout, dt_, dt_trotter, tn, Trotter_order, rates))
        if structure is None:
            structure = (kind, x, z, ny, layout, counts)
        elif not (np.array_equal(kind, structure[0]) and np.array_equal(x, structure[1]) and np.array_equal(z, structure[2])):
            raise ValueError('systems of a batch must produce the same gate structure')
        thetas.append(theta)
    kind, x, z, ny, layout, counts = structure
    theta = np.ascontiguousarray(np.array(thetas))
    from .densesim import DensityMatrix
    from .reference_path import run_ops
    dm = run_ops(DensityMatrix(layout.n_qubits), CIR.initial_state_ops(systems[0], layout))
    rho0 = np.ascontiguousarray(dm.rho, dtype=np.complex128)
    emit = np.ascontiguousarray(sorted(set(int(c) for c in counts)), dtype=np.int32)
    sysq = np.ascontiguousarray(layout.sys, dtype=np.int32)
    out = np.zeros((B, len(emit), len(sysq)))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    used = lib().qsxo_populations(C.c_int(layout.n_qubits), C.c_int(len(sysq)), p(sysq), C.c_int(len(kind)), p(kind),
                                  p(x), p(z), p(ny), p(theta), C.c_int(B), p(rho0), C.c_int(int(emit.max())),
                                  C.c_int(len(emit)), p(emit), p(out), C.c_int(n_threads))
    return out, used
