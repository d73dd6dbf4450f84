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
        _lib.qsxo_response_chains.restype = C.c_int
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
        elif op[0] in ('damp', 'dephase'):
            kind.append(2 if op[0] == 'damp' else 3), x.append(op[1]), z.append(0), ny.append(0), theta.append(op[2])
        else:
            raise ValueError(op[0])
    i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    return i32(kind), i32(x), i32(z), i32(ny), np.ascontiguousarray(theta, dtype=np.float64)


def kraus_form(ops, layout):
    """Replace every collision "gate(s) on (target, ancilla) + reset(ancilla)" of a literal step by the channel it
    induces on the target (the ancilla enters in |0> and is reset, so it can be traced out):

      rot X_m X_a (th), rot Y_m Y_a (th), reset a   ->  ('damp', m, 2 th)      qevolve.py:334-347, d = 2 pseudomodes
      rot Z_e X_a (th), reset a                     ->  ('dephase', e, th)     qevolve.py:302-304  (rzx(2 th))

    Labels lose the ancilla character (the ancilla is the highest qubit).  Returns None when the step contains any
    other use of the ancilla (d > 2 pseudomodes): the caller then keeps the literal list."""
    anc = layout.anc
    if anc is None:
        return list(ops)
    n = layout.n_qubits
    assert anc == n - 1

    def acts(label):                                     # {qubit: Pauli} of the non-identity characters
        return {n - 1 - pos: ch for pos, ch in enumerate(label) if ch != 'I'}
    out, i = [], 0
    while i < len(ops):
        op = ops[i]
        if op[0] == 'rot' and anc in acts(op[1]):
            a = acts(op[1])
            if len(a) == 2 and i + 2 < len(ops) and ops[i + 1][0] == 'rot' and ops[i + 2] == ('reset', anc):
                b = acts(ops[i + 1][1])
                (m,) = [q for q in a if q != anc]
                if a == {m: 'X', anc: 'X'} and b == {m: 'Y', anc: 'Y'} and op[2] == ops[i + 1][2]:
                    out.append(('damp', m, 2 * op[2]))
                    i += 3
                    continue
            if len(a) == 2 and i + 1 < len(ops) and ops[i + 1] == ('reset', anc):
                (e,) = [q for q in a if q != anc]
                if a == {e: 'Z', anc: 'X'}:
                    out.append(('dephase', e, op[2]))
                    i += 2
                    continue
            return None
        if op[0] == 'reset' and op[1] == anc:
            return None
        out.append(('rot', op[1][1:], op[2]) if op[0] == 'rot' else op)
        i += 1
    return out


def _step(system, dt, Trotter_number, Trotter_order, coll_rates, kraus):
    """(kind, x, z, ny, theta, layout, n_qubits of the state) of one propagation step."""
    rates = CIR.normalise_coll_rates(system, coll_rates)
    if isinstance(rates, np.ndarray):
        rates = rates.tolist()
    dt_, dt_trotter, tn, _ = CIR.resolve_time_grid(0.0, dt, Trotter_number)
    layout = CIR.Layout(system, rates is not None)
    ops = CIR.step_ops(system, layout, dt_, dt_trotter, tn, Trotter_order, rates)
    n_qubits = layout.n_qubits
    if kraus:
        reduced = kraus_form(ops, layout)
        if reduced is not None:
            ops = reduced
            n_qubits = layout.n_qubits - (1 if layout.anc is not None else 0)
    return encode_ops(ops) + (layout, n_qubits)


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


_KIND = {'X': 0, 'lower': 1, 'raise': 2}


def _interaction_kinds(side_seq, direction_seq):
    """(side, kind) of every interaction by the coefficient rules of qspectroscopy.py:144-174 (the same rules as
    ``reference_path.interaction_operators``): first interaction on a side and the emission apply X; otherwise
    |0><1| when (side == 'k') xor (direction == 'i'), else |1><0|."""
    first = {'k': True, 'b': True}
    out = []
    for n, (side, direction) in enumerate(zip(side_seq, direction_seq)):
        if first[side]:
            kind, first[side] = 'X', False
        elif n == len(side_seq) - 1:
            kind = 'X'
        elif (side == 'k') ^ (direction == 'i'):
            kind = 'lower'
        else:
            kind = 'raise'
        out.append((0 if side == 'k' else 1, _KIND[kind]))
    return out


def response_chains(system, side_seq, direction_seq, chain_steps, emit_steps, dt, Trotter_number=1, Trotter_order=1,
                    coll_rates=None, n_threads=0, kraus=False):
    """Response function along delay-grid chains with the C restatement (full space, literal gates).

    ``chain_steps`` [n_chains, n_delays - 1]: integer step counts of every delay but the last; ``emit_steps``: ascending
    step counts of the last delay shared by all chains.  ``kraus``: collisions as channels on a state without the
    ancilla qubit (``kraus_form``) instead of literal gates + reset.  -> (signal [n_chains, n_emit] complex, threads)."""
    kind, x, z, ny, theta, layout, n_qubits = _step(system, dt, Trotter_number, Trotter_order, coll_rates, kraus)
    inter = _interaction_kinds(side_seq, direction_seq)
    n_inter = len(inter)
    if n_inter > 2:
        chain_steps = np.ascontiguousarray(np.asarray(chain_steps, dtype=np.int32).reshape(-1, n_inter - 2))
        n_chains = chain_steps.shape[0]
    else:                                               # linear response: one chain, no earlier delays
        chain_steps, n_chains = np.zeros(1, dtype=np.int32), 1
    emit = np.ascontiguousarray(emit_steps, dtype=np.int32)
    assert np.all(np.diff(emit) >= 0)
    sides = np.ascontiguousarray([s for s, _ in inter], dtype=np.int32)
    kinds = np.ascontiguousarray([k for _, k in inter], dtype=np.int32)
    sysq = np.ascontiguousarray(layout.sys, dtype=np.int32)
    mu = np.ascontiguousarray(system.dipole_moments, dtype=np.float64)
    out = np.zeros((n_chains, len(emit)), dtype=np.complex128)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    used = lib().qsxo_response_chains(C.c_int(n_qubits), C.c_int(len(kind)), p(kind), p(x), p(z), p(ny), p(theta),
                                      C.c_int(n_inter), p(sides), p(kinds), C.c_int(len(sysq)), p(sysq), p(mu),
                                      C.c_int(n_chains), p(chain_steps), C.c_int(len(emit)), p(emit), p(out),
                                      C.c_int(n_threads))
    return out, used


def response_points(system, spectroscopy, index_chains, dt, n_threads=0, **kw):
    """Signal of ``spectroscopy`` on the rows ``[(i1, i2, ...)]`` of its delay grid: every entry fixes the indices of
    all delays but the last, the whole last axis is returned.  -> ([n_chains, len(last delay)] complex, threads)."""
    delays = spectroscopy.delay_time
    counts = [CIR.resolve_time_grid(np.asarray(d, dtype=float), dt, kw.get('Trotter_number', 1))[3] for d in delays]
    chains = [[int(counts[d][i]) for d, i in enumerate(idx)] for idx in index_chains] if len(delays) > 1 else [[]]
    order = np.argsort(counts[-1], kind='stable')
    sig, used = response_chains(system, spectroscopy.side_seq, spectroscopy.direction_seq, chains,
                                np.asarray(counts[-1])[order], dt, n_threads=n_threads, **kw)
    out = np.empty_like(sig)
    out[:, order] = sig
    return out, used
