"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's circuit construction.

Every function restates one piece of ``/root/reference/qsextra/qcomo/evolution/qevolve.py`` or
``.../spectroscopy/simulation/qspectroscopy.py`` as a flat list of primitive operations on numbered
qubits, in the order the reference emits them.  Operations:

    ('rot', label, theta)      exp(-i theta P), label over all circuit qubits, highest qubit leftmost
    ('reset', q)               trace out qubit q, replace by |0>
    ('x', q) ('h', q) ('cx', c, t) ('cy', c, t) ('init', amplitudes, qubits)

``system`` is duck-typed: anything with the reference's attribute names (``system_size, e_el, coupl_el,
dipole_moments, state_type, state`` and, for chromophore systems, ``mode_dict``).

Qiskit behaviours this file relies on and that cannot be verified offline (qiskit==1.0.0 is absent):
 (1) PauliEvolutionGate + LieTrotter(reps=1): terms applied in SparsePauliOp order, each as
     exp(-i * time * Re(coeff) * P); an all-identity term is a global phase.
 (2) SparsePauliOp.simplify(): duplicates summed in first-appearance order, |c| <= 1e-8 dropped.
 (3) ``a ^ b``: b on the low qubits, a-major term order.
 (4) SuzukiTrotter(order 2): half steps of terms[1:] reversed, full step of terms[0], mirrored half steps;
     order 2k by the p = 1/(4 - 4^(1/(2k-1))) recursion; odd orders raise ValueError.
 (5) rzx(theta, q0, q1) = exp(-i theta/2 Z_q0 X_q1).
 (6) reset = trace out and replace by |0>.
(1)-(4) change results only when the terms inside one gate do not commute: d >= 3 pseudomodes or order >= 2.
"""
from __future__ import annotations

import itertools
from math import prod

import numpy as np


# ---------------------------------------------------------------------------------------------
# register layout  (qevolve.py:15-43, qspectroscopy.py:18-50)
# ---------------------------------------------------------------------------------------------
def is_chromophore(system) -> bool:
    return getattr(system, 'mode_dict', None) is not None


def mode_register_size(d_k: int) -> int:
    """qevolve.py:28-35: log2(d) rounded if within 1e-10 of an integer, else ceil."""
    nq = np.log2(d_k)
    if abs(np.round(nq) - nq) > 1e-10:
        return int(np.ceil(nq))
    return int(np.round(nq))


class Layout:
    """Qubit numbers of one reference circuit.  ``offset`` = 1 in spectroscopy (ketbra is qubit 0)."""

    def __init__(self, system, has_ancilla: bool, offset: int = 0):
        n = system.system_size
        self.offset = offset
        self.sys = [offset + i for i in range(n)]
        nxt = offset + n
        self.mode = {}                       # (i, k) -> list of qubits, lowest first
        if is_chromophore(system):
            levels = system.mode_dict['lvl_mode']
            for i in range(n):
                for k in range(len(system.mode_dict['omega_mode'])):
                    q = mode_register_size(levels[k])
                    self.mode[(i, k)] = list(range(nxt, nxt + q))
                    nxt += q
        self.anc = nxt if has_ancilla else None
        self.n_qubits = nxt + (1 if has_ancilla else 0)

    def embed(self, label: str, qubits: list[int]) -> str:
        """Label on ``qubits`` (label qubit j -> qubits[j]) as a full-width label."""
        chars = ['I'] * self.n_qubits
        k = len(label)
        for j in range(k):
            chars[self.n_qubits - 1 - qubits[j]] = label[k - 1 - j]
        return ''.join(chars)


# ---------------------------------------------------------------------------------------------
# Trotter formulas  (qevolve.py:67-72 + Qiskit behaviours (1), (4))
# ---------------------------------------------------------------------------------------------
def product_formula(terms: list[tuple[str, float]], time: float, order: int) -> list[tuple[str, float]]:
    """[(label, coeff)] -> [(label, angle)] with exp(-i angle P) applied in list order."""
    terms = [(l, float(np.real(c))) for l, c in terms]
    if order == 1:
        return [(l, c * time) for l, c in terms]
    if order % 2 == 1:
        raise ValueError('Suzuki product formulae are symmetric and therefore only defined for even orders.')

    def recurse(o, t):
        if o == 2:
            halves = [(l, c * t / 2) for l, c in reversed(terms[1:])]
            return halves + [(terms[0][0], terms[0][1] * t)] + halves[::-1]
        p = 1 / (4 - 4 ** (1 / (o - 1)))
        outer = recurse(o - 2, p * t)
        return outer + outer + recurse(o - 2, (1 - 4 * p) * t) + outer + outer
    return recurse(order, time)


def simplify_terms(terms: list[tuple[str, complex]]) -> list[tuple[str, complex]]:
    """Qiskit behaviour (2)."""
    order, sums = [], {}
    for l, c in terms:
        if l not in sums:
            order.append(l)
            sums[l] = 0j
        sums[l] += c
    return [(l, sums[l]) for l in order if not np.isclose(sums[l], 0, atol=1e-8, rtol=1e-5)]


def tensor_terms(a, b):
    """Qiskit behaviour (3): ``a ^ b``."""
    return [(la + lb, ca * cb) for la, ca in a for lb, cb in b]


# ---------------------------------------------------------------------------------------------
# A3  system block  (qevolve.py:74-105)
# ---------------------------------------------------------------------------------------------
def system_terms(system) -> list[tuple[str, float]]:
    n = system.system_size
    eps, coup = system.e_el, system.coupl_el
    terms = {}
    for i in range(n):
        if eps[i] != 0.:
            terms['I' * (n - i - 1) + 'Z' + 'I' * i] = -eps[i] / 2
    for i in range(n):
        for j in range(i + 1, n):
            if coup[i, j] != 0.:
                for p in 'XY':
                    terms['I' * (n - j - 1) + p + 'I' * (j - i - 1) + p + 'I' * i] = coup[i, j] / 2
    return list(terms.items())


# ---------------------------------------------------------------------------------------------
# A4  pseudomode operators in Gray code  (qevolve.py:107-195, tools.py:77-118)
# ---------------------------------------------------------------------------------------------
def gray_strings(nq: int) -> list[str]:
    return [format(i ^ (i >> 1), '0{}b'.format(nq)) for i in range(1 << nq)]


def mode_operator_terms(d_k: int):
    """Pauli dictionaries of a, a^dagger, a + a^dagger, a^dagger a for a d_k-level mode on
    ``mode_register_size(d_k)`` qubits; level l <-> Gray(l), leftmost character = highest qubit."""
    nq = mode_register_size(d_k)
    enc = gray_strings(nq)

    def accumulate(target, ops, coefs, mul):
        labels = [''.join(t) for t in itertools.product(*ops)]
        values = [mul * prod(t) for t in itertools.product(*coefs)]
        for l, v in zip(labels, values):
            target[l] = target[l] + v if l in target else v

    a, adag, num = {}, {}, {}
    for index in range(1, d_k):                       # a = sum sqrt(l) |l-1><l|      (:152-163)
        ops, coefs = [], []
        for pos in range(nq):
            lo = enc[index - 1][pos]
            if lo == enc[index][pos]:
                ops.append(['I', 'Z'])
                coefs.append([0.5, 0.5] if lo == '0' else [0.5, -0.5])
            else:
                ops.append(['X', 'Y'])
                coefs.append([0.5, 0.5j] if lo == '0' else [0.5, -0.5j])
        accumulate(a, ops, coefs, np.sqrt(index))
    for index in range(d_k - 1):                      # a^dagger = sum sqrt(l+1) |l+1><l|   (:166-177)
        ops, coefs = [], []
        for pos in range(nq):
            hi = enc[index + 1][pos]
            if hi == enc[index][pos]:
                ops.append(['I', 'Z'])
                coefs.append([0.5, 0.5] if hi == '0' else [0.5, -0.5])
            else:
                ops.append(['X', 'Y'])
                coefs.append([0.5, 0.5j] if hi == '0' else [0.5, -0.5j])
        accumulate(adag, ops, coefs, np.sqrt(index + 1))
    a_plus_adag = {k: a[k] + adag[k] for k in a if a[k] != -adag[k]}       # (:180-183)
    for index in range(d_k):                          # a^dagger a = sum l |l><l|      (:186-193)
        ops = [['I', 'Z']] * nq
        coefs = [[0.5, 0.5] if enc[index][pos] == '0' else [0.5, -0.5] for pos in range(nq)]
        accumulate(num, ops, coefs, index)
    return a, adag, a_plus_adag, num


# ---------------------------------------------------------------------------------------------
# A5 + step assembly  (qevolve.py:197-290)
# ---------------------------------------------------------------------------------------------
def trotter_step_ops(system, layout: Layout, dt: float, order: int) -> list[tuple]:
    ops = []
    for label, angle in product_formula(system_terms(system), dt, order):          # :264-267
        ops.append(('rot', layout.embed(label, layout.sys), angle))
    if not is_chromophore(system):
        return ops
    md = system.mode_dict
    for k in range(len(md['omega_mode'])):                                          # :272
        _, _, a_plus_adag, num = mode_operator_terms(md['lvl_mode'][k])
        mode_rot, int_rot = [], []
        if md['omega_mode'][k] != 0:                                                # :210-217
            mode_rot = product_formula([(l, md['omega_mode'][k] * c) for l, c in num.items()], dt, order)
        if md['coupl_ep'][k] != 0:                                                  # :239-252
            h = tensor_terms(list(a_plus_adag.items()), [('I', 0.5), ('Z', -0.5)])
            h = [(l, md['coupl_ep'][k] * c) for l, c in simplify_terms(h)]
            int_rot = product_formula(h, dt, order)
        for i in range(system.system_size):                                         # :278-289
            mq = layout.mode[(i, k)]
            for label, angle in mode_rot:
                ops.append(('rot', layout.embed(label, mq), angle))
            for label, angle in int_rot:
                ops.append(('rot', layout.embed(label, [layout.sys[i]] + mq), angle))
    return ops


# ---------------------------------------------------------------------------------------------
# A6 / A7  collision round  (qevolve.py:292-348)
# ---------------------------------------------------------------------------------------------
def collision_ops(system, layout: Layout, dt: float, order: int, coll_rates) -> list[tuple]:
    ops = []
    if not is_chromophore(system):
        theta = 2 * np.sqrt(coll_rates) * np.sqrt(dt)                               # :303
        for i in range(system.system_size):
            ops.append(('rot', layout.embed('XZ', [layout.sys[i], layout.anc]), theta / 2))   # rzx: Z_e X_anc
            ops.append(('reset', layout.anc))
        return ops
    md = system.mode_dict
    for k in range(len(md['omega_mode'])):                                          # :328
        if coll_rates[k] == 0:
            continue
        a, adag, _, _ = mode_operator_terms(md['lvl_mode'][k])
        h = (tensor_terms(list(a.items()), [('X', 0.5), ('Y', -0.5j)])
             + tensor_terms(list(adag.items()), [('X', 0.5), ('Y', 0.5j)]))         # :337
        h = [(l, np.sqrt(coll_rates[k] / dt) * c) for l, c in simplify_terms(h)]    # :338-339
        rots = product_formula(h, dt, order)
        for i in range(system.system_size):                                         # :342-347
            for label, angle in rots:
                ops.append(('rot', layout.embed(label, [layout.anc] + layout.mode[(i, k)]), angle))
            ops.append(('reset', layout.anc))
    return ops


# ---------------------------------------------------------------------------------------------
# A2  initial state  (qevolve.py:45-65)
# ---------------------------------------------------------------------------------------------
def initial_state_ops(system, layout: Layout) -> list[tuple]:
    st = system.state_type
    if st == 'state':
        return [('init', np.asarray(system.state, dtype=complex), list(layout.sys))]
    if st == 'delocalized excitation':
        relevant = [i for i, c in enumerate(system.state) if c != 0]
        amp = np.zeros(2 ** len(relevant), dtype=complex)
        for ni, i in enumerate(relevant):
            amp[2 ** ni] = system.state[i]
        return [('init', amp, [layout.sys[i] for i in relevant])]
    if st == 'localized excitation':
        return [('x', layout.sys[system.state])]
    return []


# ---------------------------------------------------------------------------------------------
# A9  time-step logic  (qevolve.py:362-428, :504-512)
# ---------------------------------------------------------------------------------------------
def normalise_coll_rates(system, coll_rates):
    if not is_chromophore(system):
        if coll_rates is not None and not np.isscalar(coll_rates):
            raise TypeError('For an ExcitonicSystem, coll_rates must be a scalar.')
        return coll_rates
    if coll_rates is None:
        return None
    rates = [coll_rates] if np.isscalar(coll_rates) else list(coll_rates)
    if len(rates) != len(system.mode_dict['omega_mode']):
        raise ValueError('The number of dephasing rates is different from the number of pseudomodes per chromophore.')
    return rates


def resolve_time_grid(time, dt, trotter_number):
    """-> (dt, dt_trotter, trotter_number, step count of every entry of ``time``)."""
    if dt is None:
        dt_trotter = (time if np.isscalar(time) else time[1] - time[0]) / trotter_number
        dt, trotter_number = dt_trotter, 1
    else:
        dt_trotter = dt / trotter_number
    times = [time] if np.isscalar(time) else list(time)
    if not all(abs(t - np.round(t / dt) * dt) < 1e-10 for t in times):
        raise ValueError('Entries in time are not multiples of dt_collisions.')
    return dt, dt_trotter, trotter_number, np.round(np.array(times) / dt).astype(int)


def step_ops(system, layout: Layout, dt, dt_trotter, trotter_number, order, coll_rates) -> list[tuple]:
    """One propagation step: Trotter_number Trotter steps, then one collision round (:419-422)."""
    ops = trotter_step_ops(system, layout, dt_trotter, order) * trotter_number
    if coll_rates is not None:
        ops = ops + collision_ops(system, layout, dt, order, coll_rates)
    return ops


# ---------------------------------------------------------------------------------------------
# A10 / A11  Feynman-pathway circuits  (qspectroscopy.py:18-50, :65-80, :117-178)
# ---------------------------------------------------------------------------------------------
def dipole_ops(kb: int, target: int, side: str, op: str) -> list[tuple]:
    """Controlled X/Y from the ketbra qubit, X-conjugated for the bra side (:65-80)."""
    gate = ('cx', kb, target) if op == 'X' else ('cy', kb, target)
    return [('x', kb), gate, ('x', kb)] if side == 'b' else [gate]


def pathway_circuits(system, side_seq: str, direction_seq: str, evolution_ops, site_pathway):
    """Expand ONE site pathway into its circuits and coefficients, with the reference's flag logic
    (flags mutate inside the per-circuit loop, :142-176).

    ``evolution_ops[n]`` = list of ops evolving for the n-th delay.  Returns [(ops, coeff)]."""
    mu = system.dipole_moments
    first_k = first_b = True
    circuits = [([('h', 0)], 1.)]
    n_int = len(site_pathway)
    for n in range(n_int):
        last = n == n_int - 1
        evo = [] if last else evolution_ops[n]
        site, side = site_pathway[n], side_seq[n]
        new = []
        for ops, coeff in circuits:
            if side == 'b' and first_b:
                new.append((ops + dipole_ops(0, 1 + site, side, 'X') + evo, coeff * mu[site]))
                first_b = False
            elif side == 'k' and first_k:
                new.append((ops + dipole_ops(0, 1 + site, side, 'X') + evo, coeff * mu[site]))
                first_k = False
            elif last:
                new.append((ops + dipole_ops(0, 1 + site, side, 'X') + evo, coeff * mu[site]))
            else:
                for op in 'XY':
                    if op == 'X':
                        c = coeff * mu[site] / 2
                    elif (side == 'k') ^ (direction_seq[n] == 'i'):
                        c = coeff * mu[site] * (1.j / 2)
                    else:
                        c = coeff * mu[site] * (-1.j / 2)
                    new.append((ops + dipole_ops(0, 1 + site, side, op) + evo, c))
        circuits = new
    return circuits
