"""``qspectroscopy`` -- response functions on the delay grid, on the B200 engine (drop-in for the reference's
``qsextra/spectroscopy/simulation/qspectroscopy.py:193-253``).

Same signature and checks.  The reference estimates <X> + i<Y> of the ketbra ancilla of every pathway circuit
with ``shots`` samples (``Estimator(run_options={'shots': shots})``, qspectroscopy.py:114, :180-181).  The engine
computes the exact expectation value those estimates converge to (``engine/response.py``); with ``shots`` > 0 the
sampling error of the reference's estimator is added to it (``estimator_noise_sigma``), with ``shots`` 0 / None the
exact value is returned.

Checkpoints keep the reference's file format (qspectroscopy.py:103-111, :185-190; tools.py:120-137): a folder
``<uuid4 hex>/`` holding ``last_time_indices_number.npy`` and ``signal.npy``.  ``restart_from_checkpoint=<folder>`` of an
interrupted REFERENCE run is honoured: the grid points it had finished (C-order positions <= the saved number) are kept
from ``signal.npy``, the others are computed here; the folder is removed on success like the reference does.  A whole
grid is one batched computation here, so ``checkpoint=True`` has nothing to save in between and only mimics the
create/destroy of the folder.
"""
from __future__ import annotations

import os
import warnings

import numpy as np

from ...system import ChromophoreSystem, ExcitonicSystem
from ...tools.oracle import ending_sentence
from ...tools.tools import create_checkpoint_folder, destroy_checkpoint_folder
from ...engine.program import resolve_time_grid
from ...engine.response import response_function, response_function_set
from ...qcomo.evolution.qevolve import _program_for, _validate_rates, _validate_system
from ..spectroscopy import Spectroscopy

_RESERVED = ('system', 'time', 'shots', 'initialize_circuit', 'verbose')


def delay_step_counts(delay_time, dt, Trotter_number):
    """Integer step count of every delay (each delay goes through qevolve's own dt logic, qspectroscopy.py:133-139)."""
    counts = []
    grid = resolve_time_grid(0.0, dt, Trotter_number)
    for delays in delay_time:                           # dt is given, so every delay resolves against the same grid
        if len(delays) == 0:
            counts.append([])
            continue
        grid = resolve_time_grid(np.asarray(delays, dtype=float), dt, Trotter_number)
        counts.append([int(c) for c in grid[3]])
    return counts, grid


def pathway_coefficients(side_seq: str, direction_seq: str):
    """Coefficients of the circuits ONE site pathway expands into, for unit dipole moments -- the flag logic of
    qspectroscopy.py:142-176, flags mutating inside the per-circuit loop included (1 circuit for 'a', 3 for 'gsb',
    2 for 'se' and 'esa')."""
    first_k = first_b = True
    coeffs = [1.0 + 0j]
    n_int = len(side_seq)
    for n, side in enumerate(side_seq):
        new = []
        for c in coeffs:
            if side == 'b' and first_b:
                new.append(c)
                first_b = False
            elif side == 'k' and first_k:
                new.append(c)
                first_k = False
            elif n == n_int - 1:
                new.append(c)
            else:
                sign = 1j if (side == 'k') ^ (direction_seq[n] == 'i') else -1j
                new += [c / 2, c * sign / 2]
        coeffs = new
    return coeffs


def estimator_noise_sigma(side_seq: str, direction_seq: str, dipole_moments, shots: int) -> float:
    """Standard deviation (of the real and of the imaginary part) of the reference's shot-sampled signal at one grid
    point.  The reference adds coeff_c (X_c + i Y_c) over all site pathways and circuit variants c, X_c and Y_c being
    independent ``shots``-sample estimates of a +-1 observable: Var = (1 - <X_c>^2) / shots each.  The engine sums the
    pathways before it evolves anything, so the individual <X_c> do not exist here; their squares are dropped, which
    gives the variance for small single-pathway expectation values and an upper bound otherwise:
    sigma^2 = sum_c |coeff_c|^2 / shots = (sum_s mu_s^2)^(interactions) * sum_patterns |c|^2 / shots."""
    weight = sum(abs(c) ** 2 for c in pathway_coefficients(side_seq, direction_seq))
    mu2 = float(np.sum(np.asarray(dipole_moments, dtype=float) ** 2))
    return float(np.sqrt(weight * mu2 ** len(side_seq) / shots))


def add_estimator_noise(signal, side_seq, direction_seq, dipole_moments, shots, seed=None):
    """signal + complex normal sampling error of the reference's estimator (``estimator_noise_sigma``); NumPy arrays
    and CUDA tensors (noise drawn on the device) alike."""
    if not shots:
        return signal
    sigma = estimator_noise_sigma(side_seq, direction_seq, dipole_moments, int(shots))
    if isinstance(signal, np.ndarray):
        rng = np.random.default_rng(seed)
        return signal + sigma * (rng.standard_normal(signal.shape) + 1j * rng.standard_normal(signal.shape))
    import torch
    gen = torch.Generator(device=signal.device)
    gen.manual_seed(int(seed) if seed is not None else int(np.random.SeedSequence().generate_state(1)[0]))
    noise = torch.randn(tuple(signal.shape) + (2,), dtype=torch.float64, device=signal.device, generator=gen)
    return signal + sigma * torch.view_as_complex(noise)


def qspectroscopy(system: ChromophoreSystem | ExcitonicSystem,
                  spectroscopy: Spectroscopy,
                  shots: int = 1000,
                  verbose: bool = True,
                  checkpoint: bool = False,
                  restart_from_checkpoint: str | None = None,
                  device_output: bool = False,
                  shared: dict | None = None,
                  seed: int | None = None,
                  **qevolve_kwds,
                  ) -> np.ndarray:
    """Simulate ``spectroscopy`` (a ``FeynmanDiagram`` or a ``Spectroscopy`` with hand-set sequences) on ``system``.

    ``qevolve_kwds`` are forwarded to the evolution (``dt``, ``coll_rates``, ``Trotter_number``, ``Trotter_order``,
    ``GPU``); the reserved keys system/time/shots/initialize_circuit/verbose are dropped like the reference does.
    Returns a complex array with one axis per delay list.  Under ``torch.distributed`` the delay-grid chains are
    sharded over the ranks and every rank returns the full array.  ``device_output`` (an addition to the reference
    signature) returns the signal as a complex128 CUDA tensor, which ``postprocessing`` accepts as it is.  ``shared``
    (another addition): pass one dict to the calls for several diagrams of the same system, parameters and delay grid;
    the delay levels they have in common are then computed once (``engine.response.response_function``).
    ``shots`` > 0: the sampling error of the reference's estimator is added to the exact signal (``seed`` makes it
    reproducible); ``shots`` 0 or None: the exact signal.
    """
    _validate_system(system)
    if type(system) is ChromophoreSystem and system.mode_dict is None:
        warnings.warn("Pseudomodes not specified. Executing the dynamics with system as an ExcitonicSystem.")
        return qspectroscopy(system.extract_ExcitonicSystem(), spectroscopy, shots, verbose, checkpoint,
                             restart_from_checkpoint, device_output, shared, seed, **qevolve_kwds)
    for key in _RESERVED:
        qevolve_kwds.pop(key, None)
    if spectroscopy.side_seq == '' or spectroscopy.direction_seq == '':
        raise ValueError('Spectroscopy object is not correctly defined. Spectroscopy.side_seq or Spectroscopy.direction_seq are missing.')
    unknown = set(qevolve_kwds) - {'Trotter_number', 'Trotter_order', 'dt', 'coll_rates', 'GPU'}
    if unknown:
        raise TypeError("qevolve() got an unexpected keyword argument '{}'".format(sorted(unknown)[0]))

    coll_rates = _validate_rates(system, qevolve_kwds.get('coll_rates'))
    dt_in, trotter_in = qevolve_kwds.get('dt'), qevolve_kwds.get('Trotter_number', 1)
    mu = np.asarray(system.dipole_moments, dtype=float)
    # checkpoint of an interrupted run: read and validate it BEFORE any device work (qspectroscopy.py:103-111)
    folder, kept, last = None, None, -1
    shape = tuple(len(d) for d in spectroscopy.delay_time)
    if restart_from_checkpoint is not None:
        if not os.path.isdir(restart_from_checkpoint):
            raise FileNotFoundError('restart_from_checkpoint: no checkpoint folder {!r}'.format(restart_from_checkpoint))
        last = int(np.load(os.path.join(restart_from_checkpoint, 'last_time_indices_number.npy')))
        kept = np.load(os.path.join(restart_from_checkpoint, 'signal.npy'))
        if kept.shape != shape:
            raise ValueError('the checkpoint in {!r} holds a signal of shape {}, this delay grid has shape {} (folder '
                             'left untouched)'.format(restart_from_checkpoint, kept.shape, shape))
        folder = restart_from_checkpoint
    elif checkpoint:
        folder = create_checkpoint_folder()
        print(f'Checkpoint folder created ({folder})')
    if kept is not None and last + 1 >= kept.size:      # the interrupted run had finished every grid point
        signal = kept.astype(complex)
        if device_output:
            from ...engine.runtime import get_engine
            signal = get_engine().to_device(signal)
    elif dt_in is None:
        signal = _response_with_delay_sized_steps(system, spectroscopy, trotter_in, qevolve_kwds.get('Trotter_order', 1),
                                                  coll_rates, device_output)
    else:
        counts, (dt, dt_trotter, trotter_number, _) = delay_step_counts(spectroscopy.delay_time, dt_in, trotter_in)
        program = _program_for(system, dt, dt_trotter, trotter_number, qevolve_kwds.get('Trotter_order', 1), coll_rates)
        if shared is None:                              # (a one-diagram set: repeated evaluations replay a CUDA graph)
            signal = response_function_set(program, mu, [(spectroscopy.side_seq, spectroscopy.direction_seq)], counts,
                                           device_output=device_output)[0]
        else:
            signal = response_function(program, mu, spectroscopy.side_seq, spectroscopy.direction_seq, counts,
                                       device_output=device_output, shared=shared)
        signal = add_estimator_noise(signal, spectroscopy.side_seq, spectroscopy.direction_seq, mu, shots, seed)
    if kept is not None and last >= 0 and last + 1 < kept.size:
        flat, old = signal.reshape(-1), kept.reshape(-1)
        if device_output:
            import torch
            old = torch.as_tensor(old, device=signal.device)
        flat[:last + 1] = old[:last + 1]             # grid points the interrupted run had already finished
    if folder is not None:
        destroy_checkpoint_folder(folder)
    ending_sentence(verbose)
    return signal


def _response_with_delay_sized_steps(system, spectroscopy, trotter_number, trotter_order, coll_rates, device_output):
    """``dt=None``: the reference hands every scalar delay t to ``qevolve`` on its own, which then sets dt = t /
    Trotter_number and runs Trotter_number steps of that size, a collision round after each (qevolve.py:373-379,
    qspectroscopy.py:133-139); a zero delay divides by zero there.  Every distinct delay is therefore its own step
    program; the levels of the tree are evolved delay by delay (small grids: this is not the fast path)."""
    import torch
    from ...engine.response import interaction_plan
    from ...engine.runtime import get_engine
    from ...engine.sectors import Block
    eng = get_engine()
    plan = interaction_plan(spectroscopy.side_seq, spectroscopy.direction_seq)
    delays = spectroscopy.delay_time
    if len(plan) != len(delays) + 1:
        raise ValueError('side_seq must have one more entry than delay_time')
    for d in delays:
        for t in d:
            if t == 0:
                raise ZeroDivisionError('float division by zero')    # what time / dt raises in the reference (dt = 0)
    mu = np.asarray(system.dipole_moments, dtype=float)
    mu_dev = eng.to_device(mu, torch.float64)
    probe = _program_for(system, 1.0, 1.0, 1, trotter_order, coll_rates)
    n_sites, n_bits = probe.n_sites, probe.n_bits
    block = Block(n_sites, n_bits, 0, 0)
    ground = np.zeros((1, block.size), dtype=complex)
    ground[0, 0] = 1.0
    sigma = eng.to_device(ground)
    ka = kb = 0
    out = None
    for level, (side, kind) in enumerate(plan[:-1]):
        delta = +1 if kind == 'X' or (kind == 'raise') == (side == 'k') else -1
        ka, kb = (ka + delta, kb) if side == 'k' else (ka, kb + delta)
        new_block = Block(n_sites, n_bits, ka, kb)
        if not new_block.valid():
            shape = [len(d) for d in delays]
            return torch.zeros(shape, dtype=torch.complex128, device=eng.device) if device_output else np.zeros(shape, complex)
        sigma = eng.apply_dipole(block, new_block, 0 if side == 'k' else 1, mu_dev, sigma)
        block = new_block
        n_local = sigma.shape[0]
        last = level == len(plan) - 2
        pieces = []
        for t in delays[level]:
            step = float(t) / trotter_number
            program = _program_for(system, step, step, 1, trotter_order, coll_rates)
            dev = eng.upload_program(program.bind(block), program)
            if last:
                obs = eng.upload_observables(*block.emission_observable(mu), kind='emission', mu=mu)
                got, _ = eng.evolve(dev, sigma, n_local, trotter_number, [trotter_number], obs=obs)
                pieces.append(got.reshape(n_local, 1))
            else:
                _, snap = eng.evolve(dev, sigma, n_local, trotter_number, [trotter_number], snapshots=True)
                pieces.append(snap.reshape(n_local, 1, block.size))
        if last:
            out = torch.cat(pieces, dim=1)
        else:
            sigma = torch.cat(pieces, dim=1).reshape(n_local * len(pieces), block.size).contiguous()
    table = out.reshape([len(d) for d in delays])
    return table.contiguous() if device_output else eng.to_host(table)


def qspectroscopy_set(system: ChromophoreSystem | ExcitonicSystem,
                      spectroscopies,
                      shots: int = 0,
                      verbose: bool = True,
                      device_output: bool = False,
                      seed: int | None = None,
                      **qevolve_kwds,
                      ) -> list:
    """Several diagrams of one spectrum at once (new API; the reference calls ``qspectroscopy`` per diagram).

    ``spectroscopies``: ``FeynmanDiagram`` / ``Spectroscopy`` objects with the SAME ``delay_time``.  The delay levels
    the diagrams have in common are computed once and the read-out propagations of all of them run concurrently with
    the rest (``engine.response.response_function_set``).  Returns one signal per entry, each identical to what
    ``qspectroscopy(shots=0)`` returns for it; ``shots`` > 0 adds the estimator's sampling error like ``qspectroscopy``
    does (default here: exact).
    """
    spectroscopies = list(spectroscopies)
    if not spectroscopies:
        return []
    _validate_system(system)
    if type(system) is ChromophoreSystem and system.mode_dict is None:
        warnings.warn("Pseudomodes not specified. Executing the dynamics with system as an ExcitonicSystem.")
        return qspectroscopy_set(system.extract_ExcitonicSystem(), spectroscopies, shots, verbose, device_output, seed,
                                 **qevolve_kwds)
    for key in _RESERVED:
        qevolve_kwds.pop(key, None)
    delays = spectroscopies[0].delay_time
    for sp in spectroscopies:
        if sp.side_seq == '' or sp.direction_seq == '':
            raise ValueError('Spectroscopy object is not correctly defined. Spectroscopy.side_seq or Spectroscopy.direction_seq are missing.')
        if len(sp.delay_time) != len(delays) or any(not np.array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))
                                                     for a, b in zip(sp.delay_time, delays)):
            raise ValueError('qspectroscopy_set needs the same delay_time in every entry')
    unknown = set(qevolve_kwds) - {'Trotter_number', 'Trotter_order', 'dt', 'coll_rates', 'GPU'}
    if unknown:
        raise TypeError("qevolve() got an unexpected keyword argument '{}'".format(sorted(unknown)[0]))
    coll_rates = _validate_rates(system, qevolve_kwds.get('coll_rates'))
    if qevolve_kwds.get('dt') is None:
        raise ValueError('qspectroscopy needs an explicit dt (with dt=None the reference derives dt from each scalar '
                         'delay and fails at zero delay).')
    counts, (dt, dt_trotter, trotter_number, _) = delay_step_counts(delays, qevolve_kwds['dt'], qevolve_kwds.get('Trotter_number', 1))
    program = _program_for(system, dt, dt_trotter, trotter_number, qevolve_kwds.get('Trotter_order', 1), coll_rates)
    mu = np.asarray(system.dipole_moments, dtype=float)
    signals = response_function_set(program, mu, [(sp.side_seq, sp.direction_seq) for sp in spectroscopies], counts,
                                    device_output=device_output)
    if shots:
        signals = [add_estimator_noise(sig, sp.side_seq, sp.direction_seq, mu, shots, None if seed is None else seed + k)
                   for k, (sp, sig) in enumerate(zip(spectroscopies, signals))]
    ending_sentence(verbose)
    return signals


def ensemble_program(systems, spectroscopies, members=None, **qevolve_kwds):
    """Checks of ``qspectroscopy_ensemble`` and the objects its evaluation is made of:
    -> (StepProgram over the members, dipole moments, [(side_seq, direction_seq)], step counts per delay).
    ``members = (lo, hi)``: every system is checked, but only the members lo .. hi-1 are lowered (the share of one rank;
    ``None`` when the share is empty)."""
    from ...engine.program import StepProgram, SystemBatch
    systems, spectroscopies = list(systems), list(spectroscopies)
    for s_ in systems:
        _validate_system(s_)
        if type(s_) is not type(systems[0]):
            raise TypeError('all members of an ensemble must be of the same class')
    for key in _RESERVED:
        qevolve_kwds.pop(key, None)
    unknown = set(qevolve_kwds) - {'Trotter_number', 'Trotter_order', 'dt', 'coll_rates', 'GPU'}
    if unknown:
        raise TypeError("qevolve() got an unexpected keyword argument '{}'".format(sorted(unknown)[0]))
    delays = spectroscopies[0].delay_time
    for sp in spectroscopies:
        if sp.side_seq == '' or sp.direction_seq == '':
            raise ValueError('Spectroscopy object is not correctly defined. Spectroscopy.side_seq or Spectroscopy.direction_seq are missing.')
        if len(sp.delay_time) != len(delays) or any(not np.array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))
                                                     for a, b in zip(sp.delay_time, delays)):
            raise ValueError('qspectroscopy_ensemble needs the same delay_time in every entry')
    mu = np.asarray(systems[0].dipole_moments, dtype=float)
    if any(not np.array_equal(np.asarray(s_.dipole_moments, dtype=float), mu) for s_ in systems):
        raise ValueError('the members of an ensemble must have the same dipole moments')
    if qevolve_kwds.get('dt') is None:
        raise ValueError('qspectroscopy needs an explicit dt')
    B = len(systems)
    rates_in = qevolve_kwds.get('coll_rates')
    rates = None
    if rates_in is not None:
        arr = np.asarray(rates_in, dtype=float)
        per_member = arr.ndim == 2 or (arr.ndim == 1 and type(systems[0]) is ExcitonicSystem and arr.size == B and B > 1)
        rows = [_validate_rates(s_, (arr[b].tolist() if arr.ndim == 2 else float(arr[b])) if per_member else rates_in)
                for b, s_ in enumerate(systems)]
        rates = np.array([np.atleast_1d(np.asarray(r, dtype=float)) for r in rows])
        if type(systems[0]) is ExcitonicSystem:
            rates = rates.reshape(B)
    counts, (dt, dt_trotter, trotter_number, _) = delay_step_counts(delays, qevolve_kwds['dt'], qevolve_kwds.get('Trotter_number', 1))
    if members is not None:
        lo, hi = members
        systems = systems[lo:hi]
        rates = None if rates is None else rates[lo:hi]
    program = None
    if systems:
        program = StepProgram(SystemBatch.from_systems(systems), dt, trotter_number, qevolve_kwds.get('Trotter_order', 1),
                              rates, dt_trotter=dt_trotter)
    return program, mu, [(sp.side_seq, sp.direction_seq) for sp in spectroscopies], counts


def qspectroscopy_ensemble(systems,
                           spectroscopies,
                           shots: int = 0,
                           verbose: bool = True,
                           device_output: bool = False,
                           average: bool = False,
                           seed: int | None = None,
                           host_out=None,
                           **qevolve_kwds,
                           ) -> list:
    """The diagrams ``spectroscopies`` for an ENSEMBLE of systems in one batched evaluation (new API: the reference
    loops ``qspectroscopy`` over realisations by hand, e.g. for static disorder).

    ``systems``: ExcitonicSystem / ChromophoreSystem objects with the same structure (sites, pseudomode levels, which
    couplings are non-zero) and the same dipole moments; energies, couplings and bath parameters may differ.
    ``coll_rates``: one value (list) for all members, or one row per member.  The members are the parameter sets of
    one step program: every level of the response tree holds all of them, each propagates its own read-out, the last
    level is one batched product (``engine.response``).  Under ``torch.distributed`` the MEMBERS are sharded over the
    ranks (contiguous slices) and one all-gather assembles the result on every rank.
    Returns one array per diagram with a leading ensemble axis, ``[len(systems), *grid]`` (``average=True``: the
    ensemble mean, ``[*grid]``).

    ``host_out``: one pinned complex128 CPU tensor per diagram, shaped ``[members of THIS rank, *grid]`` (all members
    without ``torch.distributed``).  The exact signals (before any estimator noise) of the rank's own members are
    copied into them while the evaluation still runs -- each diagram as soon as it is complete, so that only the last
    diagram's copy is left at the end -- and the call returns after the copies have finished.  Together with
    ``device_output=True`` this is the fastest way to get the result into host memory."""
    systems, spectroscopies = list(systems), list(spectroscopies)
    if not systems or not spectroscopies:
        return []
    if type(systems[0]) is ChromophoreSystem and systems[0].mode_dict is None:
        for s_ in systems:
            _validate_system(s_)
        warnings.warn("Pseudomodes not specified. Executing the dynamics with system as an ExcitonicSystem.")
        return qspectroscopy_ensemble([s_.extract_ExcitonicSystem() for s_ in systems], spectroscopies, shots, verbose,
                                      device_output, average, seed, host_out, **qevolve_kwds)
    from ...engine import dist as qdist
    from ...engine.response import HostRows, _finish, empty_share
    B = len(systems)
    rank, size = qdist.world()
    rows = None
    if host_out is not None:
        if len(host_out) != len(spectroscopies):
            raise ValueError('host_out needs one tensor per diagram')
        rows = HostRows(host_out)
    if size > 1:                                        # every rank lowers and uploads its own members only
        share = qdist.shard_bounds(B, rank, size)
        program, mu, sequences, counts = ensemble_program(systems, spectroscopies, members=share, **qevolve_kwds)
        if program is None:                             # more ranks than members: this one only joins the gathers
            from ...engine.runtime import get_engine
            eng = get_engine()
            signals = [_finish(eng, out, info, device_output) for out, info in zip(*empty_share(eng, sequences, counts, B))]
        else:
            signals = response_function_set(program, mu, sequences, counts, device_output=device_output, global_sets=B,
                                            host_out=rows)
    else:
        program, mu, sequences, counts = ensemble_program(systems, spectroscopies, **qevolve_kwds)
        signals = response_function_set(program, mu, sequences, counts, device_output=device_output, host_out=rows)
    if B == 1 and size == 1:                            # a one-member ensemble still carries the ensemble axis
        signals = [sig[None] for sig in signals]
    if rows is not None:
        import torch
        lo, hi = qdist.shard_bounds(B, rank, size)
        for k, (target, done) in enumerate(zip(rows.tensors, rows.copied)):
            if target is not None and not done:         # (not copied on the way: from the finished signal)
                mine = signals[k][lo:hi]
                target.copy_(mine if isinstance(mine, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(mine)),
                             non_blocking=True)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
    if shots:                                           # every member is its own set of estimator runs
        signals = [add_estimator_noise(sig, a, b, mu, shots, None if seed is None else seed + k)
                   for k, ((a, b), sig) in enumerate(zip(sequences, signals))]
    if average:
        signals = [sig.mean(0) for sig in signals]
    ending_sentence(verbose)
    return signals
