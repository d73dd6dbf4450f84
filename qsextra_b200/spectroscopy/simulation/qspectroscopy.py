"""``qspectroscopy`` -- response functions on the delay grid, on the B200 engine (drop-in for the reference's
``qsextra/spectroscopy/simulation/qspectroscopy.py:193-253``).

Same signature and checks.  The reference estimates <X> + i<Y> of the ketbra ancilla of every pathway circuit
with ``shots`` samples; this implementation returns the exact expectation value those estimates converge to
(``engine/response.py``), so ``shots`` only matters when it is falsy/positive in the reference sense: it is accepted
and otherwise ignored.

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


def qspectroscopy(system: ChromophoreSystem | ExcitonicSystem,
                  spectroscopy: Spectroscopy,
                  shots: int = 1000,
                  verbose: bool = True,
                  checkpoint: bool = False,
                  restart_from_checkpoint: str | None = None,
                  device_output: bool = False,
                  shared: dict | None = None,
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
    """
    _validate_system(system)
    if type(system) is ChromophoreSystem and system.mode_dict is None:
        warnings.warn("Pseudomodes not specified. Executing the dynamics with system as an ExcitonicSystem.")
        return qspectroscopy(system.extract_ExcitonicSystem(), spectroscopy, shots, verbose, checkpoint,
                             restart_from_checkpoint, device_output, shared, **qevolve_kwds)
    for key in _RESERVED:
        qevolve_kwds.pop(key, None)
    if spectroscopy.side_seq == '' or spectroscopy.direction_seq == '':
        raise ValueError('Spectroscopy object is not correctly defined. Spectroscopy.side_seq or Spectroscopy.direction_seq are missing.')
    unknown = set(qevolve_kwds) - {'Trotter_number', 'Trotter_order', 'dt', 'coll_rates', 'GPU'}
    if unknown:
        raise TypeError("qevolve() got an unexpected keyword argument '{}'".format(sorted(unknown)[0]))

    coll_rates = _validate_rates(system, qevolve_kwds.get('coll_rates'))
    dt_in, trotter_in = qevolve_kwds.get('dt'), qevolve_kwds.get('Trotter_number', 1)
    if dt_in is None:
        raise ValueError('qspectroscopy needs an explicit dt (with dt=None the reference derives dt from each scalar '
                         'delay and fails at zero delay).')
    counts, (dt, dt_trotter, trotter_number, _) = delay_step_counts(spectroscopy.delay_time, dt_in, trotter_in)
    program = _program_for(system, dt, dt_trotter, trotter_number, qevolve_kwds.get('Trotter_order', 1), coll_rates)
    folder, kept, last = None, None, -1
    if restart_from_checkpoint is not None and os.path.exists(restart_from_checkpoint):
        folder = restart_from_checkpoint
        last = int(np.load(os.path.join(folder, 'last_time_indices_number.npy')))
        kept = np.load(os.path.join(folder, 'signal.npy'))
    elif checkpoint:
        folder = create_checkpoint_folder()
        print(f'Checkpoint folder created ({folder})')
    signal = response_function(program, np.asarray(system.dipole_moments, dtype=float), spectroscopy.side_seq,
                               spectroscopy.direction_seq, counts, device_output=device_output, shared=shared)
    if kept is not None and kept.shape == signal.shape and last >= 0:
        flat, old = signal.reshape(-1), kept.reshape(-1)
        if device_output:
            import torch
            old = torch.as_tensor(old, device=signal.device)
        flat[:last + 1] = old[:last + 1]             # grid points the interrupted run had already finished
    if folder is not None:
        destroy_checkpoint_folder(folder)
    ending_sentence(verbose)
    return signal


def qspectroscopy_set(system: ChromophoreSystem | ExcitonicSystem,
                      spectroscopies,
                      shots: int = 1000,
                      verbose: bool = True,
                      device_output: bool = False,
                      **qevolve_kwds,
                      ) -> list:
    """Several diagrams of one spectrum at once (new API; the reference calls ``qspectroscopy`` per diagram).

    ``spectroscopies``: ``FeynmanDiagram`` / ``Spectroscopy`` objects with the SAME ``delay_time``.  The delay levels
    the diagrams have in common are computed once and the read-out propagations of all of them run concurrently with
    the rest (``engine.response.response_function_set``).  Returns one signal per entry, each identical to what
    ``qspectroscopy`` returns for it.
    """
    spectroscopies = list(spectroscopies)
    if not spectroscopies:
        return []
    _validate_system(system)
    if type(system) is ChromophoreSystem and system.mode_dict is None:
        warnings.warn("Pseudomodes not specified. Executing the dynamics with system as an ExcitonicSystem.")
        return qspectroscopy_set(system.extract_ExcitonicSystem(), spectroscopies, shots, verbose, device_output,
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
    signals = response_function_set(program, np.asarray(system.dipole_moments, dtype=float),
                                    [(sp.side_seq, sp.direction_seq) for sp in spectroscopies], counts,
                                    device_output=device_output)
    ending_sentence(verbose)
    return signals


def qspectroscopy_ensemble(systems,
                           spectroscopies,
                           shots: int = 1000,
                           verbose: bool = True,
                           device_output: bool = False,
                           average: bool = False,
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
    ensemble mean, ``[*grid]``)."""
    from ...engine.program import StepProgram, SystemBatch
    systems, spectroscopies = list(systems), list(spectroscopies)
    if not systems or not spectroscopies:
        return []
    for s_ in systems:
        _validate_system(s_)
        if type(s_) is not type(systems[0]):
            raise TypeError('all members of an ensemble must be of the same class')
    if type(systems[0]) is ChromophoreSystem and systems[0].mode_dict is None:
        warnings.warn("Pseudomodes not specified. Executing the dynamics with system as an ExcitonicSystem.")
        return qspectroscopy_ensemble([s_.extract_ExcitonicSystem() for s_ in systems], spectroscopies, shots, verbose,
                                      device_output, average, **qevolve_kwds)
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
    program = StepProgram(SystemBatch.from_systems(systems), dt, trotter_number, qevolve_kwds.get('Trotter_order', 1), rates,
                          dt_trotter=dt_trotter)
    signals = response_function_set(program, mu, [(sp.side_seq, sp.direction_seq) for sp in spectroscopies], counts,
                                    device_output=device_output)
    if B == 1:                                          # a one-member ensemble still carries the ensemble axis
        signals = [sig[None] for sig in signals]
    if average:
        signals = [sig.mean(0) for sig in signals]
    ending_sentence(verbose)
    return signals
