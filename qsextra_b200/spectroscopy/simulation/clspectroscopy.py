"""``clspectroscopy`` -- Liouville-space comparison path of the reference
(``qsextra/spectroscopy/simulation/clspectroscopy.py:18-181``): the response function with the master-equation
dynamics of ``clevolve`` between the light-matter interactions.

The reference loops over site pathways and grid points and calls QuTiP for every delay (clspectroscopy.py:101-130).
Here the same sum is the engine's batched tree (``engine/response.py``): a collective dipole operator
sum_s mu_s O_s per interaction (``qsx_apply_dipole``), all delays of a level propagated together by
``qsx_lindblad_evolve`` with prefix sharing, and the emission Tr[(sum_s mu_s X_s) sigma] read on the device.
"""
from __future__ import annotations

import warnings

import numpy as np

from ...engine.lindblad import LindbladProgram
from ...engine.response import response_function
from ...system import ChromophoreSystem, ExcitonicSystem
from ...tools.oracle import ending_sentence
from ...tools.tools import if_scalar_to_list
from ..spectroscopy import Spectroscopy


def clspectroscopy(system: ChromophoreSystem | ExcitonicSystem,
                   spectroscopy: Spectroscopy,
                   verbose: bool = True,
                   **clevolve_kwds,
                   ) -> np.ndarray:
    """Response function of ``spectroscopy`` on ``system`` in the Lindblad limit (same signature as the reference).

    ``clevolve_kwds``: ``rates`` (see ``clevolve``).  Returns a complex array with one axis per delay list.
    """
    if not type(system) is ChromophoreSystem and not type(system) is ExcitonicSystem:
        raise TypeError('system must be an instance of class ChromophoreSystem or ExcitonicSystem.')
    if not system._validity:
        raise Exception('Make sure all the system parameters are specified and consistent.')
    if type(system) is ChromophoreSystem and system.mode_dict is None:
        warnings.warn("Pseudomodes not specified. Executing the dynamics with system as an ExcitonicSystem.")
        return clspectroscopy(system.extract_ExcitonicSystem(), spectroscopy, verbose, **clevolve_kwds)
    if spectroscopy.side_seq == '' or spectroscopy.direction_seq == '':
        raise ValueError('Spectroscopy object is not correctly defined. Spectroscopy.side_seq or Spectroscopy.direction_seq are missing.')
    unknown = set(clevolve_kwds) - {'rates', 'measure_populations', 'state_overwrite', 'verbose'}
    if unknown:
        raise TypeError("clevolve() got an unexpected keyword argument '{}'".format(sorted(unknown)[0]))
    rates = clevolve_kwds.get('rates')
    if type(system) is ExcitonicSystem:
        if not np.isscalar(rates) and rates is not None:
            raise TypeError('For an ExcitonicSystem, rates must be a scalar.')
    elif rates is not None:
        rates = if_scalar_to_list(rates)
        if len(rates) != len(system.mode_dict['omega_mode']):
            raise ValueError('The number of dephasing rates is different from the number of pseudomodes per chromophore.')
    program = LindbladProgram.from_system(system, rates)
    delays = [np.asarray(t, dtype=float).tolist() for t in spectroscopy.delay_time]
    signal = response_function(program, np.asarray(system.dipole_moments, dtype=float), spectroscopy.side_seq,
                               spectroscopy.direction_seq, delays)
    ending_sentence(verbose)
    return signal
