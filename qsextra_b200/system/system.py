"""Parameter containers of the hot path: ``ExcitonicSystem`` and ``ChromophoreSystem``.

Drop-in mirror of the reference's ``qsextra/system/system.py`` (ExcitonicSystem :15-302,
ChromophoreSystem :304-538): same attribute names, same setter validation and error
types, same ``todict()`` keys.  The reference returns QuTiP ``Qobj`` from the
``get_*`` helpers; QuTiP is not a dependency here, so they return dense
``numpy`` arrays with the reference's QuTiP (big-endian) ordering
``[e_{N-1} ... e_0, modes(N-1) ... modes(0)]`` (system.py:260-276, :488-527).
"""
from __future__ import annotations

import warnings
from typing import Literal

import numpy as np

from ..tools.tools import if_scalar_to_list

_FROZEN_MSG = '{} parameters cannot be changed after a state has been set.'
_STATE_TYPES = ('state', 'delocalized excitation', 'localized excitation', 'ground')

_SZ = np.diag([1.0, -1.0]).astype(complex)
_SM = np.array([[0, 1], [0, 0]], dtype=complex)        # destroy(2) = |0><1|
_I2 = np.eye(2, dtype=complex)


def _kron_chain(*mats):
    out = np.ones((1, 1), dtype=complex)
    for m in mats:
        out = np.kron(out, m)
    return out


def _as_vector(value, what: str) -> np.ndarray:
    """Scalar / list / 1-D array -> 1-D array (system.py:46-53, :79-86)."""
    if np.isscalar(value):
        value = np.array([value])
    elif isinstance(value, list):
        value = np.array(value)
    if not isinstance(value, np.ndarray):
        raise TypeError('{} is in a wrong format ({}).'.format(what, type(value)))
    if value.ndim > 1:
        raise Exception('{} must be a list, a number (for a single chromophore) or a monodimensional array.'.format(what))
    return value


class ExcitonicSystem:
    r"""Network of two-level chromophores, :math:`H^e = -\sum_i \epsilon_i/2\,\sigma^z_i
    + \sum_{i<j} J_{ij}(\sigma^+_i\sigma^-_j + h.c.)` (reference system.py:115-151).

    Parameters
    ----------
    energies : list[float] | float | numpy.ndarray
        Electronic energy gaps.
    couplings : list[list[float]] | float | numpy.ndarray
        Hermitian coupling matrix (diagonal ignored); a float is accepted for a dimer.
    dipole_moments : list[float] | float | numpy.ndarray
        Transition dipole amplitudes; ``None`` means all ones (with a warning).
    """

    def __init__(self, energies=None, couplings=None, dipole_moments=None):
        self._system_size = 0
        self._e_el = None
        self._dipole_moments = None
        self._coupl_el = None
        self._validity = False
        self._state_type = None
        self._state = None
        if energies is not None:
            self.electronics(energies=energies, couplings=couplings, dipole_moments=dipole_moments)

    # ---- guarded attributes -------------------------------------------------------
    def _require_unfrozen(self):
        if self._state_type is not None:
            raise Exception(_FROZEN_MSG.format(type(self)))

    @property
    def e_el(self):
        return self._e_el

    @e_el.setter
    def e_el(self, energies):
        self._require_unfrozen()
        energies = _as_vector(energies, 'energies')
        self._e_el = energies
        self._system_size = energies.size
        self._refresh_validity()

    @property
    def system_size(self):
        return self._system_size

    @system_size.setter
    def system_size(self, size):
        self._require_unfrozen()
        self._system_size = size
        self._refresh_validity()

    @property
    def dipole_moments(self):
        return self._dipole_moments

    @dipole_moments.setter
    def dipole_moments(self, dipole_moments):
        self._require_unfrozen()
        if dipole_moments is None:
            warnings.warn('Equal dipole moments are considered.')
            dipole_moments = [1.] * self.system_size
        self._dipole_moments = _as_vector(dipole_moments, 'dipole_moments')
        self._refresh_validity()

    @property
    def coupl_el(self):
        return self._coupl_el

    @coupl_el.setter
    def coupl_el(self, couplings):
        self._require_unfrozen()
        if couplings is None:
            warnings.warn('No coupling between chromophores is considered.')
            couplings = [[0 for _ in range(self.system_size)] for _ in range(self.system_size)]
        if np.isscalar(couplings) and self.system_size == 2:
            couplings = np.array([[0, couplings], [couplings, 0]])
        if isinstance(couplings, list) and isinstance(couplings[0], list):
            couplings = np.array(couplings)
        elif not isinstance(couplings, np.ndarray):
            raise TypeError('couplings is in a wrong format ({}).'.format(type(couplings)))
        # reference subtracts the diagonal in place (system.py:107); same result, same object
        couplings -= np.diag(np.diag(couplings))
        if not np.array_equal(couplings, couplings.conj().T):
            raise Exception('couplings is not Hermitian, but it must be.')
        self._coupl_el = couplings
        self._refresh_validity()

    @property
    def state_type(self):
        return self._state_type

    @property
    def state(self):
        return self._state

    def _refresh_validity(self):
        """system.py:153-167: valid iff the three containers agree with ``system_size``."""
        try:
            n = self.system_size
            sizes = (self.e_el.size, self.dipole_moments.size, self.coupl_el.shape[0])
        except Exception:
            self._validity = False
            return
        if sizes == (n, n, n):
            self._validity = True
            return
        kind = type(self)
        if sizes[0] != n:
            warnings.warn('The number of elements in energies ({}) is different from the system size (number of chromophores = {}). Make sure they match in order to have a valid {}.'.format(sizes[0], n, kind))
        if sizes[1] != n:
            warnings.warn('The number of elements in dipole_moments ({}) is different from the system size (number of chromophores = {}). Make sure they match in order to have a valid {}.'.format(sizes[1], n, kind))
        if sizes[2] != n:
            warnings.warn('The number of elements in couplings ({}x{}) is different from that expected ({}x{}). Make sure they match in order to have a valid {}.'.format(sizes[2], sizes[2], n, n, kind))
        self._validity = False

    # ---- public API -----------------------------------------------------------------
    def electronics(self, energies, couplings=None, dipole_moments=None):
        """Set all electronic parameters at once (system.py:169-181)."""
        self.e_el = energies
        self.system_size = self.e_el.size
        self.dipole_moments = dipole_moments
        if self.dipole_moments.size != self.system_size:
            raise Exception('The number of elements in dipole_moment ({}) is different from the system size (number of chromophores = {}).'.format(self.dipole_moments.size, self.system_size))
        self.coupl_el = couplings
        if self.coupl_el.shape[0] != self.system_size:
            raise Exception('The number of elements in couplings ({}x{}) is different from that expected ({}x{}).'.format(self.coupl_el.shape[0], self.coupl_el.shape[0], self.system_size, self.system_size))

    def set_state(self,
                  state_type: Literal['state', 'delocalized excitation', 'localized excitation', 'ground'] = 'ground',
                  state=0):
        """Set the (pure) electronic state (system.py:183-241).

        ``'state'``: 2**N amplitudes, standard binary ordering (index bit i = chromophore i);
        ``'delocalized excitation'``: N single-excitation coefficients;
        ``'localized excitation'``: index of the excited chromophore; ``'ground'``: nothing.
        Amplitude lists are normalised (with a warning) when needed.
        """
        if not self._validity:
            raise Exception('Make sure system electronic parameters (energies, dipole moment intensities and couplings) are specified and consistent before setting the state.')
        if state_type not in _STATE_TYPES:
            raise ValueError('state_type must be one of {}'.format(_STATE_TYPES))
        if state_type in ('state', 'delocalized excitation'):
            label = 'a state' if state_type == 'state' else 'a delocalized excitation'
            if not isinstance(state, (list, np.ndarray)):
                raise TypeError('For {}, state must be an list.'.format(label))
            amplitudes = np.array(state, dtype='complex')
            if amplitudes.ndim > 1:
                raise Exception('state must be a monodimensional array.')
            if state_type == 'state' and amplitudes.size != 2**self.system_size:
                raise Exception('The number of elements in state ({}) is different from the Hilbert space dimension ({}).'.format(amplitudes.size, 2**self.system_size))
            if state_type == 'delocalized excitation' and amplitudes.size != self.system_size:
                raise Exception('The number of elements in state ({}) is different from the system size (number of chromophores = {}).'.format(amplitudes.size, self.system_size))
            norm = np.sum(np.abs(amplitudes)**2)
            if norm != 1.:
                amplitudes = amplitudes / np.sqrt(norm)
                warnings.warn('Coefficients have been normalized.')
            state = amplitudes.tolist()
        elif state_type == 'localized excitation':
            if not isinstance(state, int):
                raise TypeError('For a localized excitation, state must be an integer.')
            if state >= self.system_size or state < 0:
                raise ValueError('state must be an int between 0 and {}'.format(self.system_size))
        else:
            state = 0
        self._state_type = state_type
        self._state = state

    def todict(self) -> dict:
        """Object data as a dict: attribute names without the leading underscore, arrays as
        lists, plus a ``'class'`` entry (system.py:243-258)."""
        out = {'class': type(self)}
        for key, value in self.__dict__.items():
            out[key[1:] if key[0] == '_' else key] = value.tolist() if isinstance(value, np.ndarray) else value
        return out

    def _site_operator(self, site: int, op: np.ndarray) -> np.ndarray:
        n = self.system_size
        return _kron_chain(*[_I2] * (n - site - 1), op, *[_I2] * site)

    def get_e_Hamiltonian(self) -> np.ndarray:
        """Dense Frenkel-exciton Hamiltonian, basis order e_{N-1}...e_0 (system.py:260-276)."""
        if not self._validity:
            raise Exception("The system is not valid. Please, reset the parameters.")
        n = self.system_size
        ham = np.zeros((2**n, 2**n), dtype=complex)
        for i in range(n):
            ham += -self.e_el[i] / 2 * self._site_operator(i, _SZ)
            for j in range(i + 1, n):
                hop = self._site_operator(j, _SM) @ self._site_operator(i, _SM.conj().T)
                ham += self.coupl_el[i, j] * (hop + hop.conj().T)
        return ham

    def get_e_state(self) -> np.ndarray:
        """Ket of the electronic state as a dense vector (system.py:280-300)."""
        if not self._validity:
            raise Exception("The system is not valid. Please, reset the parameters.")
        n = self.system_size
        ket = np.zeros(2**n, dtype=complex)
        if self.state_type == 'ground':
            ket[0] = 1.
        elif self.state_type == 'state':
            ket[:] = np.array(self.state)
        elif self.state_type == 'delocalized excitation':
            for site, amplitude in enumerate(self.state):
                ket[1 << site] += amplitude
        elif self.state_type == 'localized excitation':
            ket[1 << self.state] = 1.
        return ket


class ChromophoreSystem(ExcitonicSystem):
    r"""``ExcitonicSystem`` plus identical harmonic pseudomodes on every chromophore
    (reference system.py:304-538).

    Extra parameters: ``excitonic_system`` (copy electronic part and state from it),
    ``frequencies_pseudomode``, ``levels_pseudomode``, ``couplings_ep`` (one entry per
    pseudomode of a chromophore; scalars accepted for a single pseudomode).
    """

    def __init__(self, energies=None, couplings=None, dipole_moments=None,
                 excitonic_system: ExcitonicSystem = None,
                 frequencies_pseudomode=None, levels_pseudomode=None, couplings_ep=None):
        if type(excitonic_system) is ExcitonicSystem:
            super().__init__(energies=excitonic_system.e_el,
                             couplings=excitonic_system.coupl_el,
                             dipole_moments=excitonic_system.dipole_moments)
            try:
                self.set_state(excitonic_system.state_type, excitonic_system.state)
            except Exception:
                pass
        else:
            super().__init__(energies=energies, couplings=couplings, dipole_moments=dipole_moments)
        self._mode_dict = None
        if frequencies_pseudomode is not None:
            self.pseudomodes(frequencies_pseudomode=frequencies_pseudomode,
                             levels_pseudomode=levels_pseudomode,
                             couplings_ep=couplings_ep)

    @property
    def mode_dict(self):
        return self._mode_dict

    @mode_dict.setter
    def mode_dict(self, md):
        self._mode_dict = md

    def pseudomodes(self, frequencies_pseudomode=None, levels_pseudomode=None, couplings_ep=None):
        """Attach pseudomodes; every pseudomode starts in its lowest level (system.py:378-439)."""
        if frequencies_pseudomode is None:
            raise Exception('frequencies_pseudomode is not specified.')
        frequencies_pseudomode = if_scalar_to_list(frequencies_pseudomode)
        if not isinstance(frequencies_pseudomode, list):
            raise TypeError('frequencies_pseudomode is in a wrong format ({}).'.format(type(frequencies_pseudomode)))
        n_modes = len(frequencies_pseudomode)

        if levels_pseudomode is None:
            raise Exception('levels_pseudomode is not specified.')
        levels_pseudomode = if_scalar_to_list(levels_pseudomode)
        if not isinstance(levels_pseudomode, list):
            raise TypeError('levels_pseudomode is in a wrong format ({}).'.format(type(levels_pseudomode)))
        if len(levels_pseudomode) != n_modes:
            raise Exception('The number of elements in levels_pseudomode ({}) is different from the nummer of elements in frequencies_pseudomode ({}).'.format(len(levels_pseudomode), n_modes))
        if min(levels_pseudomode) < 2:
            raise Exception('All the entries in levels_pseudomode must be >= 2.')

        if couplings_ep is None:
            raise Exception('couplings_ep is not specified.')
        couplings_ep = if_scalar_to_list(couplings_ep)
        if not isinstance(couplings_ep, list):
            raise TypeError('couplings_ep is in a wrong format ({}).'.format(type(couplings_ep)))
        if len(couplings_ep) != n_modes:
            raise Exception('The number of elements in couplings_ep ({}) is different from the nummer of elements in frequencies_pseudomode ({}).'.format(len(couplings_ep), n_modes))

        self.mode_dict = {'omega_mode': frequencies_pseudomode,
                          'lvl_mode': levels_pseudomode,
                          'coupl_ep': couplings_ep,
                          'state_mode': [[1. + 0.j] + [0. + 0.j] * (d - 1) for d in levels_pseudomode],
                          }

    def get_state_mode(self, mode_number: int | None = None):
        """Pseudomode kets (system.py:452-473; the reference's indexed branch is broken, this one works)."""
        kets = [np.array(amplitudes, dtype=complex) for amplitudes in self.mode_dict['state_mode']]
        if mode_number is None:
            return kets
        if not isinstance(mode_number, int):
            raise TypeError('mode_number must be int or None.')
        if mode_number >= len(kets):
            raise ValueError('mode_number must be lower than {} or None.'.format(len(kets)))
        return kets[mode_number]

    @staticmethod
    def _destroy(levels: int) -> np.ndarray:
        return np.diag(np.sqrt(np.arange(1, levels)), 1).astype(complex)

    def get_p_Hamiltonian(self) -> np.ndarray:
        """Free Hamiltonian of the W pseudomodes of ONE chromophore (system.py:475-486)."""
        levels = self.mode_dict['lvl_mode']
        eyes = [np.eye(d, dtype=complex) for d in levels]
        dim = int(np.prod(levels))
        ham = np.zeros((dim, dim), dtype=complex)
        for k, omega in enumerate(self.mode_dict['omega_mode']):
            a = self._destroy(levels[k])
            ham += omega * _kron_chain(*eyes[:k], a.conj().T @ a, *eyes[k + 1:])
        return ham

    def get_ep_Hamiltonian(self) -> np.ndarray:
        """Exciton-pseudomode coupling on [e_{N-1}..e_0, modes(N-1)..modes(0)] (system.py:488-508)."""
        n = self.system_size
        levels = self.mode_dict['lvl_mode']
        eyes = [np.eye(d, dtype=complex) for d in levels]
        dim = 2**n * int(np.prod(levels))**n
        ham = np.zeros((dim, dim), dtype=complex)
        for i in range(n):
            for k, g in enumerate(self.mode_dict['coupl_ep']):
                a = self._destroy(levels[k])
                ham += g / 2 * _kron_chain(*[_I2] * (n - i - 1), _I2 - _SZ, *[_I2] * i,
                                           *eyes * (n - i - 1),
                                           *eyes[:k], a.conj().T + a, *eyes[k + 1:],
                                           *eyes * i)
        return ham

    def get_global_Hamiltonian(self) -> np.ndarray:
        """H_e + sum_i H_p(i) + H_ep on the full space (system.py:510-527)."""
        n = self.system_size
        eyes = [np.eye(d, dtype=complex) for d in self.mode_dict['lvl_mode']]
        h_p = self.get_p_Hamiltonian()
        ham = _kron_chain(self.get_e_Hamiltonian(), *eyes * n) + self.get_ep_Hamiltonian()
        for i in range(n):
            ham += _kron_chain(*[_I2] * n, *eyes * (n - i - 1), h_p, *eyes * i)
        return ham

    def extract_ExcitonicSystem(self) -> ExcitonicSystem:
        """Electronic part (and state, when set) as an ``ExcitonicSystem`` (system.py:529-538)."""
        e_sys = ExcitonicSystem(energies=self.e_el, couplings=self.coupl_el, dipole_moments=self.dipole_moments)
        try:
            e_sys.set_state(self.state_type, self.state)
        except Exception:
            warnings.warn('State is not defined.')
        return e_sys
