"""Specification objects for the response-function path: ``Spectroscopy`` and ``FeynmanDiagram``.

Mirror of the reference's ``qsextra/spectroscopy/spectroscopy.py`` (Spectroscopy :7-104,
FeynmanDiagram :106-197): ``delay_time`` is normalised to a list of lists, and a diagram label
maps to the side / direction sequences of the light-matter interactions (:138-150):

    'a' | 'abs' -> 'kk'   / 'io'        'gsb' -> 'bbkk' / 'ioio'
    'se'        -> 'bkbk' / 'iioo'      'esa' -> 'bkkk' / 'iiio'
"""
from __future__ import annotations

import numpy as np


def _normalise_delays(delays) -> list[list]:
    """float -> [[T]]; flat list / 1-D array -> [[...]]; list of arrays/lists/scalars -> list of lists
    (spectroscopy.py:23-58)."""
    if np.isscalar(delays):
        return [[delays]]
    if isinstance(delays, np.ndarray):
        return [delays.tolist()]
    if not isinstance(delays, list):
        raise TypeError('delay_time is a {}. However, a list, numpy.ndarray or a float are expected.'.format(type(delays)))
    if all(np.isscalar(d) for d in delays):
        return [delays]
    out = []
    for d in delays:
        if np.isscalar(d):
            out.append([d])
        elif isinstance(d, np.ndarray):
            out.append(d.tolist())
        elif isinstance(d, list):
            out.append(d)
        else:       # the reference silently re-appends the previous entry here; refuse instead
            raise TypeError('delay_time entries must be floats, lists or numpy.ndarray, not {}.'.format(type(d)))
    return out


class Spectroscopy:
    """Delay-time grid plus the side ('k'/'b') and direction ('i'/'o') sequences of the interactions."""

    def __init__(self, delay_time, label: str = ""):
        self.label = label
        self.delay_time = delay_time
        self.side_seq = ''
        self.direction_seq = ''

    @property
    def label(self) -> str:
        return self._label

    @label.setter
    def label(self, label):
        self._label = label

    @property
    def delay_time(self) -> list[list]:
        return self._delay_time

    @delay_time.setter
    def delay_time(self, delays):
        self._delay_time = _normalise_delays(delays)

    @property
    def side_seq(self) -> str:
        return self._side_seq

    @side_seq.setter
    def side_seq(self, side_sequence):
        self._side_seq = side_sequence

    @property
    def direction_seq(self) -> str:
        return self._direction_seq

    @direction_seq.setter
    def direction_seq(self, direction_sequence):
        self._direction_seq = direction_sequence

    def todict(self) -> dict:
        """Object data as a dict (attribute names without the leading underscore, plus 'class')."""
        out = {'class': type(self)}
        for key, value in self.__dict__.items():
            out[key[1:] if key[0] == '_' else key] = value.tolist() if isinstance(value, np.ndarray) else value
        return out


class FeynmanDiagram(Spectroscopy):
    """One double-sided Feynman diagram: linear absorption ('a'/'abs') or a rephasing third-order
    contribution ('gsb', 'se', 'esa').  ``delay_time``: one list for linear, ``[T1, T2, T3]`` for
    third order."""

    _linear_list = ["a", "abs"]
    _thirdorder_list = ["gsb", "se", "esa"]
    _sequences = {'a': ('kk', 'io'), 'abs': ('kk', 'io'),
                  'gsb': ('bbkk', 'ioio'), 'se': ('bkbk', 'iioo'), 'esa': ('bkkk', 'iiio')}

    def __init__(self, FD_label: str, delay_time):
        self.set(FD_label, delay_time)

    def set(self, FD_label: str, delay_time):
        """(Re)define the diagram; ``ValueError`` for an unknown label or a wrong number of delay
        lists (spectroscopy.py:153-197)."""
        if not isinstance(FD_label, str):
            raise TypeError('FD_label must be a str.')
        FD_label = FD_label.casefold()
        if FD_label not in self._sequences:
            raise ValueError('Unsupported Feynman diagram.')
        self.label = FD_label
        self.delay_time = delay_time
        expected = 1 if FD_label in self._linear_list else 3
        if len(self.delay_time) != expected:
            raise ValueError('delay_time contains {} items. However, {} expected.'.format(
                len(self.delay_time), '1 is' if expected == 1 else '3 are'))
        self.side_seq, self.direction_seq = self._sequences[FD_label]
