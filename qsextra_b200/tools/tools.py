"""Small helpers shared by the hot path (mirror of the reference's ``qsextra/tools/tools.py``).

``gray_code_list``/``to_gray`` fix the pseudomode level -> qubit-register encoding used by the
collision-model step (tools.py:77-118); ``unit_converter`` (tools.py:139-190) and
``spectral_function`` (tools.py:27-75) are what the spectroscopy notebooks call around the path.
``kron`` works on dense ``numpy`` arrays here (the reference's works on QuTiP objects, tools.py:15-25).
"""
from __future__ import annotations

import os
import shutil
import uuid

import numpy as np

H_BAR_EV_FS = 0.658212  # [eV * fs], tools.py:168


def if_scalar_to_list(a):
    """Wrap a scalar in a one-item list; leave anything else untouched (tools.py:7-13)."""
    return [a] if np.isscalar(a) else a


def kron(*args):
    """Kronecker product of an ordered sequence of dense operators."""
    out = np.asarray(args[0])
    for nxt in args[1:]:
        out = np.kron(out, np.asarray(nxt))
    return out


def spectral_function(central_frequencies, height, width, frequency_range=None, negative_frequencies=True):
    """Sum of Lorentzians ``h w^2 / ((f - f0)^2 + w^2)`` (tools.py:27-75).

    Returns ``(frequencies, values)``; with ``negative_frequencies`` the positive half is
    mirrored onto negative frequencies.
    """
    centres = np.atleast_1d(np.asarray(if_scalar_to_list(central_frequencies), dtype=float))
    heights = np.atleast_1d(np.asarray(if_scalar_to_list(height), dtype=float))
    widths = np.atleast_1d(np.asarray(if_scalar_to_list(width), dtype=float))
    if frequency_range is None:
        step = widths.min() / 100
        fr = np.arange(0, centres.max() + 4 * widths.max() + step, step, dtype=float)
    else:
        fr = frequency_range
    sf = np.zeros_like(fr, dtype=float)
    for f0, h, w in zip(centres, heights, widths):
        sf += h * w**2 / ((fr - f0)**2 + w**2)
    if negative_frequencies:
        fr = np.concatenate((-np.flip(fr[1:]), fr))
        sf = np.concatenate((np.flip(sf[1:]), sf))
    return fr, sf


def to_gray(input, binary: bool = False) -> int:
    """Gray code ``n ^ (n >> 1)`` of a decimal (int/str) or, with ``binary``, a bit string (tools.py:77-99)."""
    value = int(input, 2) if binary else int(input)
    return value ^ (value >> 1)


def gray_code_list(n: int) -> list[str]:
    """Gray codes of 0 .. 2**n - 1 as ``n``-character bit strings (tools.py:101-118)."""
    return [format(to_gray(i), '0{}b'.format(n)) for i in range(1 << n)]


def create_checkpoint_folder() -> str:
    """Make ``./<uuid4 hex>/`` and return its name (tools.py:120-130)."""
    folder_name = uuid.uuid4().hex
    try:
        os.mkdir(folder_name)
    except OSError as e:
        print(f"Error creating directory '{folder_name}': {e}")
    return folder_name


def destroy_checkpoint_folder(folder_name: str):
    """Remove a checkpoint folder and its contents (tools.py:132-137)."""
    try:
        shutil.rmtree(folder_name)
    except OSError as e:
        print(f"Error deleting directory '{folder_name}': {e}")


def unit_converter(quantity: float, initial_unit: str, final_unit: str) -> float:
    """Convert between ``eV``, ``eV-1``, ``fs`` and ``fs-1`` with hbar = 0.658212 eV fs (tools.py:139-190)."""
    accepted_units = ["ev", "ev-1", "fs", "fs-1"]
    src, dst = initial_unit.casefold(), final_unit.casefold()
    if src not in accepted_units or dst not in accepted_units:
        raise ValueError(f'Input unit not accepted. Accepted units are: {accepted_units}')
    if src == dst:
        return quantity
    if src[:2] == dst[:2]:                      # a unit and its own inverse
        return 1. / quantity
    src_inverse, dst_inverse = src.endswith("-1"), dst.endswith("-1")
    # energy <-> time: E = hbar / t
    value = quantity * H_BAR_EV_FS if src_inverse else quantity / H_BAR_EV_FS
    return value if (src_inverse != dst_inverse) else 1. / value
