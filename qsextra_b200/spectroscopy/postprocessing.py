"""Rotating frame + Fourier transform of a simulated response function on the GPU (mirror of the reference's
``qsextra/spectroscopy/postprocessing.py:9-122``; same function names, arguments and return values).

Linear:  R(T) e^{+i w_RF T} e^{-g T}, zero-pad, ``ifft``, ``ifftshift``                  (postprocessing.py:9-19, :50-62)
2-D:     R(T1, T2[idx], T3) e^{-i w_RF (T1 - T3)} e^{-g (T1 + T3)}, zero-pad, ``fft`` along T1, ``ifft`` along T3,
         ``fftshift`` of both axes, prefactor len(padded) / len(original)**2             (postprocessing.py:21-37, :64-86)

The frame/damping/padding pass and the shift/prefactor pass are the library's own kernels (``qsx_rotating_frame``,
``qsx_spectrum_shift``); the transforms between them are cuFFT Z2Z through ``torch.fft``.  A NumPy signal is uploaded
and a NumPy spectrum returned; a CUDA tensor (e.g. the signal ``engine.response`` produced) stays on the device and a
CUDA tensor is returned.  The frequency axes are small host arrays in both cases.
"""
from __future__ import annotations

import numpy as np

from .spectroscopy import Spectroscopy


def _fft_axis(n: int, pad: int, dt: float, rf_freq: float) -> np.ndarray:
    """Frequency axis of one transformed delay (postprocessing.py:56, :74-75); a small host array."""
    return np.fft.fftshift(2 * np.pi * np.fft.fftfreq(n * pad, dt)) + rf_freq


def _validate(delay_time, t2_index):
    if len(delay_time) == 3 and t2_index >= len(delay_time[1]):
        raise ValueError('T2_index exceed length of T2')
    return len(delay_time) in (1, 3)


class _Device:
    """Uploads the signal if it lives on the host and gives results back in the caller's form."""

    def __init__(self, signal):
        import torch
        from ..engine.runtime import get_engine
        self.engine = get_engine()
        self.torch = torch
        self.on_host = not isinstance(signal, torch.Tensor)
        if self.on_host:
            signal = self.engine.to_device(np.asarray(signal, dtype=np.complex128))
        self.signal = signal.to(torch.complex128).contiguous()

    def axis(self, t):
        return self.engine.to_device(np.asarray(t, dtype=np.float64))

    def give(self, tensor):
        return tensor.cpu().numpy() if self.on_host else tensor


def _frame(dev: _Device, signal, delay_time, rf_freq, damping_rate, t2_index, pad):
    """Device tensor of the (padded) rotating-frame signal."""
    if len(delay_time) == 1:
        return dev.engine.rotating_frame(signal, dev.axis(delay_time[0]), None, 0, pad, rf_freq, damping_rate)
    if len(delay_time) == 3:
        t1, t2, t3 = delay_time
        if signal.dim() == 2:                                 # an already sliced (T1, T3) array
            signal, t2_index = signal.unsqueeze(1).contiguous(), 0
        return dev.engine.rotating_frame(signal, dev.axis(t1), dev.axis(t3), t2_index, pad, rf_freq, damping_rate)


def _transform(dev: _Device, padded, delay_time, rf_freq, pad):
    fft = dev.torch.fft
    if len(delay_time) == 1:
        t = delay_time[0]
        n = len(t)
        omega = _fft_axis(n, pad, t[1] - t[0], rf_freq)
        m = n * pad
        return omega, dev.engine.spectrum_shift(fft.ifft(padded), m // 2, 0, 1.0)                 # ifftshift
    t1, _, t3 = delay_time
    n1, n3 = len(t1), len(t3)
    omega1 = _fft_axis(n1, pad, t1[1] - t1[0], rf_freq)
    omega3 = _fft_axis(n3, pad, t3[1] - t3[0], rf_freq)
    omega = list(np.meshgrid(omega1, omega3, indexing='ij'))
    m1, m3 = n1 * pad, n3 * pad
    both = fft.ifft(fft.fft(padded, dim=0), dim=1).contiguous()
    return omega, dev.engine.spectrum_shift(both, m1 - m1 // 2, m3 - m3 // 2, m1 / n1 ** 2)       # fftshift x 2


def rotating_frame(response_function, delay_time, RF_freq=0, damping_rate=0, T2_index=0):
    if not _validate(delay_time, T2_index):
        return None
    dev = _Device(response_function)
    return dev.give(_frame(dev, dev.signal, delay_time, RF_freq, damping_rate, T2_index, 1))


def fourier_transform(response_function, delay_time, RF_freq=0, T2_index=0, pad_extension=1):
    """``response_function`` is the rotating-frame signal: [n] (linear) or the (T1, T3) slice [n1, n3]."""
    if not _validate(delay_time, T2_index):
        return None
    dev = _Device(response_function)
    padded = _frame(dev, dev.signal, delay_time, 0.0, 0.0, T2_index, pad_extension)
    omega, spectrum = _transform(dev, padded, delay_time, RF_freq, pad_extension)
    return omega, dev.give(spectrum)


def postprocessing(spectroscopy: Spectroscopy, signal, RF_freq: float = 0, damping_rate: float = 0,
                   pad_extension: int = 1, T2_index: int = 0):
    """(omega, spectrum) of ``signal`` simulated for ``spectroscopy`` (postprocessing.py:99-122); frame, damping and
    padding are one pass here."""
    delay_time = spectroscopy.delay_time
    if not _validate(delay_time, T2_index):
        return None
    dev = _Device(signal)
    padded = _frame(dev, dev.signal, delay_time, RF_freq, damping_rate, T2_index, pad_extension)
    omega, spectrum = _transform(dev, padded, delay_time, RF_freq, pad_extension)
    return omega, dev.give(spectrum)


def postprocessing_all_T2(spectroscopy: Spectroscopy, signal, RF_freq: float = 0, damping_rate: float = 0,
                          pad_extension: int = 1):
    """Every T2 slice of a third-order signal at once (an addition to the reference API, which takes one ``T2_index``
    per call): (omega, spectra [n2, n1 * pad, n3 * pad]); ``spectra[k]`` equals ``postprocessing(..., T2_index=k)[1]``.
    The frame/padding and shift passes run per slice into one buffer, the transforms are two batched cuFFT calls."""
    delay_time = spectroscopy.delay_time
    if len(delay_time) != 3:
        raise ValueError('postprocessing_all_T2 needs a third-order signal (three delay lists)')
    dev = _Device(signal)
    t1, t2, t3 = delay_time
    n1, n2, n3 = len(t1), len(t2), len(t3)
    m1, m3 = n1 * pad_extension, n3 * pad_extension
    torch = dev.torch
    ax1, ax3 = dev.axis(t1), dev.axis(t3)
    padded = torch.empty((n2, m1, m3), dtype=torch.complex128, device=dev.engine.device)
    for k in range(n2):
        dev.engine.rotating_frame(dev.signal, ax1, ax3, k, pad_extension, RF_freq, damping_rate, out=padded[k])
    both = torch.fft.ifft(torch.fft.fft(padded, dim=1), dim=2).contiguous()
    spectra = padded                                            # the padded buffer is free again: reuse it for the result
    for k in range(n2):
        dev.engine.spectrum_shift(both[k], m1 - m1 // 2, m3 - m3 // 2, m1 / n1 ** 2, out=spectra[k])
    omega1 = _fft_axis(n1, pad_extension, t1[1] - t1[0], RF_freq)
    omega3 = _fft_axis(n3, pad_extension, t3[1] - t3[0], RF_freq)
    return list(np.meshgrid(omega1, omega3, indexing='ij')), dev.give(spectra)
