"""TEST INFRASTRUCTURE -- CPU restatement of the reference's post-processing, never used by the product path.

NumPy/SciPy restatement of ``qsextra/spectroscopy/postprocessing.py:9-122`` (rotating frame, zero padding, Fourier
transforms); pinned by ``tests/golden/postprocessing.npz``, which the unmodified reference produced
(``tests/golden/make_golden.py``).  Only ``tests/`` may import it: it is the checker of the device implementation in
``qsextra_b200/spectroscopy/postprocessing.py``.

Linear:  R(T) e^{+i w_RF T} e^{-g T}, zero-pad, ``ifft``              (postprocessing.py:9-19, :50-62)
2-D:     R(T1, T2[idx], T3) e^{-i w_RF (T1 - T3)} e^{-g (T1 + T3)}, zero-pad, ``fft`` along T1, ``ifft`` along T3,
         prefactor len(padded) / len(original)**2                     (postprocessing.py:21-37, :64-86)
"""
from __future__ import annotations

import numpy as np
from scipy.fft import fft, fftfreq, fftshift, ifft, ifftshift



def rotating_frame(response_function, delay_time, RF_freq=0, damping_rate=0, T2_index=0):
    if len(delay_time) == 1:
        t = np.array(delay_time[0])
        return response_function * np.exp(1.j * RF_freq * t) * np.exp(-damping_rate * t)
    if len(delay_time) == 3:
        t1, t2, t3 = (np.array(t) for t in delay_time)
        if T2_index >= len(t2):
            raise ValueError('T2_index exceed length of T2')
        t1, t3 = np.meshgrid(t1, t3, indexing='ij')
        return response_function[:, T2_index, :] * np.exp(-1.j * RF_freq * (t1 - t3)) * np.exp(-damping_rate * (t1 + t3))


def fourier_transform(response_function, delay_time, RF_freq=0, T2_index=0, pad_extension=1):
    if len(delay_time) == 1:
        t = delay_time[0]
        n = len(t)
        omega = fftshift(2 * np.pi * fftfreq(n * pad_extension, t[1] - t[0])) + RF_freq
        padded = np.pad(response_function, (0, n * (pad_extension - 1)), 'constant')
        return omega, ifftshift(ifft(padded))
    if len(delay_time) == 3:
        t1, t2, t3 = delay_time
        if T2_index >= len(t2):
            raise ValueError('T2_index exceed length of T2')
        n1, n3 = len(t1), len(t3)
        omega1 = fftshift(2 * np.pi * fftfreq(n1 * pad_extension, t1[1] - t1[0])) + RF_freq
        omega3 = fftshift(2 * np.pi * fftfreq(n3 * pad_extension, t3[1] - t3[0])) + RF_freq
        omega = list(np.meshgrid(omega1, omega3, indexing='ij'))
        padded = np.pad(response_function, ((0, n1 * (pad_extension - 1)), (0, n3 * (pad_extension - 1))), 'constant')
        along_t1 = fftshift(fft(padded, axis=0), axes=0)
        spectrum = len(padded) / len(response_function) ** 2 * fftshift(ifft(along_t1, axis=1), axes=1)
        return omega, spectrum


def postprocessing(spectroscopy, signal: np.ndarray, RF_freq: float = 0, damping_rate: float = 0,
                   pad_extension: int = 1, T2_index: int = 0):
    """(omega, spectrum) of ``signal`` simulated for ``spectroscopy`` (postprocessing.py:99-122)."""
    delay_time = spectroscopy.delay_time
    rf = rotating_frame(signal, delay_time, RF_freq, damping_rate, T2_index)
    return fourier_transform(rf, delay_time, RF_freq, T2_index, pad_extension)
