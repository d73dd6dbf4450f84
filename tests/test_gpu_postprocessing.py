"""Device post-processing (qsx_rotating_frame + cuFFT + qsx_spectrum_shift) against the reference's golden spectra and
against the NumPy restatement (oracle/postprocessing.py) on random signals, odd lengths and every padding/slice option.
Tolerance: 1e-12 relative to the largest magnitude (FP64; the device sincos/exp and cuFFT differ from libm/pocketfft in
the last bits)."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from oracle import postprocessing as oracle_pp
from qsextra_b200.spectroscopy import FeynmanDiagram
from qsextra_b200.spectroscopy.postprocessing import fourier_transform, postprocessing, rotating_frame

pytestmark = pytest.mark.gpu
TOL = 1e-12


def close(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.abs(a - b).max() <= TOL * max(1.0, np.abs(b).max())


def test_golden_spectra(engine):
    meta, g = load_golden('postprocessing')
    t = np.array(meta['t'])
    w1, s1 = postprocessing(FeynmanDiagram('a', t), g['lin'], RF_freq=1.505, damping_rate=0.01, pad_extension=3)
    assert np.array_equal(w1, g['omega_lin']) and close(s1, g['spec_lin'])
    w2, s2 = postprocessing(FeynmanDiagram('gsb', [t, [0., 1.], t]), g['sig'], RF_freq=1.505, damping_rate=0.01,
                            pad_extension=3, T2_index=1)
    assert np.array_equal(w2[0], g['omega1']) and np.array_equal(w2[1], g['omega3']) and close(s2, g['spec_2d'])


@pytest.mark.parametrize('n1,n2,n3,pad,idx', [(8, 1, 8, 1, 0), (7, 3, 5, 2, 2), (33, 2, 17, 3, 1), (128, 4, 128, 3, 3),
                                              (2, 1, 2, 1, 0)])
def test_third_order_random(engine, n1, n2, n3, pad, idx):
    rng = np.random.default_rng(n1 * 1000 + n3 * 10 + pad)
    sig = rng.normal(size=(n1, n2, n3)) + 1j * rng.normal(size=(n1, n2, n3))
    delays = [np.arange(n1) * 0.7, np.arange(n2) * 5.0, np.arange(n3) * 1.3]
    spec = FeynmanDiagram('se', delays)
    want_rf = oracle_pp.rotating_frame(sig, delays, 1.4, 0.02, idx)
    assert close(rotating_frame(sig, delays, 1.4, 0.02, idx), want_rf)
    w_want, s_want = oracle_pp.fourier_transform(want_rf, delays, 1.4, idx, pad)
    w_got, s_got = fourier_transform(want_rf, delays, 1.4, idx, pad)
    assert all(np.array_equal(a, b) for a, b in zip(w_got, w_want)) and close(s_got, s_want)
    w_full, s_full = postprocessing(spec, sig, 1.4, 0.02, pad, idx)
    assert all(np.array_equal(a, b) for a, b in zip(w_full, w_want)) and close(s_full, s_want)
    # a signal that is already on the device stays there
    dev_sig = torch.as_tensor(sig, device='cuda')
    w_dev, s_dev = postprocessing(spec, dev_sig, 1.4, 0.02, pad, idx)
    assert isinstance(s_dev, torch.Tensor) and s_dev.is_cuda and np.array_equal(s_dev.cpu().numpy(), s_full)


@pytest.mark.parametrize('n,pad', [(1000, 1), (501, 3), (2, 2)])
def test_linear_random(engine, n, pad):
    rng = np.random.default_rng(n + pad)
    sig = rng.normal(size=n) + 1j * rng.normal(size=n)
    delays = [np.arange(n) * 0.25]
    w_want, s_want = oracle_pp.postprocessing(FeynmanDiagram('a', delays[0]), sig, 1.5, 0.01, pad)
    w_got, s_got = postprocessing(FeynmanDiagram('a', delays[0]), sig, 1.5, 0.01, pad)
    assert np.array_equal(w_got, w_want) and close(s_got, s_want)
    assert close(rotating_frame(sig, delays, 1.5, 0.01), oracle_pp.rotating_frame(sig, delays, 1.5, 0.01))


def test_spectrum_of_a_simulated_response(engine):
    """The signal of the engine goes straight into the transforms: one absorptive peak near the site energy."""
    from helpers import build_system
    from qsextra_b200.spectroscopy import qspectroscopy
    system = build_system(dict(energies=[1.55, 1.46], couplings=0.01, dipole_moments=[1., 1.]))
    t = np.arange(0, 60., 1.0)
    sig = qspectroscopy(system, FeynmanDiagram('a', t), shots=0, verbose=False, dt=1.0, coll_rates=0.02)
    omega, spec = postprocessing(FeynmanDiagram('a', t), sig, RF_freq=1.5, damping_rate=0.05, pad_extension=4)
    w_want, s_want = oracle_pp.postprocessing(FeynmanDiagram('a', t), sig, RF_freq=1.5, damping_rate=0.05, pad_extension=4)
    assert close(spec, s_want)
    peak = omega[np.argmax(np.abs(spec))]
    assert 1.40 < peak < 1.62


def test_c_abi_argument_errors(engine):
    from qsextra_b200.engine import cabi
    sig = torch.zeros((4, 2, 4), dtype=torch.complex128, device='cuda')
    t = torch.arange(4, dtype=torch.float64, device='cuda')
    with pytest.raises(cabi.QsxError, match='T2_index'):
        engine.rotating_frame(sig, t, t, 2, 1, 0.0, 0.0)
    with pytest.raises(cabi.QsxError, match='pad_extension'):
        engine.rotating_frame(sig, t, t, 0, 0, 0.0, 0.0)
    assert engine.rotating_frame(sig[:0].contiguous(), t[:0], t, 0, 2, 0.0, 0.0).shape == (0, 8)


def test_signal_stays_on_the_device_from_simulation_to_spectrum(engine):
    from helpers import build_system
    from qsextra_b200.spectroscopy import qspectroscopy
    system = build_system(dict(energies=[1.55, 1.46], couplings=0.01, dipole_moments=[1., 0.8]))
    t1, t2, t3 = np.arange(0, 12., 1.0), [0., 3., 3., 1.], np.arange(0, 9., 1.0)
    diagram = FeynmanDiagram('esa', [t1, t2, t3])
    host = qspectroscopy(system, diagram, shots=0, verbose=False, dt=1.0, coll_rates=0.02)
    dev = qspectroscopy(system, diagram, shots=0, verbose=False, dt=1.0, coll_rates=0.02, device_output=True)
    assert isinstance(dev, torch.Tensor) and dev.is_cuda and tuple(dev.shape) == host.shape
    assert np.array_equal(dev.cpu().numpy(), host)
    omega, spec = postprocessing(diagram, dev, RF_freq=1.5, damping_rate=0.03, pad_extension=2, T2_index=2)
    w_want, s_want = oracle_pp.postprocessing(diagram, host, RF_freq=1.5, damping_rate=0.03, pad_extension=2, T2_index=2)
    assert spec.is_cuda and close(spec.cpu().numpy(), s_want)


def test_all_t2_slices_at_once(engine):
    from qsextra_b200.spectroscopy.postprocessing import postprocessing_all_T2
    rng = np.random.default_rng(4)
    n1, n2, n3, pad = 12, 5, 9, 3
    sig = rng.normal(size=(n1, n2, n3)) + 1j * rng.normal(size=(n1, n2, n3))
    delays = [np.arange(n1) * 0.7, np.arange(n2) * 5.0, np.arange(n3) * 1.3]
    spec = FeynmanDiagram('gsb', delays)
    omega, spectra = postprocessing_all_T2(spec, sig, 1.4, 0.02, pad)
    assert spectra.shape == (n2, n1 * pad, n3 * pad)
    for k in range(n2):
        w, one = postprocessing(spec, sig, 1.4, 0.02, pad, k)
        assert np.array_equal(spectra[k], one) and all(np.array_equal(a, b) for a, b in zip(omega, w))
    dev_spectra = postprocessing_all_T2(spec, torch.as_tensor(sig, device='cuda'), 1.4, 0.02, pad)[1]
    assert dev_spectra.is_cuda and np.array_equal(dev_spectra.cpu().numpy(), spectra)
    with pytest.raises(ValueError):
        postprocessing_all_T2(FeynmanDiagram('a', delays[0]), sig[:, 0, 0])
