"""Statistical sanity of the in-kernel uniform noise stream, checked on its host restatement
(tests/gpu_util.py; the GPU tests prove that restatement bit-exact against the kernels)."""
import numpy as np

from tests.gpu_util import in_kernel_noise


def test_uniform_moments_and_histogram():
    U = in_kernel_noise(1234, 2, 128).ravel()                 # 131072 draws
    assert U.min() >= 0.0 and U.max() < 1.0
    assert abs(U.mean() - 0.5) < 3e-3
    assert abs(U.var() - 1.0 / 12.0) < 1e-3
    hist, _ = np.histogram(U, bins=64, range=(0.0, 1.0))
    expected = U.size / 64
    chi2 = ((hist - expected) ** 2 / expected).sum()
    assert chi2 < 120.0, chi2                                  # 63 dof: mean 63, p(chi2 > 120) ~ 1e-5


def test_no_serial_correlation_along_time_or_along_a_lane_stream():
    U = in_kernel_noise(99, 1, 256)[0].astype(np.float64) - 0.5
    n = U.size
    for lag in (1, 2, 31, 32, 33, 64, 512, 1024):             # 32 = consecutive draws of one lane's LCG stream
        rho = float((U[:-lag] * U[lag:]).sum() / (U * U).sum())
        assert abs(rho) < 5.0 / np.sqrt(n), (lag, rho)


def test_flat_spectrum():
    U = in_kernel_noise(7, 1, 512)[0].astype(np.float64) * 2.0 - 1.0
    frames = U.reshape(-1, 1024) * np.hanning(1024)
    psd = (np.abs(np.fft.rfft(frames, axis=1)) ** 2).mean(axis=0)[4:-4]
    bands = psd.reshape(-1, 101).mean(axis=1) if psd.size % 101 == 0 else psd[:500].reshape(5, 100).mean(axis=1)
    assert bands.max() / bands.min() < 1.25, bands


def test_streams_differ_between_clips_and_seeds():
    a = in_kernel_noise(5, 2, 4)
    assert not np.array_equal(a[0], a[1])
    assert not np.array_equal(a, in_kernel_noise(6, 2, 4))
    assert np.array_equal(a, in_kernel_noise(5, 2, 4))
