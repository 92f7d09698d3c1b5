import numpy as np
import pytest

try:
    import torch
    HAS_CUDA = torch.cuda.is_available()
except Exception:        # pragma: no cover
    torch = None
    HAS_CUDA = False

SPLITS = {'combsubfast': (513, 513, 513), 'combsub': (256, 512, 256), 'sins': (128, 256, 256)}


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def ctrl_views(ctrl_np, model):
    """The strided views a real Unit2Control emits: torch.split of one (B,F,sumK) tensor."""
    packed = dev(ctrl_np)
    return torch.split(packed, list(SPLITS[model]), dim=-1)


def snr_db(ref, out):
    ref = np.asarray(ref, np.float64)
    out = np.asarray(out, np.float64)
    den = np.sum((out - ref) ** 2)
    return float('inf') if den == 0 else 10 * np.log10(np.sum(ref ** 2) / den)


# north-star tolerances (BASELINE.json): max-abs <= 1e-4 of full scale, SNR >= 60 dB
MAX_ABS_TOL = 1e-4
SNR_TOL_DB = 60.0


def assert_waveform(out, ref, max_abs=MAX_ABS_TOL, snr=SNR_TOL_DB, what=''):
    out = np.asarray(out, np.float64)
    ref = np.asarray(ref, np.float64)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    assert np.all(np.isfinite(out)), f'{what}: non-finite output'
    err = np.abs(out - ref).max()
    s = snr_db(ref, out)
    assert err <= max_abs, f'{what}: max-abs {err:.3e} > {max_abs:.1e} (snr {s:.1f} dB)'
    assert s >= snr, f'{what}: snr {s:.1f} dB < {snr} (max-abs {err:.3e})'
    return err, s
