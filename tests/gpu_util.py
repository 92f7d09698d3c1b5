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


# ---- host restatement of the in-kernel counter-based noise (csrc/common.cuh: noise_key64 / noise_seed /
# noise_next / noise_u24).  Integer work, so the comparison with the kernels is bit-exact.
def noise_key64(seed, clip):
    m = (1 << 64) - 1
    z = (seed + 0x9E3779B97F4A7C15 * (clip + 1)) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


def hop_noise(k64, hop):
    """(512,) uniforms of one hop: lane l owns samples 32*i + l, a 24-bit multiplicative stream seeded per (hop, lane)
    from the clip's 64-bit key (low word additive before the first hash round, high word between the rounds)."""
    key, key2 = k64 & 0xffffffff, k64 >> 32
    lane = np.arange(32, dtype=np.uint64)
    x = ((hop * 32 + lane) * 0x9E3779B1 + key) & 0xffffffff
    x ^= x >> 16; x = (x * 0x7feb352d) & 0xffffffff
    x = (x + key2) & 0xffffffff
    x ^= x >> 15; x = (x * 0x846ca68b) & 0xffffffff
    x ^= x >> 16
    x = (x & 0xffffff00) | 0x100        # 24-bit stream in the top bits, odd; low byte clear (exact int->float)
    out = np.zeros(512, np.float32)
    for i in range(16):
        x = (x * 747796405) & 0xffffffff
        u24 = (x >> 8) ^ 0x800000
        out[32 * i + np.arange(32)] = u24.astype(np.float32) * np.float32(2.0 ** -24)
    return out


def in_kernel_noise(seed, B, F):
    """(B, F*512) uniforms the kernels draw for `seed` when no noise tensor is injected."""
    return np.stack([np.concatenate([hop_noise(noise_key64(seed, b), h) for h in range(F)]) for b in range(B)])
