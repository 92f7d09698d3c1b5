"""Deterministic synthetic inputs for the synthesizer path.

Used by the tests, the golden-fixture generator and bench.py so that all three see the same
tensors for a given (B, F, seed).  numpy `default_rng` (PCG64) streams are stable across
platforms and numpy versions.
"""
import numpy as np

SR = 44100
HOP = 512


def make_f0(B, F, rng, zero_f0_fraction=0.0, f0_min=65.0, f0_max=800.0, sr=SR, hop=HOP):
    """(B,F) fp32 contour: per-clip log-uniform base in [f0_min,f0_max], slow +-0.5 octave
    glide, 5.5 Hz +-50 cent vibrato, clipped to the config range (configs/*.yaml:3-4);
    optionally a fraction of frames forced to 0 (unvoiced, vocoder.py:460)."""
    t = np.arange(F) * (hop / sr)
    base = np.exp(rng.uniform(np.log(f0_min), np.log(f0_max), size=(B, 1)))
    glide_T = rng.uniform(2.0, 6.0, size=(B, 1))
    glide_ph = rng.uniform(0, 2 * np.pi, size=(B, 1))
    vib_ph = rng.uniform(0, 2 * np.pi, size=(B, 1))
    octave = 0.5 * np.sin(2 * np.pi * t[None, :] / glide_T + glide_ph) \
        + (50.0 / 1200.0) * np.sin(2 * np.pi * 5.5 * t[None, :] + vib_ph)
    f0 = np.clip(base * 2.0 ** octave, f0_min, f0_max).astype(np.float32)
    if zero_f0_fraction > 0:
        # unvoiced runs of 1..4 frames
        mask = rng.random((B, F)) < zero_f0_fraction / 2.5
        for s in range(1, min(4, F)):
            mask[:, s:] |= mask[:, :-s] & (rng.random((B, F - s)) < 0.5)
        f0 = np.where(mask, np.float32(0), f0)
    return f0


def make_inputs(B, F, sum_k, seed=1234, zero_f0_fraction=0.0, hop=HOP, ctrl_std=0.57, noise=True):
    """dict(f0_frames (B,F) f32, ctrl (B,F,sum_k) f32 ~ N(0, ctrl_std^2) -- what a random-init
    Unit2Control emits --, U (B,F*hop) f32 in [0,1))."""
    rng = np.random.default_rng(seed)
    f0 = make_f0(B, F, rng, zero_f0_fraction, hop=hop)
    ctrl = (ctrl_std * rng.standard_normal((B, F, sum_k), dtype=np.float32)).astype(np.float32)
    out = dict(f0_frames=f0, ctrl=ctrl)
    if noise:
        out['U'] = rng.random((B, F * hop), dtype=np.float32)
    return out


def synthetic_state_dict(template, seed=0):
    """Deterministic stand-in for a trained checkpoint: for every floating entry of `template` (a state_dict, only
    names / shapes / dtypes are used) a tensor drawn from a numpy PCG64 stream keyed by (seed, crc32(name)), scaled
    like an initialised network (matrices ~ N(0, 1/fan_in), biases ~ 0.1 N(0,1), norm gains ~ 1 + 0.1 N(0,1),
    weight-norm gains positive, FAVOR+ projection ~ N(0,1)).  Integer entries and the synthesis `window` are kept.
    The reference module (tests/golden/make_golden_control.py) and the drop-in (tests) both `load_state_dict` the
    result with strict=True, so the same weights run through both without shipping megabytes of parameters."""
    import zlib

    import torch
    out = {}
    for name, t in template.items():
        if not torch.is_floating_point(t) or t.dim() == 0 or name == 'window':
            out[name] = t.clone()
            continue
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        x = rng.standard_normal(tuple(t.shape))
        if name.endswith('projection_matrix'):
            pass
        elif name.endswith('weight_g'):
            x = 0.5 + 0.2 * np.abs(x)
        elif t.dim() >= 2:
            x = x / np.sqrt(max(1, t.numel() // t.shape[0]))
        elif name.endswith('bias'):
            x = 0.1 * x
        else:
            x = 1.0 + 0.1 * x
        out[name] = torch.from_numpy(x).to(t.dtype)
    return out
