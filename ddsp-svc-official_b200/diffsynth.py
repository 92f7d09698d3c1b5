"""Differentiable restatement of stage B of the `frequency_filter` synthesizers (Sins, CombSub-old) in stock
torch ops -- the BACKWARD path of the drop-in modules.

The forward of `ddsp_b200.vocoder.Sins / CombSub` always runs the hand-written kernels; when the control tensors
require grad (solver.py:111-113: `model(..., infer=False)` -> loss -> `backward()`), `vocoder._FilterStageB`
re-evaluates stage B here under autograd in its backward and hands the gradients of the three control tensors
back -- the gradients autograd derives for ddsp/vocoder.py:397-421 / :521-548 and ddsp/core.py:185-336, because
this file computes the same function: per-frame impulse responses `irfft(magnitudes)` rolled and windowed,
Bartlett-framed FFT convolution, overlap-add, delay-compensated crop.  f0 / phase / noise are data (no gradient),
as in the reference graph.  Runs on any device (the CPU tests pin it to gradients recorded from the reference).
"""
import math

import torch
import torch.nn.functional as F


def impulse_response(magnitudes, window='hann', half_width_frames=None):
    """(B, Frame, n_mag) complex -> (B, Frame, L) real, L = 2 (n_mag - 1): inverse real DFT, rolled to the centre,
    windowed (core.py:306-328: none `:326`, periodic Hann `:242-289`, dynamic cosine `:292-303`)."""
    ir = torch.fft.irfft(magnitudes)
    L = ir.shape[-1]
    ir = torch.roll(ir, L // 2, dims=-1)
    if window == 'hann':
        ir = ir * torch.hann_window(L, periodic=True, dtype=ir.dtype, device=ir.device)
    elif window == 'dynamic':
        n = torch.arange(-(L // 2), (L + 1) // 2, dtype=ir.dtype, device=ir.device)
        x = n / half_width_frames                                   # (B, Frame, L); half width in samples
        x = torch.where(x > 1, torch.zeros_like(x), x)              # only x > 1 is cleared (-> weight 1), as the reference does
        ir = ir * ((1 + torch.cos(math.pi * x)) / 2)
    elif window != 'none':
        raise ValueError(window)
    return ir


def ltv_fir(audio, ir, hop):
    """Linear time-varying FIR of core.py:185-239: y[t] = sum_s x[s] h_s[t + L//2 - s] with h_s the frame impulse
    responses linearly interpolated over input time (Bartlett frames of 2 hop, one FFT convolution per frame)."""
    B, T = audio.shape
    n_frames, L = ir.shape[1], ir.shape[2]
    if audio.shape[0] != ir.shape[0]:
        raise ValueError('Batch size of audio and impulse response must be the same')
    frame = 2 * hop
    frames = F.pad(audio, (hop, hop)).unfold(1, frame, hop)                                  # (B, Frame + 1, 2 hop)
    frames = frames * torch.bartlett_window(frame, periodic=True, dtype=audio.dtype, device=audio.device)
    ir = torch.cat([ir, ir[:, -1:]], dim=1)                                                  # last response held
    nfft = 1 << (frame + L - 2).bit_length()                                                 # covers the linear convolution
    y = torch.fft.irfft(torch.fft.rfft(frames, nfft) * torch.fft.rfft(ir, nfft), nfft)       # (B, Frame + 1, nfft)
    total = n_frames * hop + nfft
    out = F.fold(y.transpose(1, 2), (1, total), (1, nfft), stride=(1, hop)).reshape(B, total)
    start = hop + L // 2
    return out[:, start:start + T]


def frequency_filter(audio, magnitudes, window='hann', half_width_frames=None):
    hop = audio.shape[1] // magnitudes.shape[1]
    return ltv_fir(audio, impulse_response(magnitudes, window, half_width_frames), hop)


def upsample(x, hop):
    """(B, Frame, C) -> (B, Frame*hop, C), linear with the last frame held (core.py:7-21)."""
    x = torch.cat([x, x[:, -1:]], dim=1).transpose(1, 2)
    y = F.interpolate(x, size=(x.shape[-1] - 1) * hop + 1, mode='linear', align_corners=True)
    return y[..., :-1].transpose(1, 2)


def combsub_stage(group_delay, harmonic_magnitude, noise_magnitude, f0_frames, rot, noise_u, hop, sr):
    """vocoder.py:521-548.  rot (B,T): wrapped rotation of stage A; noise_u (B,T) in [0,1)."""
    f0 = upsample(f0_frames.reshape(f0_frames.shape[0], -1, 1), hop)[..., 0]
    comb = torch.sinc(sr * rot / (f0 + 1e-3))
    allpass = torch.exp(1j * torch.cumsum(math.pi * torch.tanh(group_delay), dim=-1))
    harmonic = frequency_filter(comb, allpass, 'none')
    hw = 1.5 * sr / (f0_frames.reshape(f0_frames.shape[0], -1, 1) + 1e-3)
    harmonic = frequency_filter(harmonic, torch.complex(torch.exp(harmonic_magnitude), torch.zeros_like(harmonic_magnitude)),
                                'dynamic', hw)
    noise_param = torch.exp(noise_magnitude) / 128
    noise = frequency_filter(noise_u * 2 - 1, torch.complex(noise_param, torch.zeros_like(noise_param)), 'hann')
    return harmonic + noise, harmonic, noise


def sins_stage(amplitudes, group_delay, noise_magnitude, f0_frames, phase, noise_u, hop, sr, chunk=32):
    """vocoder.py:397-421.  phase (B,T): full-rate phase 2 pi rot of stage A."""
    f0f = f0_frames.reshape(f0_frames.shape[0], -1, 1)
    n_harm = amplitudes.shape[-1]
    level = torch.arange(1, n_harm + 1, dtype=amplitudes.dtype, device=amplitudes.device)
    amp = torch.exp(amplitudes) / 128 * ((f0f * level < sr / 2).to(amplitudes.dtype) + 1e-7)     # core.py:24-28
    sinusoids = torch.zeros_like(phase)
    for k0 in range(0, n_harm, chunk):                                                            # bounds the (B,T,chunk) temporaries
        a = upsample(amp[..., k0:k0 + chunk], hop)
        sinusoids = sinusoids + (a * torch.sin(phase.unsqueeze(-1) * level[k0:k0 + chunk])).sum(-1)
    allpass = torch.exp(1j * torch.cumsum(math.pi * torch.tanh(group_delay), dim=-1))
    harmonic = frequency_filter(sinusoids, allpass, 'none')
    noise_param = torch.exp(noise_magnitude) / 128
    noise = frequency_filter(noise_u * 2 - 1, torch.complex(noise_param, torch.zeros_like(noise_param)), 'hann')
    return harmonic + noise, harmonic, noise
