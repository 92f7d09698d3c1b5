"""Stock-PyTorch restatement of the CombSubFast synthesizer path (stage A + stage B), op for op as
`ddsp/vocoder.py:446-492` + `ddsp/core.py:7-51` issue them, runnable on CUDA.

TEST / BENCH INFRASTRUCTURE ONLY (same rule as ddsp_oracle.py).  It exists because the reference
tree cannot travel to the GPU box: `bench.py --torch-port` times this on the same B200 as "what
the stock PyTorch ops cost", the practical bar named in BASELINE.md §3, and the GPU tests use it as
a second fp32 reference of the same op sequence.  Nothing in the product imports it.
"""
import numpy as np
import torch
import torch.nn.functional as F


def upsample(signal, factor):                                   # core.py:7-21
    signal = signal.permute(0, 2, 1)
    signal = F.interpolate(torch.cat((signal, signal[:, :, -1:]), 2), size=signal.shape[-1] * factor + 1,
                           mode='linear', align_corners=True)
    return signal[:, :, :-1].permute(0, 2, 1)


def fo_to_rot(fo, sr, initial_phase=None, precise=False):       # core.py:31-51
    _fo = fo.double() if precise else fo
    rot = torch.cumsum(_fo / sr, axis=1)
    if initial_phase is not None:
        rot += initial_phase.unsqueeze(-1).to(rot) / 2 / np.pi
    rot = rot - torch.round(rot)
    return rot.to(fo)


def combsubfast_forward(harmo_mag, harmo_phase, noise_mag, f0_frames, window, noise_u=None, sr=44100, block_size=512,
                        initial_phase=None, infer=True):
    """vocoder.py:446-492 with `unit2ctrl` replaced by the given control tensors and `torch.rand_like`
    by `noise_u` when supplied.  f0_frames (B,F,1).  Returns (signal, phase_frames)."""
    pad = block_size
    f0 = upsample(f0_frames, block_size).squeeze(-1)
    rot = fo_to_rot(f0, sr, initial_phase, infer)
    phase_frames = 2 * np.pi * rot[:, ::block_size]
    combtooth = torch.sinc(sr * rot / (f0 + 1e-3))
    combtooth[f0 <= 0.] = 0.
    noise = (torch.rand_like(combtooth) if noise_u is None else noise_u) * 2 - 1
    combtooth = F.pad(combtooth, (pad, pad))
    noise = F.pad(noise, (pad, pad))
    combtooth_frames = combtooth.unfold(1, 2 * block_size, block_size) * window
    noise_frames = noise.unfold(1, 2 * block_size, block_size) * window
    _src = torch.exp(harmo_mag + 1.j * np.pi * harmo_phase)
    src_filter = torch.cat((_src, _src[:, -1:, :]), 1)
    _nf = torch.exp(noise_mag) / 128
    noise_filter = torch.cat((_nf, _nf[:, -1:, :]), 1)
    signal_fft = torch.fft.rfft(combtooth_frames, 2 * block_size) * src_filter \
        + torch.fft.rfft(noise_frames, 2 * block_size) * noise_filter
    frames_out = torch.fft.irfft(signal_fft, 2 * block_size) * window
    fold = torch.nn.Fold(output_size=(1, (frames_out.size(1) + 1) * block_size), kernel_size=(1, 2 * block_size),
                         stride=(1, block_size))
    signal = fold(frames_out.transpose(1, 2))[:, 0, 0, pad:-pad]
    return signal, phase_frames
