"""Host-side mirror of the reference's `ddsp/core.py` for the synthesizer path.

Same function names and argument meaning as the reference (`upsample`, `fo_to_rot`,
`remove_above_fmax`, `frequency_filter`), but every function takes CUDA float32 tensors and
enqueues hand-written sm_100a kernels through the C ABI (include/ddsp_b200.h) on the current
stream.  There is no CPU path: a CPU tensor or a missing library raises.
"""
import math

import torch

from . import _cabi

_TWO62 = 1 << 62


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _cabi.DDSPB200Error(f'{name} must be a CUDA tensor (ddsp_b200 has no CPU path)')
    if t.dtype != torch.float32:
        raise TypeError(f'{name} must be float32, got {t.dtype}')
    return t


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
def upsample(signal, factor):
    """(B, Frame, C) -> (B, Frame*factor, C); reference ddsp/core.py:7-21."""
    signal = _need_cuda_f32(signal, 'signal')
    if signal.dim() != 3:
        raise ValueError('signal must be (B, Frame, C)')
    B, F, Cc = signal.shape
    factor = int(factor)
    y = torch.empty((B, F * factor, Cc), dtype=torch.float32, device=signal.device)
    with torch.cuda.device(signal.device):
        _cabi.check(_cabi.lib().ddsp_b200_upsample(signal.data_ptr(), signal.stride(0), signal.stride(1),
                                                   signal.stride(2), B, F, Cc, factor, y.data_ptr(), _stream()))
    return y


def fo_to_rot(fo, sr, initial_phase=None, precise=False):
    """(B, T) Hz -> wrapped rotation in [-0.5, 0.5]; reference ddsp/core.py:31-51."""
    fo = _need_cuda_f32(fo, 'fo').contiguous()
    B, T = fo.shape
    L = _cabi.lib()
    ws = torch.empty(max(1, L.ddsp_b200_fo_to_rot_workspace_bytes(B, T)), dtype=torch.uint8, device=fo.device)
    ip = None if initial_phase is None else _need_cuda_f32(initial_phase.to(fo.device, torch.float32), 'initial_phase').contiguous()
    rot = torch.empty_like(fo)
    with torch.cuda.device(fo.device):
        _cabi.check(L.ddsp_b200_fo_to_rot(fo.data_ptr(), B, T, float(sr), _ptr(ip), int(bool(precise)),
                                          rot.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return rot


def remove_above_fmax(amplitudes, pitch, fmax, level_start=1):
    """amplitudes (B,F,K) * ((pitch*k < fmax) + 1e-7); reference ddsp/core.py:24-28."""
    amplitudes = _need_cuda_f32(amplitudes, 'amplitudes')
    pitch = _need_cuda_f32(pitch, 'pitch')
    if amplitudes.stride(-1) != 1:
        amplitudes = amplitudes.contiguous()
    B, F, K = amplitudes.shape
    p2 = pitch.reshape(B, F)
    out = torch.empty((B, F, K), dtype=torch.float32, device=amplitudes.device)
    with torch.cuda.device(amplitudes.device):
        _cabi.check(_cabi.lib().ddsp_b200_remove_above_fmax(
            amplitudes.data_ptr(), amplitudes.stride(0), amplitudes.stride(1), p2.data_ptr(), p2.stride(0),
            p2.stride(1), float(fmax), int(level_start), B, F, K, out.data_ptr(), _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# fused stages used by the synthesizer modules
# ------------------------------------------------------------------------------------------------
def _f0_2d(f0_frames):
    f0_frames = _need_cuda_f32(f0_frames, 'f0_frames')
    if f0_frames.dim() == 3:
        if f0_frames.shape[-1] != 1:
            raise ValueError('f0_frames must be (B, Frame, 1)')
        f0_frames = f0_frames[..., 0]
    if f0_frames.dim() != 2:
        raise ValueError('f0_frames must be (B, Frame, 1) or (B, Frame)')
    return f0_frames


def _init_phase(initial_phase, B, device):
    if initial_phase is None:
        return None
    ip = torch.as_tensor(initial_phase, dtype=torch.float32, device=device).reshape(-1).contiguous()
    if ip.numel() != B:
        raise ValueError('initial_phase must have one entry per clip')
    return ip


def phase_stage(f0_frames, block_size, sampling_rate, initial_phase=None, infer=True, full_rate=False):
    """Stage A (vocoder.py:391-393 / 449-451 / 515-517).

    Returns (phase_frames (B,F) fp32, prefix (B,F) fp64 workspace for stage B,
    phase (B,T) fp32 at sample rate if `full_rate` else None)."""
    f0 = _f0_2d(f0_frames)
    B, F = f0.shape
    hop = int(block_size)
    dev = f0.device
    phase_frames = torch.empty((B, F), dtype=torch.float32, device=dev)
    prefix = torch.empty((B, F), dtype=torch.float64, device=dev)
    phase_full = torch.empty((B, F * hop), dtype=torch.float32, device=dev) if full_rate else None
    ip = _init_phase(initial_phase, B, dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().ddsp_b200_phase(f0.data_ptr(), f0.stride(0), f0.stride(1), B, F, hop,
                                                float(sampling_rate), _ptr(ip), int(bool(infer)),
                                                phase_frames.data_ptr(), prefix.data_ptr(), _ptr(phase_full),
                                                _stream()))
    return phase_frames, prefix, phase_full


def _common_views(tensors, names):
    """Control tensors arrive as non-contiguous `torch.split` views of one (B,F,sumK) tensor
    (unit2control.py:10-20).  Pass them through untouched when they share (batch,row) strides and
    have unit inner stride; otherwise gather them into one packed buffer (still on the GPU)."""
    ts = [_need_cuda_f32(t, n) for t, n in zip(tensors, names)]
    s0 = (ts[0].stride(0), ts[0].stride(1))
    if all(t.dim() == 3 and t.stride(2) == 1 and (t.stride(0), t.stride(1)) == s0 for t in ts):
        return ts
    packed = torch.cat([t.contiguous() for t in ts], dim=-1)
    return list(torch.split(packed, [t.shape[-1] for t in ts], dim=-1))


def combsubfast_stage(harmonic_magnitude, harmonic_phase, noise_magnitude, f0_frames, prefix, block_size,
                      sampling_rate, initial_phase=None, noise_u=None, seed=0, window=None, out=None):
    """Stage B of CombSubFast.forward (vocoder.py:455-492) -> signal (B,T)."""
    hm, hp, nm = _common_views((harmonic_magnitude, harmonic_phase, noise_magnitude),
                               ('harmonic_magnitude', 'harmonic_phase', 'noise_magnitude'))
    f0 = _f0_2d(f0_frames)
    B, F = f0.shape
    hop = int(block_size)
    for t in (hm, hp, nm):
        if tuple(t.shape) != (B, F, hop + 1):
            raise ValueError(f'control tensors must be (B, Frame, {hop + 1}); got {tuple(t.shape)}')
    dev = f0.device
    T = F * hop
    if noise_u is not None:
        noise_u = _need_cuda_f32(noise_u, 'noise_u').contiguous()
        if tuple(noise_u.shape) != (B, T):
            raise ValueError('noise_u must be (B, T)')
    if window is not None:
        window = _need_cuda_f32(window, 'window').contiguous()
        if window.numel() != 2 * hop:
            raise ValueError('window must have 2*block_size entries')
    ip = _init_phase(initial_phase, B, dev)
    signal = out if out is not None else torch.empty((B, T), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().ddsp_b200_combsubfast(
            hm.data_ptr(), hp.data_ptr(), nm.data_ptr(), hm.stride(0), hm.stride(1), f0.data_ptr(), f0.stride(0),
            f0.stride(1), prefix.data_ptr(), _ptr(ip), _ptr(noise_u), int(seed) % _TWO62, _ptr(window), B, F, hop,
            float(sampling_rate), signal.data_ptr(), _stream()))
    return signal


def last_launch_count():
    return _cabi.lib().ddsp_b200_last_launch_count()
