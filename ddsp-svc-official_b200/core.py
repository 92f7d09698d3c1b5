"""Host-side mirror of the reference's `ddsp/core.py` for the synthesizer path.

Same function names and argument meaning as the reference (`upsample`, `fo_to_rot`,
`remove_above_fmax`, `frequency_filter`), but every function takes CUDA float32 tensors and
enqueues hand-written sm_100a kernels through the C ABI (include/ddsp_b200.h) on the current
stream.  There is no CPU path: a CPU tensor or a missing library raises.
"""
import math

import torch

from . import _cabi
from . import _torchext

_TWO62 = 1 << 62


class _OnDevice:
    """`with _OnDevice(dev) as stream:` -- makes `dev` current only when it is not already (the
    torch.cuda.device context manager and torch.cuda.current_stream() cost ~15 us per call, more than
    a streaming-size synthesis itself) and yields the raw handle of torch's current stream on it."""
    __slots__ = ('idx', 'prev')

    def __init__(self, dev):
        self.idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self.prev = -1

    def __enter__(self):
        cur = torch._C._cuda_getDevice()
        if cur != self.idx:
            self.prev = cur
            torch._C._cuda_setDevice(self.idx)
        return torch._C._cuda_getCurrentRawStream(self.idx)

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch._C._cuda_setDevice(self.prev)
        return False


def _need_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _cabi.DDSPB200Error(f'{name} must be a CUDA tensor (ddsp_b200 has no CPU path)')
    if t.dtype != torch.float32:
        raise TypeError(f'{name} must be float32, got {t.dtype}')
    return t


def as_f32(t):
    """Module-level convenience: the reference accepts any floating dtype (AMP / fp64 experiments);
    the kernels compute in fp32 (+fp64 phase), so other float dtypes are cast on the way in."""
    if isinstance(t, torch.Tensor) and t.is_floating_point() and t.dtype != torch.float32:
        return t.float()
    return t


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _op(fn, *args):
    """Call an operator of the PyTorch-extension host; ValueError / TypeError pass through (shape, dtype), every
    other failure (CPU tensor, unsupported configuration, CUDA error) surfaces as DDSPB200Error like the ctypes path."""
    try:
        return fn(*args)
    except (ValueError, TypeError):
        raise
    except RuntimeError as e:
        raise _cabi.DDSPB200Error(str(e)) from None


# ------------------------------------------------------------------------------------------------
def upsample(signal, factor):
    """(B, Frame, C) -> (B, Frame*factor, C); reference ddsp/core.py:7-21."""
    signal = _need_cuda_f32(signal, 'signal')
    if signal.dim() != 3:
        raise ValueError('signal must be (B, Frame, C)')
    B, F, Cc = signal.shape
    factor = int(factor)
    y = torch.empty((B, F * factor, Cc), dtype=torch.float32, device=signal.device)
    with _OnDevice(signal.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_upsample(signal.data_ptr(), signal.stride(0), signal.stride(1),
                                                   signal.stride(2), B, F, Cc, factor, y.data_ptr(), _st))
    return y


def fo_to_rot(fo, sr, initial_phase=None, precise=False):
    """(B, T) Hz -> wrapped rotation in [-0.5, 0.5]; reference ddsp/core.py:31-51."""
    fo = _need_cuda_f32(fo, 'fo').contiguous()
    B, T = fo.shape
    L = _cabi.lib()
    ws = torch.empty(max(1, L.ddsp_b200_fo_to_rot_workspace_bytes(B, T)), dtype=torch.uint8, device=fo.device)
    ip = None if initial_phase is None else _need_cuda_f32(initial_phase.to(fo.device, torch.float32), 'initial_phase').contiguous()
    rot = torch.empty_like(fo)
    with _OnDevice(fo.device) as _st:
        _cabi.check(L.ddsp_b200_fo_to_rot(fo.data_ptr(), B, T, float(sr), _ptr(ip), int(bool(precise)),
                                          rot.data_ptr(), ws.data_ptr(), ws.numel(), _st))
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
    with _OnDevice(amplitudes.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_remove_above_fmax(
            amplitudes.data_ptr(), amplitudes.stride(0), amplitudes.stride(1), p2.data_ptr(), p2.stride(0),
            p2.stride(1), float(fmax), int(level_start), B, F, K, out.data_ptr(), _st))
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
    ops = _torchext.ops()
    if ops is not None:     # the PyTorch-extension host: checks, allocation and the C-ABI call in C++
        ip = None if initial_phase is None else torch.as_tensor(initial_phase, dtype=torch.float32, device=f0_frames.device)
        pf, prefix, full = _op(ops.phase, f0_frames, int(block_size), float(sampling_rate), ip, bool(infer), bool(full_rate), None)
        return pf, prefix, (full if full_rate else None)
    f0 = _f0_2d(f0_frames)
    B, F = f0.shape
    hop = int(block_size)
    dev = f0.device
    phase_frames = torch.empty((B, F), dtype=torch.float32, device=dev)
    prefix = torch.empty((B, F), dtype=torch.float64, device=dev)
    phase_full = torch.empty((B, F * hop), dtype=torch.float32, device=dev) if full_rate else None
    ip = _init_phase(initial_phase, B, dev)
    with _OnDevice(dev) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_phase(f0.data_ptr(), f0.stride(0), f0.stride(1), B, F, hop,
                                                float(sampling_rate), _ptr(ip), int(bool(infer)),
                                                phase_frames.data_ptr(), prefix.data_ptr(), _ptr(phase_full),
                                                _st))
    return phase_frames, prefix, phase_full


def phase_stage_stream(f0_frames, block_size, sampling_rate, carry=None, initial_phase=None, full_rate=False):
    """Stage A for a block that continues a stream (gui.py:373-388).

    `carry` (B,) float64 view (any stride): the prefix the stream reached at this block's first frame --
    a column of the previous block's `prefix`; None at stream start (then `initial_phase` applies).
    Returns (phase_frames (B,F) fp32, prefix (B,F) fp64) and, with `full_rate` (Sins), the sample-rate phase (B,T)."""
    ops = _torchext.ops()
    if ops is not None and carry is not None and not full_rate:
        ip = None if initial_phase is None else torch.as_tensor(initial_phase, dtype=torch.float32, device=f0_frames.device)
        pf, prefix, _ = _op(ops.phase, f0_frames, int(block_size), float(sampling_rate), ip, True, False, carry)
        return pf, prefix
    f0 = _f0_2d(f0_frames)
    B, F = f0.shape
    dev = f0.device
    phase_frames = torch.empty((B, F), dtype=torch.float32, device=dev)
    prefix = torch.empty((B, F), dtype=torch.float64, device=dev)
    ip = _init_phase(initial_phase, B, dev)
    cptr, cstride = 0, 0
    if carry is not None:
        if not carry.is_cuda or carry.dtype != torch.float64 or carry.dim() != 1 or carry.numel() != B:
            raise ValueError('carry must be a CUDA float64 tensor with one entry per clip')
        cptr, cstride = carry.data_ptr(), carry.stride(0)
    if full_rate:
        phase_full = torch.empty((B, F * int(block_size)), dtype=torch.float32, device=dev)
        with _OnDevice(dev) as _st:
            _cabi.check(_cabi.lib().ddsp_b200_phase_stream_full(f0.data_ptr(), f0.stride(0), f0.stride(1), B, F, int(block_size),
                                                                float(sampling_rate), _ptr(ip), cptr, cstride,
                                                                phase_frames.data_ptr(), prefix.data_ptr(), phase_full.data_ptr(), _st))
        return phase_frames, prefix, phase_full
    with _OnDevice(dev) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_phase_stream(f0.data_ptr(), f0.stride(0), f0.stride(1), B, F, int(block_size),
                                                       float(sampling_rate), _ptr(ip), cptr, cstride,
                                                       phase_frames.data_ptr(), prefix.data_ptr(), _st))
    return phase_frames, prefix


def combsubfast_synth(harmonic_magnitude, harmonic_phase, noise_magnitude, f0_frames, block_size, sampling_rate,
                      initial_phase=None, noise_u=None, seed=0, window=None, seed_device=None):
    """Stage A + stage B of CombSubFast (vocoder.py:449-451, 455-492) in one operator of the extension host, for callers
    whose control rows do not depend on the phase (rows from outside the module: a streaming plugin, the benchmarks).
    Returns (signal (B,T), phase_frames (B,F), prefix (B,F) fp64) -- the same tensors as `phase_stage` followed by
    `combsubfast_stage`."""
    ops = _torchext.ops()
    if ops is None:
        pf, prefix, _ = phase_stage(f0_frames, block_size, sampling_rate, initial_phase)
        return combsubfast_stage(harmonic_magnitude, harmonic_phase, noise_magnitude, f0_frames, prefix, block_size,
                                 sampling_rate, noise_u=noise_u, seed=seed, window=window, seed_device=seed_device), pf, prefix
    ip = None if initial_phase is None else torch.as_tensor(initial_phase, dtype=torch.float32, device=f0_frames.device)
    return _op(ops.combsubfast_ab, harmonic_magnitude, harmonic_phase, noise_magnitude, f0_frames, int(block_size),
               float(sampling_rate), ip, None, noise_u, int(seed) % _TWO62, window, seed_device, 0)


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
                      sampling_rate, initial_phase=None, noise_u=None, seed=0, window=None, out=None, hop_offset=None,
                      seed_device=None):
    """Stage B of CombSubFast.forward (vocoder.py:455-492) -> signal (B,T).
    `hop_offset` (streaming only): stream index of this block's first hop, see `phase_stage_stream`.
    `seed_device` (CUDA graphs only): one-element int64 CUDA tensor added to `seed` on the device when the
    kernel starts, so that a captured call draws fresh in-kernel noise on every replay."""
    ops = _torchext.ops()
    if ops is not None:     # (initial_phase is carried by `prefix`)
        return _op(ops.combsubfast, harmonic_magnitude, harmonic_phase, noise_magnitude, f0_frames, prefix, int(block_size),
                               float(sampling_rate), noise_u, int(seed) % _TWO62, window, seed_device, int(hop_offset or 0), out)
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
    with _OnDevice(dev) as _st:
        if seed_device is not None and (not seed_device.is_cuda or seed_device.dtype != torch.int64
                                        or seed_device.numel() != 1):
            raise ValueError('seed_device must be a one-element int64 CUDA tensor')
        if hop_offset is None and seed_device is None:
            _cabi.check(_cabi.lib().ddsp_b200_combsubfast(
                hm.data_ptr(), hp.data_ptr(), nm.data_ptr(), hm.stride(0), hm.stride(1), f0.data_ptr(), f0.stride(0),
                f0.stride(1), prefix.data_ptr(), _ptr(ip), _ptr(noise_u), int(seed) % _TWO62, _ptr(window), B, F, hop,
                float(sampling_rate), signal.data_ptr(), _st))
        else:
            _cabi.check(_cabi.lib().ddsp_b200_combsubfast_stream(
                hm.data_ptr(), hp.data_ptr(), nm.data_ptr(), hm.stride(0), hm.stride(1), f0.data_ptr(), f0.stride(0),
                f0.stride(1), prefix.data_ptr(), _ptr(noise_u), int(seed) % _TWO62, _ptr(seed_device),
                int(hop_offset or 0), _ptr(window), B, F, hop, float(sampling_rate), signal.data_ptr(), _st))
    return signal


def combsubfast_backward_stage(grad_signal, harmonic_magnitude, harmonic_phase, noise_magnitude, f0_frames, prefix,
                               block_size, sampling_rate, noise_u=None, seed=0, window=None):
    """Gradient of `combsubfast_stage` w.r.t. the three control tensors (what autograd derives for
    vocoder.py:455-492).  All other arguments must be those of the forward call.  Returns three
    (B,F,513) views of one (B,F,1539) tensor, in the order of the inputs."""
    hm, hp, nm = _common_views((harmonic_magnitude, harmonic_phase, noise_magnitude),
                               ('harmonic_magnitude', 'harmonic_phase', 'noise_magnitude'))
    f0 = _f0_2d(f0_frames)
    B, F = f0.shape
    hop = int(block_size)
    for t in (hm, hp, nm):
        if tuple(t.shape) != (B, F, hop + 1):
            raise ValueError(f'control tensors must be (B, Frame, {hop + 1}); got {tuple(t.shape)}')
    grad_signal = _need_cuda_f32(grad_signal, 'grad_signal').contiguous()
    if tuple(grad_signal.shape) != (B, F * hop):
        raise ValueError('grad_signal must be (B, T)')
    if noise_u is not None:
        noise_u = _need_cuda_f32(noise_u, 'noise_u').contiguous()
        if tuple(noise_u.shape) != (B, F * hop):
            raise ValueError('noise_u must be (B, T)')
    if window is not None:
        window = _need_cuda_f32(window, 'window').contiguous()
        if window.numel() != 2 * hop:
            raise ValueError('window must have 2*block_size entries')
    grads = torch.empty((B, F, 3 * (hop + 1)), dtype=torch.float32, device=f0.device)
    ghm, ghp, gnm = torch.split(grads, hop + 1, dim=-1)
    with _OnDevice(f0.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_combsubfast_backward(
            hm.data_ptr(), hp.data_ptr(), nm.data_ptr(), hm.stride(0), hm.stride(1), f0.data_ptr(), f0.stride(0),
            f0.stride(1), prefix.data_ptr(), _ptr(noise_u), int(seed) % _TWO62, _ptr(window), grad_signal.data_ptr(),
            B, F, hop, float(sampling_rate), ghm.data_ptr(), ghp.data_ptr(), gnm.data_ptr(), ghm.stride(0),
            ghm.stride(1), _st))
    return ghm, ghp, gnm


WINDOW_NONE, WINDOW_HANN, WINDOW_DYNAMIC = 0, 1, 2
MAG_REAL, MAG_EXP, MAG_ALLPASS_TANH, MAG_COMPLEX = 0, 1, 2, 3


def frequency_filter(audio, magnitudes, hann_window=True, half_width_frames=None, *, f0_frames=None,
                     sampling_rate=44100, encoding=None, mag_scale=1.0):
    """Apply a per-frame linear-phase LTV-FIR (reference ddsp/core.py:331-336).

    audio (B,T) float32; magnitudes (B,F,n_mag) real float32 or complex64, n_mag in {256, 512} (complex with
    n_mag = 512 only with a zero imaginary part, the reference's `torch.complex(x, zeros)` idiom), T = 512*F (other hop
    sizes raise).  `half_width_frames` (B,F,1) selects the dynamic cosine window exactly as in the
    reference; it must be the synthesizer's `1.5*sr/(f0_frames+1e-3)` (vocoder.py:542) -- pass the
    generating `f0_frames` instead and the kernel derives it.  `encoding`/`mag_scale` let the
    synthesizer modules hand over raw control tensors (exp / all-pass) without materialising the
    complex magnitudes.
    """
    audio = _need_cuda_f32(audio, 'audio').contiguous()
    B, T = audio.shape
    if magnitudes.shape[0] != B:
        raise ValueError(f'Batch size of audio ({B}) and impulse response ({magnitudes.shape[0]}) must be the same.')
    if encoding is None:
        if magnitudes.is_complex() and magnitudes.shape[-1] == 512:
            # The reference's own call with n_mag = 512 is `torch.complex(src_param, zeros)` (vocoder.py:541): the L = 1022
            # kernels evaluate real (symmetric) magnitude responses.  A zero imaginary part is verified here (one device
            # sync; this mirror is not on the modules' hot path) and the real part handed on; a genuinely complex 512-bin
            # response is not supported.
            if bool((magnitudes.imag != 0).any()):
                raise _cabi.DDSPB200Error('frequency_filter: complex magnitudes with n_mag = 512 must have a zero imaginary '
                                          'part (vocoder.py:541); only n_mag = 256 takes arbitrary complex responses')
            magnitudes = magnitudes.real.to(torch.float32)
            encoding, n_mag = MAG_REAL, 512
        elif magnitudes.is_complex():
            magnitudes = torch.view_as_real(magnitudes.to(torch.complex64).contiguous()).reshape(
                magnitudes.shape[0], magnitudes.shape[1], -1)
            encoding, n_mag = MAG_COMPLEX, magnitudes.shape[-1] // 2
        else:
            encoding, n_mag = MAG_REAL, magnitudes.shape[-1]
    else:
        n_mag = magnitudes.shape[-1]
    mags = _need_cuda_f32(magnitudes, 'magnitudes')
    if mags.stride(-1) != 1:
        mags = mags.contiguous()
    F = mags.shape[1]
    if T != F * 512:
        raise _cabi.DDSPB200Error('frequency_filter: only hop = T/Frame = 512 is supported')
    if not hann_window:
        window = WINDOW_NONE
    elif half_width_frames is None and f0_frames is None:
        window = WINDOW_HANN
    else:
        window = WINDOW_DYNAMIC
        if f0_frames is None:      # invert half_width = 1.5*sr/(f0+1e-3)
            f0_frames = 1.5 * float(sampling_rate) / half_width_frames.reshape(B, F).to(torch.float32) - 1e-3
    f0 = None if f0_frames is None else _f0_2d(f0_frames)
    out = torch.empty_like(audio)
    L = _cabi.lib()
    ws = torch.empty(L.ddsp_b200_frequency_filter_workspace_bytes(B, F, n_mag), dtype=torch.uint8, device=audio.device)
    with _OnDevice(audio.device) as _st:
        _cabi.check(L.ddsp_b200_frequency_filter(
            audio.data_ptr(), mags.data_ptr(), mags.stride(0), mags.stride(1), n_mag, encoding, float(mag_scale),
            window, _ptr(f0), 0 if f0 is None else f0.stride(0), 0 if f0 is None else f0.stride(1),
            float(sampling_rate), B, F, 512, out.data_ptr(), 0, ws.data_ptr(), ws.numel(), _st))
    return out


def _check_noise(noise_u, B, T):
    if noise_u is None:
        return None
    noise_u = _need_cuda_f32(noise_u, 'noise_u').contiguous()
    if tuple(noise_u.shape) != (B, T):
        raise ValueError('noise_u must be (B, T)')
    return noise_u


def combsub_stage(group_delay, harmonic_magnitude, noise_magnitude, f0_frames, prefix, block_size, sampling_rate,
                  noise_u=None, seed=0, hop_offset=0):
    """Stage B of CombSub.forward (old) (vocoder.py:521-548) -> (signal, harmonic, noise), each (B,T).
    `hop_offset` (streaming): stream index of the first hop, so that the in-kernel noise of a hop repeats."""
    ops = _torchext.ops()
    if ops is not None and not hop_offset:
        return _op(ops.combsub, group_delay, harmonic_magnitude, noise_magnitude, f0_frames, prefix, int(block_size),
                           float(sampling_rate), noise_u, int(seed) % _TWO62)
    gd, hm, nm = _common_views((group_delay, harmonic_magnitude, noise_magnitude),
                               ('group_delay', 'harmonic_magnitude', 'noise_magnitude'))
    f0 = _f0_2d(f0_frames)
    B, F = f0.shape
    hop = int(block_size)
    dev = f0.device
    T = F * hop
    noise_u = _check_noise(noise_u, B, T)
    L = _cabi.lib()
    ws = torch.empty(L.ddsp_b200_combsub_workspace_bytes(B, F, gd.shape[-1], hm.shape[-1], nm.shape[-1]),
                     dtype=torch.uint8, device=dev)
    signal, harmonic, noise = (torch.empty((B, T), dtype=torch.float32, device=dev) for _ in range(3))
    with _OnDevice(dev) as _st:
        _cabi.check(L.ddsp_b200_combsub_stream(
            gd.data_ptr(), gd.shape[-1], hm.data_ptr(), hm.shape[-1], nm.data_ptr(), nm.shape[-1], gd.stride(0),
            gd.stride(1), f0.data_ptr(), f0.stride(0), f0.stride(1), prefix.data_ptr(), _ptr(noise_u),
            int(seed) % _TWO62, int(hop_offset), B, F, hop, float(sampling_rate), signal.data_ptr(), harmonic.data_ptr(),
            noise.data_ptr(), ws.data_ptr(), ws.numel(), _st))
    return signal, harmonic, noise


def sins_stage(amplitudes, group_delay, noise_magnitude, f0_frames, phase, block_size, sampling_rate, noise_u=None,
               seed=0, hop_offset=0):
    """Stage B of Sins.forward (vocoder.py:397-421) -> (signal, harmonic, noise), each (B,T).
    `phase` is the full-rate phase (B,T) from `phase_stage(..., full_rate=True)`; `hop_offset` as in `combsub_stage`."""
    ops = _torchext.ops()
    if ops is not None and not hop_offset:
        return _op(ops.sins, amplitudes, group_delay, noise_magnitude, f0_frames, phase, int(block_size), float(sampling_rate),
                        noise_u, int(seed) % _TWO62)
    am, gd, nm = _common_views((amplitudes, group_delay, noise_magnitude),
                               ('amplitudes', 'group_delay', 'noise_magnitude'))
    f0 = _f0_2d(f0_frames)
    B, F = f0.shape
    hop = int(block_size)
    dev = f0.device
    T = F * hop
    noise_u = _check_noise(noise_u, B, T)
    phase = _need_cuda_f32(phase, 'phase').contiguous()
    L = _cabi.lib()
    ws = torch.empty(L.ddsp_b200_sins_workspace_bytes(B, F, am.shape[-1], gd.shape[-1], nm.shape[-1]),
                     dtype=torch.uint8, device=dev)
    signal, harmonic, noise = (torch.empty((B, T), dtype=torch.float32, device=dev) for _ in range(3))
    with _OnDevice(dev) as _st:
        _cabi.check(L.ddsp_b200_sins_stream(
            am.data_ptr(), am.shape[-1], gd.data_ptr(), gd.shape[-1], nm.data_ptr(), nm.shape[-1], am.stride(0),
            am.stride(1), f0.data_ptr(), f0.stride(0), f0.stride(1), phase.data_ptr(), _ptr(noise_u),
            int(seed) % _TWO62, int(hop_offset), B, F, hop, float(sampling_rate), signal.data_ptr(), harmonic.data_ptr(),
            noise.data_ptr(), ws.data_ptr(), ws.numel(), _st))
    return signal, harmonic, noise


def apply_frame_mask_(signal, mask_frames, block_size=512):
    """In place `signal *= upsample(mask_frames, block_size).squeeze(-1)` -- the silence-mask epilogue of
    main.py:112-116,159 and gui.py:108-112,127 -- in one pass, without the (B,T) mask tensor.
    signal (B,T) fp32 contiguous; mask_frames (B,F) or (B,F,1)."""
    signal = _need_cuda_f32(signal, 'signal')
    if not signal.is_contiguous():
        raise ValueError('signal must be contiguous (it is modified in place)')
    m = _f0_2d(mask_frames)
    B, F = m.shape
    if tuple(signal.shape) != (B, F * int(block_size)):
        raise ValueError('signal must be (B, Frame*block_size)')
    with _OnDevice(signal.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_apply_frame_mask(signal.data_ptr(), m.data_ptr(), m.stride(0), m.stride(1),
                                                           B, F, int(block_size), _st))
    return signal


def apply_volume_mask_(signal, volume_frames, threshold_db=-60.0, block_size=512):
    """The callers' whole silence-mask epilogue in one in-place pass (main.py:112-116,159; gui.py:108-112,127):
    `mask = volume > 10**(threshold_db/20)`, padded by 4 frames with its edge values, 9-frame running maximum,
    `signal *= upsample(mask, block_size)`.  signal (B,T) fp32 contiguous; volume_frames (B,F) or (B,F,1)."""
    signal = _need_cuda_f32(signal, 'signal')
    if not signal.is_contiguous():
        raise ValueError('signal must be contiguous (it is modified in place)')
    v = _f0_2d(volume_frames)
    B, F = v.shape
    if tuple(signal.shape) != (B, F * int(block_size)):
        raise ValueError('signal must be (B, Frame*block_size)')
    with _OnDevice(signal.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_apply_volume_mask(signal.data_ptr(), v.data_ptr(), v.stride(0), v.stride(1),
                                                            10.0 ** (float(threshold_db) / 20.0), B, F, int(block_size), _st))
    return signal


def last_launch_count():
    return _cabi.lib().ddsp_b200_last_launch_count()


# ------------------------------------------------------------------------------------------------
# fused elementwise stages of the control network (not on the synthesizer path)
def performer_features(dash, x, heads, is_query, eps=1e-4):
    """FAVOR+ softmax-kernel features (pcmer.py:124-160).  dash (B,N,H*M) or (B,N,H,M) =
    (64^-0.25 * x) @ projection^T, x (B,N,H*64) -> (B,H,N,M)."""
    dash = _need_cuda_f32(dash, 'dash').contiguous()
    x = _need_cuda_f32(x, 'x').contiguous()
    B, N = x.shape[0], x.shape[1]
    H = int(heads)
    if x.numel() != B * N * H * 64 or dash.numel() % (B * N * H):
        raise ValueError('x must be (B, N, heads*64) and dash (B, N, heads, M)')
    M = dash.numel() // (B * N * H)
    out = torch.empty((B, H, N, M), dtype=torch.float32, device=x.device)
    with _OnDevice(x.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_performer_features(dash.data_ptr(), x.data_ptr(), B, N, H, M, int(bool(is_query)),
                                                             float(eps), out.data_ptr(), _st))
    return out


def performer_project_features(x, projection, heads, is_query, eps=1e-4, x_bias=None):
    """`performer_features` with the random-feature projection fused: x (B,N,H*64), projection (M,64)
    -> (B,H,N,M); the (B,N,H,M) projections never touch HBM.  `x_bias` (H*64): bias of the Linear that
    produced x, added on load (lets the GEMM run without its separate bias epilogue)."""
    x = _need_cuda_f32(x, 'x').contiguous()
    if x_bias is not None:
        x_bias = _need_cuda_f32(x_bias, 'x_bias').contiguous()
        if x_bias.numel() != int(heads) * 64:
            raise ValueError('x_bias must have heads*64 entries')
    projection = _need_cuda_f32(projection, 'projection').contiguous()
    B, N = x.shape[0], x.shape[1]
    H = int(heads)
    if x.numel() != B * N * H * 64 or projection.dim() != 2 or projection.shape[1] != 64:
        raise ValueError('x must be (B, N, heads*64) and projection (M, 64)')
    M = projection.shape[0]
    out = torch.empty((B, H, N, M), dtype=torch.float32, device=x.device)
    with _OnDevice(x.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_performer_project_features(x.data_ptr(), _ptr(x_bias), projection.data_ptr(), B, N, H, M,
                                                                     int(bool(is_query)), float(eps), out.data_ptr(), _st))
    return out


def glu_dwconv_silu(u, weight, bias, u_bias=None):
    """GLU -> depthwise Conv1d(k=31, 'same') -> SiLU in channels-last layout (pcmer.py:53-55).
    u (B,T,2C), weight (C,1,31) or (C,31), bias (C) -> (B,T,C).  `u_bias` (2C): bias of the pointwise
    conv that produced u, added on load."""
    u = _need_cuda_f32(u, 'u').contiguous()
    B, T, C2 = u.shape
    C = C2 // 2
    if u_bias is not None:
        u_bias = _need_cuda_f32(u_bias, 'u_bias').contiguous()
        if u_bias.numel() != C2:
            raise ValueError('u_bias must have 2C entries')
    weight = _need_cuda_f32(weight, 'weight').reshape(C, -1).contiguous()
    if weight.shape[1] != 31:
        raise ValueError('depthwise kernel size must be 31')
    bias = _need_cuda_f32(bias, 'bias').contiguous()
    out = torch.empty((B, T, C), dtype=torch.float32, device=u.device)
    with _OnDevice(u.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_glu_dwconv_silu(u.data_ptr(), _ptr(u_bias), weight.data_ptr(), bias.data_ptr(), B, T, C,
                                                          out.data_ptr(), _st))
    return out


def linear(x, weight, bias=None, residual=None, out=None):
    """`F.linear(x, weight, bias) (+ residual)` on the tensor cores with fp32-faithful 3xTF32 accumulation
    (csrc/gemm_tc.cuh; replaces the nn.Linear / 1x1 Conv1d calls of ddsp/unit2control.py:56-62 and
    ddsp/pcmer.py:41-63, :191-251).  x (..., K) whose leading dimensions collapse to rows with one stride,
    weight (N, K) with contiguous rows; `residual` (..., N) is added in the epilogue; `out` may be given
    (row stride a multiple of 4 floats for 128-bit stores) and may alias `residual`."""
    x = _need_cuda_f32(x, 'x')
    weight = _need_cuda_f32(weight, 'weight')
    N, K = weight.shape
    if x.shape[-1] != K:
        raise ValueError(f'linear: x has {x.shape[-1]} features, weight expects {K}')
    lead = x.shape[:-1]
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1 or (x2.stride(0) & 3) or (x2.data_ptr() & 15):
        x2 = x2.contiguous()
    if weight.stride(1) != 1 or (weight.stride(0) & 3) or (weight.data_ptr() & 15):
        weight = weight.contiguous()
    M = x2.shape[0]
    if K & 3:
        raise ValueError('linear: the reduction length must be a multiple of 4 (16-byte rows for TMA)')
    if out is None:
        out = torch.empty(lead + (N,), dtype=torch.float32, device=x.device)
    o2 = out.view(-1, N) if out.is_contiguous() else out.reshape(-1, out.shape[-1])[:, :N]
    if o2.data_ptr() != out.data_ptr() or o2.stride(1) != 1:
        raise ValueError('linear: `out` must have contiguous rows')
    r2, ldr = None, 0
    if residual is not None:
        r2 = _need_cuda_f32(residual, 'residual').reshape(-1, N)
        if r2.stride(1) != 1:
            r2 = r2.contiguous()
        ldr = r2.stride(0)
    if bias is not None:
        bias = _need_cuda_f32(bias, 'bias').contiguous()
    with _OnDevice(x.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_linear_tf32x3(x2.data_ptr(), x2.stride(0), weight.data_ptr(), weight.stride(0),
                                                        _ptr(bias), _ptr(r2), ldr, o2.data_ptr(), o2.stride(0), M, N, K, _st))
    return out


def split_tf32(w):
    """(hi, lo) with hi = w rounded to TF32 (the kernel's rounding: add half a TF32 ulp to the bit pattern, clear the low
    13 bits) and lo = w - hi (exact): static weights are split once here instead of in shared memory on every tile."""
    w = _need_cuda_f32(w, 'weight').contiguous()
    hi = ((w.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    return hi, w - hi


def linear_ex(x, weight, bias=None, residual=None, out=None, weight_lo=None, ln=None, ln_out=None):
    """`linear` through the second-generation GEMM (csrc/gemm_attn.cuh): TMA-store epilogue, optional pre-split weight
    (`weight` = hi, `weight_lo` = lo from `split_tf32`) and optional fused LayerNorm: `ln = (gamma, beta, eps)` makes the
    call return `(out, layer_norm(out))` -- the normalised copy is produced while the row is still in tensor memory
    (pcmer.py:25,44: every LayerNorm of PCmer follows a residual GEMM)."""
    x = _need_cuda_f32(x, 'x')
    weight = _need_cuda_f32(weight, 'weight')
    N, K = weight.shape
    if x.shape[-1] != K:
        raise ValueError(f'linear: x has {x.shape[-1]} features, weight expects {K}')
    lead = x.shape[:-1]
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1 or (x2.stride(0) & 3) or (x2.data_ptr() & 15):
        x2 = x2.contiguous()
    if weight.stride(1) != 1 or (weight.stride(0) & 3) or (weight.data_ptr() & 15):
        if weight_lo is not None:
            raise ValueError('linear_ex: a pre-split weight must be contiguous')
        weight = weight.contiguous()
    if weight_lo is not None and (weight_lo.shape != weight.shape or weight_lo.stride() != weight.stride()):
        raise ValueError('linear_ex: weight_lo must match weight')
    M = x2.shape[0]
    if K & 3:
        raise ValueError('linear: the reduction length must be a multiple of 4 (16-byte rows for TMA)')
    if out is None:
        out = torch.empty(lead + (N,), dtype=torch.float32, device=x.device)
    o2 = out.view(-1, N) if out.is_contiguous() else out.reshape(-1, out.shape[-1])[:, :N]
    if o2.data_ptr() != out.data_ptr() or o2.stride(1) != 1:
        raise ValueError('linear: `out` must have contiguous rows')
    r2, ldr = None, 0
    if residual is not None:
        r2 = _need_cuda_f32(residual, 'residual').reshape(-1, N)
        if r2.stride(1) != 1:
            r2 = r2.contiguous()
        ldr = r2.stride(0)
    if bias is not None:
        bias = _need_cuda_f32(bias, 'bias').contiguous()
    g = b_ = None
    eps = 0.0
    l2 = None
    if ln is not None:
        g, b_, eps = ln
        g = _need_cuda_f32(g, 'ln gamma').contiguous()
        b_ = _need_cuda_f32(b_, 'ln beta').contiguous()
        if ln_out is None:
            ln_out = torch.empty(lead + (N,), dtype=torch.float32, device=x.device)
        l2 = ln_out.view(-1, N)
    with _OnDevice(x.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_linear_tf32x3_ex(
            x2.data_ptr(), x2.stride(0), weight.data_ptr(), _ptr(weight_lo), weight.stride(0), _ptr(bias), _ptr(r2), ldr,
            o2.data_ptr(), o2.stride(0), _ptr(g), _ptr(b_), float(eps), _ptr(l2), l2.stride(0) if l2 is not None else 0,
            M, N, K, _st))
    return (out, ln_out) if ln is not None else out


def glu_interleave(weight, bias=None):
    """Row order the GLU-fused GEMM expects (include/ddsp_b200.h, ddsp_b200_linear_glu): per 256-row tile t the 128 value
    channels 128 t.. followed by their 128 gate channels N/2 + 128 t..  weight (N, K) [, bias (N)] -> permuted copies."""
    N = weight.shape[0]
    if N % 256:
        raise ValueError('glu_interleave: N must be a multiple of 256')
    half = N // 2
    idx = torch.arange(N, device=weight.device).view(N // 256, 2, 128)
    perm = (idx[:, 0] % 128 + 128 * torch.arange(N // 256, device=weight.device).view(-1, 1))
    perm = torch.stack([perm, perm + half], dim=1).reshape(-1)
    w = weight.reshape(N, -1)[perm].contiguous()
    return (w, bias[perm].contiguous()) if bias is not None else w


def linear_glu(x, weight_il, bias_il=None, weight_lo=None):
    """GLU(F.linear(x, W, b)) (pcmer.py:52-53 in channels-last layout) with the gating done in the GEMM epilogue:
    `weight_il` / `bias_il` / `weight_lo` in the interleaved row order of `glu_interleave`.  x (..., K) -> (..., N/2)."""
    x = _need_cuda_f32(x, 'x')
    N, K = weight_il.shape
    lead = x.shape[:-1]
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1 or (x2.stride(0) & 3) or (x2.data_ptr() & 15):
        x2 = x2.contiguous()
    out = torch.empty(lead + (N // 2,), dtype=torch.float32, device=x.device)
    with _OnDevice(x.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_linear_glu(x2.data_ptr(), x2.stride(0), weight_il.data_ptr(), _ptr(weight_lo),
                                                     weight_il.stride(0), _ptr(bias_il), out.data_ptr(), N // 2, x2.shape[0], N, K, _st))
    return out


def dwconv_silu(g, weight, bias):
    """Depthwise Conv1d(k=31, 'same') -> SiLU on a channels-last tensor g (B,T,C) (pcmer.py:54-55)."""
    g = _need_cuda_f32(g, 'g').contiguous()
    B, T, C = g.shape
    weight = _need_cuda_f32(weight, 'weight').reshape(C, -1).contiguous()
    if weight.shape[1] != 31:
        raise ValueError('depthwise kernel size must be 31')
    bias = _need_cuda_f32(bias, 'bias').contiguous()
    out = torch.empty_like(g)
    with _OnDevice(g.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_dwconv_silu(g.data_ptr(), weight.data_ptr(), bias.data_ptr(), B, T, C, out.data_ptr(), _st))
    return out


_FAVOR_FEATURES, _FAVOR_PAD, _FAVOR_VT_ROWS = 266, 272, 80
_favor_ws = {}


def favor_workspace(B, H, F, device):
    """Buffers of the tensor-core attention path, allocated once per (B, H, F, device) and reused by every layer and
    call: q, k (B,H,F,64); vt (B,H,80,Fp) with its row of ones; q' (B*H,F,272); k'^T (B*H,272,Fp) zero-initialised
    (its pad rows / columns are never written); ctxT (B*H,80,272)."""
    key = (B, H, F, device.index if device.index is not None else torch.cuda.current_device())
    ws = _favor_ws.get(key)
    if ws is None:
        Fp = (F + 3) // 4 * 4
        Z = B * H
        f32 = dict(dtype=torch.float32, device=device)
        vt = torch.zeros((B, H, _FAVOR_VT_ROWS, Fp), **f32)
        vt[:, :, 64, :F] = 1.0
        ws = {'Fp': Fp, 'q': torch.empty((B, H, F, 64), **f32), 'k': torch.empty((B, H, F, 64), **f32), 'vt': vt,
              'qf': torch.empty((Z, F, _FAVOR_PAD), **f32), 'kt': torch.zeros((Z, _FAVOR_PAD, Fp), **f32),
              'ctx': torch.empty((Z, _FAVOR_VT_ROWS, _FAVOR_PAD), **f32), 'ctx_lo': torch.empty((Z, _FAVOR_VT_ROWS, _FAVOR_PAD), **f32)}
        if len(_favor_ws) >= 4:
            _favor_ws.clear()
        _favor_ws[key] = ws
    return ws


def favor_attention(x, w_qkv, w_qkv_lo, b_qkv, proj_scaled, heads, eps=1e-4):
    """Non-causal Performer self-attention (pcmer.py:191-251 up to, not including, `to_out`) as five tensor-core
    launches: merged q|k|v projection with head-split / transposed stores, the FAVOR+ feature GEMM for q and for k,
    the context GEMM and the normalised output GEMM.  x (B,F,C) layer-normed input; w_qkv (3*H*64, C) = [W_q; W_k; W_v]
    (optionally pre-split: w_qkv = hi, w_qkv_lo = lo), b_qkv (3*H*64); proj_scaled = 64^-0.25 * projection_matrix
    (266, 64).  Returns the head-merged attention output (B, F, H*64)."""
    x = _need_cuda_f32(x, 'x')
    B, F, Cc = x.shape
    H = int(heads)
    if proj_scaled.shape != (_FAVOR_FEATURES, 64) or w_qkv.shape != (3 * H * 64, Cc):
        raise ValueError('favor_attention: dim_head 64 / 266 features / merged (3*H*64, C) weight expected')
    x2 = x.reshape(B * F, Cc)
    if x2.stride(1) != 1 or (x2.stride(0) & 3) or (x2.data_ptr() & 15):
        x2 = x2.contiguous()
    ws = favor_workspace(B, H, F, x.device)
    Fp, Z = ws['Fp'], B * H
    out = torch.empty((B, F, H * 64), dtype=torch.float32, device=x.device)
    L = _cabi.lib()
    launches = 0
    with _OnDevice(x.device) as _st:
        _cabi.check(L.ddsp_b200_qkv_heads(x2.data_ptr(), x2.stride(0), w_qkv.data_ptr(), _ptr(w_qkv_lo), w_qkv.stride(0),
                                          _ptr(b_qkv), ws['q'].data_ptr(), ws['k'].data_ptr(), ws['vt'].data_ptr(), 0, B, F, Fp, H, Cc, _st))
        _cabi.check(L.ddsp_b200_favor_features(ws['q'].data_ptr(), proj_scaled.data_ptr(), _FAVOR_FEATURES, 1, float(eps),
                                               ws['qf'].data_ptr(), Z, F, Fp, _st))
        _cabi.check(L.ddsp_b200_favor_features(ws['k'].data_ptr(), proj_scaled.data_ptr(), _FAVOR_FEATURES, 0, float(eps),
                                               ws['kt'].data_ptr(), Z, F, Fp, _st))
        _cabi.check(L.ddsp_b200_favor_context(ws['vt'].data_ptr(), 0, ws['kt'].data_ptr(), ws['ctx'].data_ptr(), ws['ctx_lo'].data_ptr(), Z, Fp, _st))
        _cabi.check(L.ddsp_b200_favor_output(ws['qf'].data_ptr(), ws['ctx'].data_ptr(), ws['ctx_lo'].data_ptr(), out.data_ptr(), B, H, F, _st))
    return out


def tc_microbench(n, k, block_n, virtual_tiles, device='cuda'):
    """Launch the tensor-pipe microbenchmark (operands resident in L2, nothing stored); returns nothing --
    time it with CUDA events."""
    a = torch.randn(128, k, device=device)
    w = torch.randn(n, k, device=device)
    c = torch.empty(128, n, device=device)
    with _OnDevice(a.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_tc_microbench(a.data_ptr(), w.data_ptr(), c.data_ptr(), n, k, block_n, virtual_tiles, _st))


def embed_sum(x, f0, phase, volume, f0_embed, phase_embed, volume_embed, spk_rows):
    """Input embedding sum of Unit2Control.forward (unit2control.py:80-95) in one kernel.
    x (B,N,C) any strides; f0 (B,N,1) or (B,N); phase, volume (B,N); *_embed: nn.Linear(1,C);
    spk_rows (1,C), (B,C) or (B,1,C).  Returns a contiguous (B,N,C) tensor."""
    x = _need_cuda_f32(x, 'x')
    B, N, Cc = x.shape
    f0 = _f0_2d(f0)
    phase = _need_cuda_f32(phase, 'phase').reshape(B, N)
    volume = _need_cuda_f32(volume, 'volume').reshape(B, N)
    spk = _need_cuda_f32(spk_rows, 'spk_rows').reshape(-1, Cc).contiguous()
    if spk.shape[0] not in (1, B):
        raise ValueError('spk_rows must have 1 or B rows')
    ws = []
    for lin in (f0_embed, phase_embed, volume_embed):
        ws += [lin.weight.detach().reshape(-1).contiguous(), lin.bias.detach().contiguous()]
    out = torch.empty((B, N, Cc), dtype=torch.float32, device=x.device)
    with _OnDevice(x.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_embed_sum(
            x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), f0.data_ptr(), f0.stride(0), f0.stride(1),
            phase.data_ptr(), phase.stride(0), phase.stride(1), volume.data_ptr(), volume.stride(0), volume.stride(1),
            ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr(), ws[3].data_ptr(), ws[4].data_ptr(), ws[5].data_ptr(),
            spk.data_ptr(), 0 if spk.shape[0] == 1 else Cc, B, N, Cc, out.data_ptr(), _st))
    return out


def conv3_frames(xp, weight_hi, bias, out_rows, weight_lo=None):
    """Conv1d(C_in, C_out, 3, padding='same') over channels-last frames on the tensor cores (unit2control.py:40,43).
    xp (B, N+2, C_in): frames with one zero frame in front of and behind every clip (`pad_frames`); weight_hi [/ weight_lo]:
    the (C_out, 3*C_in) weight W'[o, t*C_in + c] = W[o, c, t] (`conv3_weight`, split by `split_tf32`); out_rows: the
    (B*(N+2) - 2, C_out) rows that receive the result -- row b*(N+2) + n is frame n of clip b (the two rows per clip
    boundary are garbage).  The A operand is xp itself read with row stride C_in and K = 3*C_in (overlapping rows)."""
    B, Np2, Cin = xp.shape
    M = B * Np2 - 2
    a = xp.as_strided((M, 3 * Cin), (Cin, 1))
    return linear_ex(a, weight_hi, bias, out=out_rows, weight_lo=weight_lo)


def conv3_weight(weight):
    """Conv1d weight (C_out, C_in, 3) -> (C_out, 3*C_in) with the tap as the slow index of a row."""
    return weight.permute(0, 2, 1).reshape(weight.shape[0], -1).contiguous()


def pad_frames(x):
    """(B, N, C) [unit channel stride] -> contiguous (B, N+2, C) with zero frames 0 and N+1."""
    x = _need_cuda_f32(x, 'x')
    B, N, Cc = x.shape
    if x.stride(2) != 1 or (x.stride(0) & 3) or (x.stride(1) & 3) or (x.data_ptr() & 15):
        x = x.contiguous()
    out = torch.empty((B, N + 2, Cc), dtype=torch.float32, device=x.device)
    with _OnDevice(x.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_pad_frames(x.data_ptr(), x.stride(0), x.stride(1), B, N, Cc, out.data_ptr(), _st))
    return out


def groupnorm_leaky_(hp, gamma, beta, eps, groups, slope=0.01):
    """In place on a padded (B, N+2, C) buffer: GroupNorm(groups, C) over the N real frames of every clip, LeakyReLU(slope),
    zeros on the pad frames (unit2control.py:41-42)."""
    B, Np2, Cc = hp.shape
    sums = torch.empty((B, groups, 2), dtype=torch.float64, device=hp.device)
    with _OnDevice(hp.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_groupnorm_leaky(hp.data_ptr(), gamma.detach().contiguous().data_ptr(),
                                                          beta.detach().contiguous().data_ptr(), float(eps), float(slope),
                                                          int(groups), B, Np2 - 2, Cc, sums.data_ptr(), _st))
    return hp


def embed_sum_ln(x, f0, phase, volume, f0_embed, phase_embed, volume_embed, spk_rows, ln):
    """`embed_sum` (unit2control.py:80-95) that also returns the first LayerNorm of PCmer (pcmer.py:25) of the finished
    rows; C = 256.  x (B,N,256) with unit channel stride.  Returns (x_sum, layer_norm(x_sum)), both contiguous."""
    x = _need_cuda_f32(x, 'x')
    B, N, Cc = x.shape
    f0 = _f0_2d(f0)
    phase = _need_cuda_f32(phase, 'phase').reshape(B, N)
    volume = _need_cuda_f32(volume, 'volume').reshape(B, N)
    spk = _need_cuda_f32(spk_rows, 'spk_rows').reshape(-1, Cc).contiguous()
    if spk.shape[0] not in (1, B):
        raise ValueError('spk_rows must have 1 or B rows')
    ws = []
    for lin in (f0_embed, phase_embed, volume_embed):
        ws += [lin.weight.detach().reshape(-1).contiguous(), lin.bias.detach().contiguous()]
    g, b_ = ln.weight.detach().contiguous(), ln.bias.detach().contiguous()
    out = torch.empty((B, N, Cc), dtype=torch.float32, device=x.device)
    out_ln = torch.empty_like(out)
    with _OnDevice(x.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_embed_sum_ln(
            x.data_ptr(), x.stride(0), x.stride(1), f0.data_ptr(), f0.stride(0), f0.stride(1),
            phase.data_ptr(), phase.stride(0), phase.stride(1), volume.data_ptr(), volume.stride(0), volume.stride(1),
            ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr(), ws[3].data_ptr(), ws[4].data_ptr(), ws[5].data_ptr(),
            spk.data_ptr(), 0 if spk.shape[0] == 1 else Cc, g.data_ptr(), b_.data_ptr(), float(ln.eps), B, N, Cc,
            out.data_ptr(), out_ln.data_ptr(), _st))
    return out, out_ln


def performer_attention(q, k, v, projection, heads, q_bias=None, k_bias=None, v_bias=None, eps=1e-4):
    """Non-causal Performer attention after the q/k/v projections as one kernel (blocks of up to 16 frames)
    or three kernels over 8-frame tiles (pcmer.py:69-78,124-160):
    q, k, v (B,N,H*64) -- optionally without the biases of their Linears, passed separately --,
    projection (M,64) -> (B,N,H*64) head-merged attention output (before `to_out`)."""
    q, k, v = (_need_cuda_f32(t, n) for t, n in ((q, 'q'), (k, 'k'), (v, 'v')))
    projection = _need_cuda_f32(projection, 'projection').contiguous()
    B, N, HD = q.shape
    H = int(heads)
    if HD != H * 64 or k.shape != q.shape or v.shape != q.shape or projection.shape[1] != 64:
        raise ValueError('q, k, v must be (B, N, heads*64) and projection (M, 64)')
    # accepted as they are: slices of one merged (B, N, 3*heads*64) projection (common frame stride, clips
    # back to back); anything else is made contiguous
    rs = q.stride(1)
    ok = all(t.stride(2) == 1 and t.stride(1) == rs and t.stride(0) == N * rs and t.data_ptr() % 16 == 0 for t in (q, k, v))
    if not ok or rs % 4 or rs < HD:
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        rs = HD
    biases = []
    for bvec in (q_bias, k_bias, v_bias):
        if bvec is not None:
            bvec = _need_cuda_f32(bvec.detach(), 'bias').contiguous()
            if bvec.numel() != HD:
                raise ValueError('bias must have heads*64 entries')
        biases.append(bvec)
    out = torch.empty((B, N, HD), dtype=torch.float32, device=q.device)
    nbytes = _cabi.lib().ddsp_b200_performer_attention_workspace_bytes(B, N, H)
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=q.device)
    with _OnDevice(q.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_performer_attention(
            q.data_ptr(), k.data_ptr(), v.data_ptr(), rs, _ptr(biases[0]), _ptr(biases[1]), _ptr(biases[2]),
            projection.data_ptr(), B, N, H, projection.shape[0], float(eps), out.data_ptr(), ws.data_ptr(), nbytes, _st))
    return out
