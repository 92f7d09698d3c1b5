"""CPU oracle for the DDSP-SVC synthesizer forward path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`ddsp_b200`) may import this
module; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs do, and there only as the checker / the CPU baseline.

What it is: a numpy restatement of the reference's algorithm for the hot path
(`/root/reference/ddsp/core.py` and `/root/reference/ddsp/vocoder.py:372-550`), function by
function, each citing the reference lines it follows.  It keeps the reference's *rounding
points* that the north star pins bit-exactly (fp32 `upsample` of f0, the fp32 Nyquist mask,
fp64 phase accumulation rounded back to fp32, the fp32 sinc argument) and carries all the
linear filtering (FFT / overlap-add) in the working dtype `wd` (np.float64 for the golden
arbiter, np.float32 to mimic the reference's precision).

Pinning: the five known-answer tests of `ddsp/core.py:54-97` are replayed in
`tests/test_oracle.py`, and every function here is compared against the outputs of the
reference code itself (imported from /root/reference in the build container by
`tests/golden/make_golden.py`, outputs committed under `tests/golden/*.npz`).

The third-party arithmetic underneath the reference is torch 2.11.0 ATen (`F.interpolate`
linear/align_corners, `torch.cumsum`, `torch.fft.rfft/irfft`, `torch.nn.Fold`,
`torch.sinc`); those ops are restated here from their documented semantics.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64


# --------------------------------------------------------------------------------------
# a1  upsample                                                   ddsp/core.py:7-21
# --------------------------------------------------------------------------------------
def upsample(signal: np.ndarray, factor: int) -> np.ndarray:
    """(B, Frame, C) -> (B, Frame*factor, C), linear, hold-last.   core.py:7-21

    `F.interpolate(cat(x, x[-1]), size=F*factor+1, mode='linear', align_corners=True)[:-1]`
    is y[t] = w0*x[m] + w1*x[m+1], m = t // factor, w1 = (t % factor)/factor, w0 = 1 - w1,
    x[F] := x[F-1].  For float32 input the result is bit-identical to torch's kernels, which
    evaluate it as fma(w0, x[m], fl32(w1*x[m+1]))  (verified against torch CPU for
    factor=512; torch's CUDA kernel is the same expression under nvcc's FMA contraction).
    For float64 input the plain expression is evaluated in float64.
    """
    signal = np.asarray(signal)
    B, Fr, C = signal.shape
    factor = int(factor)
    ext = np.concatenate([signal, signal[:, -1:, :]], axis=1)
    t = np.arange(Fr * factor)
    m = t // factor
    x0 = ext[:, m, :]
    x1 = ext[:, m + 1, :]
    if signal.dtype == np.float32:
        # rwidth = Fr / (Fr*factor) in fp32 is exactly 1/factor when factor is a power of two
        # (the only case on the path, block_size=512); keep the general form otherwise.
        rw = F32(Fr) / F32(Fr * factor)
        w1r = (rw * t.astype(F32)).astype(F32)
        # guard against fp32 rounding pushing the source index across an integer
        mm = np.minimum(w1r.astype(np.int64), Fr)
        if not np.array_equal(mm, m):
            x0 = ext[:, mm, :]
            x1 = ext[:, np.minimum(mm + 1, Fr), :]
        w1 = (w1r - mm.astype(F32)).astype(F32)
        w0 = (F32(1) - w1).astype(F32)
        p1 = (w1[None, :, None] * x1).astype(F32)                      # fl32(w1*x1)
        y = (w0[None, :, None].astype(F64) * x0.astype(F64) + p1.astype(F64)).astype(F32)
        return y
    w1 = (t % factor) / factor
    w0 = 1.0 - w1
    return w0[None, :, None] * x0 + w1[None, :, None] * x1


# --------------------------------------------------------------------------------------
# a3  remove_above_fmax                                           ddsp/core.py:24-28
# --------------------------------------------------------------------------------------
def nyquist_mask(pitch: np.ndarray, n_harm: int, fmax, level_start: int = 1) -> np.ndarray:
    """The factor `(pitch*k < fmax).float() + 1e-7` of core.py:26-27, in fp32: values are
    exactly fl32(1+1e-7) = 1.00000012 (kept) or fl32(1e-7) (masked)."""
    pitch = np.asarray(pitch, dtype=F32)
    k = np.arange(level_start, n_harm + level_start).astype(F32)
    pitches = (pitch * k).astype(F32)                                 # fp32 multiply
    aa = (pitches < F32(fmax)).astype(F32) + F32(1e-7)
    return aa.astype(F32)


def remove_above_fmax(amplitudes: np.ndarray, pitch: np.ndarray, fmax, level_start: int = 1):
    """core.py:24-28.  amplitudes (B,F,K), pitch (B,F,1)."""
    n_harm = amplitudes.shape[-1]
    aa = nyquist_mask(pitch, n_harm, fmax, level_start)
    return (amplitudes * aa.astype(amplitudes.dtype)).astype(amplitudes.dtype)


# --------------------------------------------------------------------------------------
# a2  fo_to_rot                                                   ddsp/core.py:31-51
# --------------------------------------------------------------------------------------
def fo_to_rot(fo: np.ndarray, sr, initial_phase=None, precise: bool = False, exact_cumsum: bool = False) -> np.ndarray:
    """core.py:31-51: cumsum(fo/sr) (+init/2/pi), wrap with round-half-even, cast back.

    `exact_cumsum=True` accumulates in extended precision instead of replaying the sequential
    fp64 loop of torch's CPU cumsum.  Over 5 minutes (13 M terms near 1e5 rotations) the sequential
    loop drifts by ~1e-5 rotations of correlated rounding error [measured here], whereas a
    block-parallel scan (torch's CUDA cumsum, and this repo's kernels) stays within 1e-9 of the exact
    sum; the long-form tests use this switch so that the arbiter is the mathematically exact phase."""
    fo = np.asarray(fo)
    _fo = fo.astype(F64) if precise else fo                           # :40
    if exact_cumsum:
        rot = np.cumsum(_fo.astype(np.longdouble) / np.longdouble(sr), axis=1)
        if initial_phase is not None:
            rot = rot + np.asarray(initial_phase).astype(np.longdouble)[:, None] / 2 / np.longdouble(np.pi)
        rot = rot - np.rint(rot)
        return rot.astype(fo.dtype)
    rot = np.cumsum(_fo / _fo.dtype.type(sr), axis=1, dtype=_fo.dtype)  # :43
    if initial_phase is not None:                                      # :44-45
        ip = np.asarray(initial_phase).astype(rot.dtype)
        rot = rot + ip[:, None] / rot.dtype.type(2) / rot.dtype.type(np.pi)
    rot = rot - np.rint(rot)                                           # :46 (half-to-even)
    return rot.astype(fo.dtype)                                        # :49


# --------------------------------------------------------------------------------------
# a8-a10  impulse responses                                       ddsp/core.py:242-328
# --------------------------------------------------------------------------------------
def hann_window_periodic(n: int, dtype=F64) -> np.ndarray:
    """torch.hann_window(n) (periodic=True): 0.5 - 0.5 cos(2 pi i / n)."""
    i = np.arange(n, dtype=F64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * i / n)).astype(dtype)


def bartlett_window_periodic(n: int, dtype=F64) -> np.ndarray:
    """torch.bartlett_window(n) (periodic=True): 1 - |2 i / n - 1|."""
    i = np.arange(n, dtype=F64)
    return (1.0 - np.abs(2.0 * i / n - 1.0)).astype(dtype)


def _irfft_hermitian(magnitudes: np.ndarray) -> np.ndarray:
    """torch.fft.irfft(x) with default n = 2*(n_mag-1): imaginary parts of the DC and
    Nyquist bins are ignored (C2R semantics)."""
    n = 2 * (magnitudes.shape[-1] - 1)
    out = np.fft.irfft(magnitudes, n=n, axis=-1)
    return out


def frequency_impulse_response(magnitudes, hann_window=True, half_width_frames=None, wd=F64):
    """core.py:306-328 (+ :242-289 static window, :292-303 dynamic window).

    magnitudes complex (B,F,n_mag) -> causal-form IR (B,F,L), L = 2(n_mag-1).
    """
    mags = np.asarray(magnitudes)
    ir = _irfft_hermitian(mags.astype(np.complex128 if wd == F64 else np.complex64)).astype(wd)  # :316
    L = ir.shape[-1]
    if hann_window:
        if half_width_frames is None:
            # :262-287 with window_size=0, causal=False: ir*roll(hann,L//2) then roll(L//2)
            win = np.roll(hann_window_periodic(L, F32), L // 2).astype(wd)
            ir = np.roll(ir * win, L // 2, axis=-1)
        else:
            # :296-301 -- arange/hw in the IR dtype, only x>1 zeroed (-> weight 1), x<-1 kept
            hw = np.asarray(half_width_frames).astype(wd)
            x = np.arange(-(L // 2), (L + 1) // 2).astype(wd) / hw
            x = np.where(x > 1, wd(0), x)
            win = (1 + np.cos(wd(np.pi) * x)) / 2
            ir = np.roll(ir, L // 2, axis=-1) * win.astype(wd)
    else:
        ir = np.roll(ir, L // 2, axis=-1)                               # :326
    return ir.astype(wd)


# --------------------------------------------------------------------------------------
# a11-a12  _fft_convolve                                           ddsp/core.py:128-239
# --------------------------------------------------------------------------------------
def get_fft_size(frame_size: int, ir_size: int, power_of_2: bool = True) -> int:
    """core.py:128-144."""
    n = ir_size + frame_size - 1
    return int(2 ** np.ceil(np.log2(n))) if power_of_2 else n


def fft_convolve(audio: np.ndarray, impulse_response: np.ndarray, wd=F64) -> np.ndarray:
    """core.py:185-239: 50%-overlap Bartlett framing, per-frame FFT convolution with that
    frame's IR (last IR repeated), overlap-add, crop with L//2 delay compensation."""
    audio = np.asarray(audio).astype(wd)
    ir = np.asarray(impulse_response).astype(wd)
    if ir.ndim == 2:
        ir = ir[:, None, :]
    Bi, n_ir_frames, L = ir.shape
    B, T = audio.shape
    if B != Bi:                                                         # :212-213
        raise ValueError(f'Batch size of audio ({B}) and impulse response ({Bi}) must be the same.')
    hop = int(T / n_ir_frames)                                          # :216
    frame = 2 * hop
    padded = np.pad(audio, ((0, 0), (hop, hop)))                        # :218
    n_frames = (padded.shape[1] - frame) // hop + 1
    idx = np.arange(frame)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = padded[:, idx] * bartlett_window_periodic(frame, F32).astype(wd)   # :221-222
    fft_size = get_fft_size(frame, L, power_of_2=False)                 # :226
    ir_ext = np.concatenate([ir, ir[:, -1:, :]], axis=1)                # :228
    if ir_ext.shape[1] != n_frames:
        raise ValueError('frame count mismatch between audio frames and IR frames')
    spec = np.fft.rfft(frames, fft_size, axis=-1) * np.fft.rfft(ir_ext, fft_size, axis=-1)
    out_frames = np.fft.irfft(spec, fft_size, axis=-1).astype(wd)       # :230
    total = (n_frames - 1) * hop + fft_size                             # :234
    ola = np.zeros((B, total), dtype=wd)
    for m in range(n_frames):                                           # Fold == overlap-add
        ola[:, m * hop:m * hop + fft_size] += out_frames[:, m]
    ola = ola[:, hop:]                                                  # :238
    start = L // 2                                                      # :177
    return ola[:, start:start + T]                                      # :182 (padding='same')


def frequency_filter(audio, magnitudes, hann_window=True, half_width_frames=None, wd=F64):
    """core.py:331-336."""
    ir = frequency_impulse_response(magnitudes, hann_window, half_width_frames, wd=wd)
    return fft_convolve(audio, ir, wd=wd)


def ltv_fir_direct(audio: np.ndarray, ir: np.ndarray, hop: int) -> np.ndarray:
    """Time-domain statement of what `_fft_convolve` computes:
    y[t] = sum_s x[s] * h_s[t + L//2 - s], h_s = lerp of the frame IRs at input time s.
    O(T*L) -- small cases only; used to cross-check `fft_convolve`."""
    audio = np.asarray(audio, dtype=F64)
    ir = np.asarray(ir, dtype=F64)
    B, T = audio.shape
    _, Fr, L = ir.shape
    ir_ext = np.concatenate([ir, ir[:, -1:, :]], axis=1)
    y = np.zeros((B, T + L), dtype=F64)
    for s in range(T):
        m, lam = s // hop, (s % hop) / hop
        h = (1 - lam) * ir_ext[:, m] + lam * ir_ext[:, m + 1]
        y[:, s:s + L] += audio[:, s:s + 1] * h
    return y[:, L // 2:L // 2 + T]


# --------------------------------------------------------------------------------------
# stage A shared by the three synthesizers            vocoder.py:391-393,449-451,515-517
# --------------------------------------------------------------------------------------
def stage_a(f0_frames: np.ndarray, sr: int, hop: int, initial_phase=None, infer: bool = True,
            exact_cumsum: bool = False):
    """f0 (B,T), rot (B,T), phase_frames (B,F) = 2pi*rot[:, ::hop], all in the dtype of
    `f0_frames`: float32 is the reference's inference path (fp32 rounding points kept);
    float64 input reproduces the reference fed with float64 tensors (the "ref64" arbiter)."""
    f0_frames = np.asarray(f0_frames)
    if f0_frames.dtype != F64:
        f0_frames = f0_frames.astype(F32)
    if f0_frames.ndim == 2:
        f0_frames = f0_frames[..., None]
    dt = f0_frames.dtype.type
    f0 = upsample(f0_frames, hop)[..., 0]
    rot = fo_to_rot(f0, sr, initial_phase, infer, exact_cumsum=exact_cumsum)
    phase_frames = (dt(2 * np.pi) * rot[:, ::hop]).astype(dt)
    return f0, rot, phase_frames


def sinc_f32(x: np.ndarray) -> np.ndarray:
    """torch.sinc on fp32: 1 at 0 else sin(pi x)/(pi x); evaluated from the fp32 argument in
    float64 and returned as float64 (the reference's own fp32 evaluation sits within 1 ulp
    of this)."""
    x = np.asarray(x, dtype=F32).astype(F64)
    return np.sinc(x)


def combtooth(f0: np.ndarray, rot: np.ndarray, sr: int, zero_unvoiced: bool) -> np.ndarray:
    """vocoder.py:459-460 / :539: sinc(fl(fl(sr*rot)/fl(f0+1e-3))), fp32 rounding points
    (plain float64 evaluation when f0/rot are float64)."""
    if f0.dtype == F64:
        c = np.sinc(sr * rot / (f0 + 1e-3))
    else:
        num = (F32(sr) * rot.astype(F32)).astype(F32)
        den = (f0.astype(F32) + F32(1e-3)).astype(F32)
        x = (num / den).astype(F32)
        c = sinc_f32(x)
    if zero_unvoiced:
        c = np.where(f0 <= 0, 0.0, c)
    return c


# --------------------------------------------------------------------------------------
# a6  CombSubFast.forward stage B                                vocoder.py:455-492
# --------------------------------------------------------------------------------------
def combsubfast_forward(harmo_mag, harmo_phase, noise_mag, f0_frames, U, sr=44100, hop=512,
                        initial_phase=None, infer=True, wd=F64, exact_cumsum=False):
    """Returns (signal (B,T), phase_frames (B,F))."""
    f0, rot, phase_frames = stage_a(f0_frames, sr, hop, initial_phase, infer, exact_cumsum)
    B, T = f0.shape
    comb = combtooth(f0, rot, sr, zero_unvoiced=True).astype(wd)         # :459-460
    noise = (np.asarray(U, dtype=F32) * F32(2) - F32(1)).astype(wd)      # :461
    win = np.sqrt(hann_window_periodic(2 * hop, F32)).astype(wd)         # :434 (fp32 buffer)
    n_frames = T // hop + 1
    idx = np.arange(2 * hop)[None, :] + hop * np.arange(n_frames)[:, None]
    cf = np.pad(comb, ((0, 0), (hop, hop)))[:, idx] * win                # :463-468
    nf = np.pad(noise, ((0, 0), (hop, hop)))[:, idx] * win
    cw = np.complex128 if wd == F64 else np.complex64
    hm = np.asarray(harmo_mag).astype(wd)
    hp = np.asarray(harmo_phase).astype(wd)
    nm = np.asarray(noise_mag).astype(wd)
    src = np.exp(hm + 1j * wd(np.pi) * hp).astype(cw)                    # :472
    src = np.concatenate([src, src[:, -1:, :]], axis=1)                  # :473
    nfil = (np.exp(nm) / wd(128)).astype(wd)                             # :475
    nfil = np.concatenate([nfil, nfil[:, -1:, :]], axis=1)               # :476
    spec = np.fft.rfft(cf, 2 * hop, axis=-1) * src + np.fft.rfft(nf, 2 * hop, axis=-1) * nfil  # :479-481
    frames_out = np.fft.irfft(spec, 2 * hop, axis=-1).astype(wd) * win   # :482,486
    ola = np.zeros((B, (n_frames + 1) * hop), dtype=wd)                  # :485-487
    for m in range(n_frames):
        ola[:, m * hop:m * hop + 2 * hop] += frames_out[:, m]
    signal = ola[:, hop:-hop]                                            # :490
    return signal, phase_frames


# --------------------------------------------------------------------------------------
# a5  CombSub.forward (old) stage B                              vocoder.py:521-550
# --------------------------------------------------------------------------------------
def combsub_forward(group_delay, harmonic_magnitude, noise_magnitude, f0_frames, U, sr=44100,
                    hop=512, initial_phase=None, infer=True, wd=F64):
    """Returns (signal, phase_frames, harmonic, noise)."""
    f0, rot, phase_frames = stage_a(f0_frames, sr, hop, initial_phase, infer)
    gd = wd(np.pi) * np.tanh(np.asarray(group_delay).astype(wd))         # :521
    src = np.exp(np.asarray(harmonic_magnitude).astype(wd))              # :522
    npar = np.exp(np.asarray(noise_magnitude).astype(wd)) / wd(128)      # :523
    comb = combtooth(f0, rot, sr, zero_unvoiced=False).astype(wd)        # :539
    allpass = np.exp(1j * np.cumsum(gd, axis=-1))                        # :540
    harmonic = frequency_filter(comb, allpass, hann_window=False, wd=wd)
    f0f = np.asarray(f0_frames, dtype=F32)
    f0f = f0f.reshape(f0f.shape[0], -1, 1)
    hw = (wd(1.5 * sr) / (f0f.astype(wd) + wd(1e-3))).astype(wd)         # :542
    harmonic = frequency_filter(harmonic, src.astype(np.complex128), hann_window=True,
                                half_width_frames=hw, wd=wd)             # :541-542
    noise = (np.asarray(U, dtype=F32) * F32(2) - F32(1)).astype(wd)      # :545
    noise = frequency_filter(noise, npar.astype(np.complex128), hann_window=True, wd=wd)  # :546
    signal = harmonic + noise                                            # :548
    return signal, phase_frames, harmonic, noise


# --------------------------------------------------------------------------------------
# a4  Sins.forward stage B                                       vocoder.py:397-423
# --------------------------------------------------------------------------------------
def sins_forward(amplitudes, group_delay, noise_magnitude, f0_frames, U, sr=44100, hop=512,
                 initial_phase=None, infer=True, max_upsample_dim=32, wd=F64):
    """Returns (signal, phase (B,T) fp32, harmonic, noise)."""
    f0, rot, _ = stage_a(f0_frames, sr, hop, initial_phase, infer)
    f0_frames = np.asarray(f0_frames, dtype=F32)
    phase = (rot.dtype.type(2 * np.pi) * rot).astype(rot.dtype)         # :392
    amp_ctrl = np.asarray(amplitudes)
    # :397 exp()/128 in the control dtype, :402 fp32 Nyquist mask factor
    amp_frames = np.exp(amp_ctrl.astype(wd)) / wd(128)
    n_harm = amp_frames.shape[-1]
    mask = nyquist_mask(f0_frames.reshape(f0_frames.shape[0], -1, 1), n_harm, F32(sr / 2))   # fmax = sr/2 (:402)
    amp_frames = (amp_frames * mask.astype(wd)).astype(wd)
    gd = wd(np.pi) * np.tanh(np.asarray(group_delay).astype(wd))         # :398
    npar = np.exp(np.asarray(noise_magnitude).astype(wd)) / wd(128)      # :399
    B, T = f0.shape
    sinusoids = np.zeros((B, T), dtype=wd)
    level = np.arange(1, n_harm + 1).astype(F32)                         # :404
    for n in range((n_harm - 1) // max_upsample_dim + 1):                # :406-412
        s, e = n * max_upsample_dim, (n + 1) * max_upsample_dim
        phases = (phase[:, :, None] * level[None, None, s:e].astype(phase.dtype)).astype(phase.dtype)   # fp32 product
        amps = upsample(amp_frames[:, :, s:e].astype(F64), hop).astype(wd)
        sinusoids += (amps * np.sin(phases.astype(F64)).astype(wd)).sum(-1)
    allpass = np.exp(1j * np.cumsum(gd, axis=-1))                        # :415
    harmonic = frequency_filter(sinusoids, allpass, hann_window=False, wd=wd)
    noise = (np.asarray(U, dtype=F32) * F32(2) - F32(1)).astype(wd)      # :418
    noise = frequency_filter(noise, npar.astype(np.complex128), hann_window=True, wd=wd)  # :419
    signal = harmonic + noise                                            # :421
    return signal, phase, harmonic, noise


def snr_db(ref: np.ndarray, out: np.ndarray) -> float:
    ref = np.asarray(ref, dtype=F64)
    out = np.asarray(out, dtype=F64)
    num = float(np.sum(ref * ref))
    den = float(np.sum((out - ref) ** 2))
    if den == 0.0:
        return float('inf')
    return 10.0 * np.log10(num / den)
