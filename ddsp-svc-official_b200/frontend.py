"""Host side of the steps immediately downstream of the synthesizer, mirroring the
reference's call sites:

    SincResampler(orig, new, lowpass_filter_width)(audio)   torchaudio.transforms.Resample at enhancer.py:47,69, gui.py:398-401
    MelSTFT(sr, n_mels, n_fft, win, hop, fmin, fmax, mel_basis=...).get_mel(y)       nsf_hifigan/nvSTFT.py:52-121 (STFT.get_mel)
    resample_f0(...)                                                                  enhancer.py:57-63
    EnhancerFrontEnd(...).prepare(audio, sample_rate, f0, hop_size, adaptive_key, silence_front)
                                                                                      enhancer.py:24-66: everything before the
                                                                                      neural vocoder is called
    SolaSplicer(block, crossfade, search).splice(window)                              gui.py:408-426

All of it runs as hand-written kernels (csrc/frontend.cuh) on CUDA float32 tensors; there is no CPU path.  The neural
vocoder itself (NSF-HiFiGAN) stays the reference's PyTorch model (out of scope).
"""
import math

import numpy as np
import torch

from . import _cabi
from .core import _OnDevice, _need_cuda_f32, _f0_2d


def sinc_resample_table(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """The polyphase filter table of torchaudio's `_get_sinc_resample_kernel` (sinc_interp_hann; built in float64 with
    torch's float32 phase offsets, stored as float32), as a host tensor (new, K).  Returns (table, width, orig, new) with
    orig / new divided by their gcd and K = 2 * width + orig."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    if lowpass_filter_width <= 0:
        raise ValueError('Low pass filter width should be positive.')
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, :] / orig
    t = (torch.arange(0, -new, -1)[:, None] / new).to(torch.float64) + idx          # int64 / int -> float32, as torchaudio
    t = t * base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    kernels = kernels * window * (base_freq / orig)
    return kernels.to(torch.float32), width, orig, new


class SincResampler:
    """`torchaudio.transforms.Resample(orig_freq, new_freq, lowpass_filter_width=...)` for CUDA float32 waveforms
    (..., T) -> (..., ceil(new * T / orig))."""

    def __init__(self, orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)
        table, self.width, self.orig, self.new = sinc_resample_table(orig_freq, new_freq, lowpass_filter_width, rolloff)
        self._table_t_host = table.t().contiguous()       # (K, new): consecutive output phases are consecutive words
        self._table_t = {}

    def to(self, device):
        return self

    def __call__(self, waveform):
        waveform = _need_cuda_f32(waveform, 'waveform')
        if self.orig_freq == self.new_freq:
            return waveform
        shape = waveform.shape
        x = waveform.reshape(-1, shape[-1]).contiguous()
        B, T = x.shape
        dev = x.device
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        if key not in self._table_t:
            self._table_t[key] = self._table_t_host.to(dev)
        T_out = int(math.ceil(self.new * T / self.orig))
        y = torch.empty((B, T_out), dtype=torch.float32, device=dev)
        with _OnDevice(dev) as _st:
            _cabi.check(_cabi.lib().ddsp_b200_sinc_resample(x.data_ptr(), B, T, self._table_t[key].data_ptr(), self.orig, self.new,
                                                            self.width, y.data_ptr(), T_out, _st))
        return y.view(shape[:-1] + (T_out,))


def mel_band_ranges(mel_basis):
    """First and one-past-last non-zero bin of every row of a mel filterbank (n_mels, n_bins) -> two int32 host tensors."""
    nz = (mel_basis != 0)
    n_bins = mel_basis.shape[1]
    any_nz = nz.any(dim=1)
    first = torch.where(any_nz, nz.float().argmax(dim=1), torch.zeros(1, dtype=torch.long))
    last = torch.where(any_nz, n_bins - nz.flip(1).float().argmax(dim=1), torch.zeros(1, dtype=torch.long))
    return first.to(torch.int32), last.to(torch.int32)


class MelSTFT:
    """`nsf_hifigan.nvSTFT.STFT` for the synthesizer's output (keyshift = 0, speed = 1, center = False).  The mel
    filterbank is an argument (the reference builds it with librosa, which this package does not depend on):
    `mel_basis` (n_mels, n_fft/2+1) float32."""

    def __init__(self, sr=44100, n_mels=128, n_fft=2048, win_size=2048, hop_length=512, fmin=40, fmax=16000, clip_val=1e-5,
                 mel_basis=None):
        if mel_basis is None:
            raise ValueError('MelSTFT needs the mel filterbank (librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax))')
        self.target_sr, self.n_mels, self.n_fft, self.win_size, self.hop_length = sr, n_mels, n_fft, win_size, hop_length
        self.fmin, self.fmax, self.clip_val = fmin, fmax, clip_val
        basis = torch.as_tensor(np.asarray(mel_basis), dtype=torch.float32).contiguous()
        if tuple(basis.shape) != (n_mels, n_fft // 2 + 1):
            raise ValueError('mel_basis must be (n_mels, n_fft/2 + 1)')
        self._basis_host = basis
        self._ranges_host = mel_band_ranges(basis)
        self._dev = {}

    def get_mel(self, y):
        """y (B,T) in [-1,1] -> log-mel (B, n_mels, n_frames)."""
        y = _need_cuda_f32(y, 'y').contiguous()
        B, T = y.shape
        dev = y.device
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        if key not in self._dev:
            self._dev[key] = (self._basis_host.to(dev), self._ranges_host[0].to(dev), self._ranges_host[1].to(dev))
        basis, b0, b1 = self._dev[key]
        win, hop, n_fft = self.win_size, self.hop_length, self.n_fft
        pad_left = (win - hop) // 2
        pad_right = max((win - hop + 1) // 2, win - T - pad_left)
        n_frames = 1 + (T + pad_left + pad_right - n_fft) // hop
        out = torch.empty((B, self.n_mels, n_frames), dtype=torch.float32, device=dev)
        with _OnDevice(dev) as _st:
            _cabi.check(_cabi.lib().ddsp_b200_mel_spectrogram(y.data_ptr(), B, T, n_fft, win, hop, basis.data_ptr(), b0.data_ptr(),
                                                              b1.data_ptr(), self.n_mels, float(self.clip_val), out.data_ptr(),
                                                              n_frames, _st))
        return out


def resample_f0(f0, hop_size, sample_rate, real_factor, n_frames, enhancer_hop_size, enhancer_sample_rate):
    """enhancer.py:57-63 on the device: f0 (B,n) or (B,n,1) -> (B, n_frames): scaled by real_factor and linearly re-gridded
    (np.interp in double, ends held) from the synthesizer's frame times to the enhancer's."""
    f = _f0_2d(f0)
    B, n = f.shape
    out = torch.empty((B, int(n_frames)), dtype=torch.float32, device=f.device)
    with _OnDevice(f.device) as _st:
        _cabi.check(_cabi.lib().ddsp_b200_interp_frames(f.data_ptr(), f.stride(0), f.stride(1), B, n, float(np.float32(real_factor)),
                                                        hop_size / sample_rate, float(real_factor),
                                                        enhancer_hop_size / enhancer_sample_rate, out.data_ptr(), int(n_frames), _st))
    return out


class EnhancerFrontEnd:
    """Everything `Enhancer.enhance` (enhancer.py:24-66) does before it calls the neural vocoder: silence trimming, the
    adaptive-key sample rate, resampling of the synthesizer output, f0 re-gridding and the mel spectrogram.

        fe = EnhancerFrontEnd(enhancer_sample_rate=44100, enhancer_hop_size=512, mel=MelSTFT(..., mel_basis=...))
        audio_res, mel, f0_res, info = fe.prepare(audio, 44100, f0, 512, adaptive_key=0, silence_front=0)
        enhanced = vocoder(mel, f0_res)                        # the reference's NSF-HiFiGAN, unchanged
        enhanced = fe.finish(enhanced, info)                   # enhancer.py:68-76: resample back, pad the silence
    """

    def __init__(self, enhancer_sample_rate, enhancer_hop_size, mel):
        self.enhancer_sample_rate = int(enhancer_sample_rate)
        self.enhancer_hop_size = int(enhancer_hop_size)
        self.mel = mel
        self.resample_kernel = {}

    def _resampler(self, a, b):
        key = f'{a}_{b}'
        if key not in self.resample_kernel:
            self.resample_kernel[key] = SincResampler(a, b, lowpass_filter_width=128)
        return self.resample_kernel[key]

    def prepare(self, audio, sample_rate, f0, hop_size, adaptive_key=0, silence_front=0):
        audio = _need_cuda_f32(audio, 'audio')
        start_frame = int(silence_front * sample_rate / hop_size)
        real_silence_front = start_frame * hop_size / sample_rate
        audio = audio[:, int(np.round(real_silence_front * sample_rate)):]
        f0 = f0[:, start_frame:, :]
        if adaptive_key == 'auto':
            adaptive_key = 12 * np.log2(float(torch.max(f0)) / 760)       # the one host sync the reference has too (enhancer.py:36)
            adaptive_key = max(0, np.ceil(adaptive_key))
        else:
            adaptive_key = float(adaptive_key)
        adaptive_factor = 2 ** (-adaptive_key / 12)
        adaptive_sample_rate = 100 * int(np.round(self.enhancer_sample_rate / adaptive_factor / 100))
        real_factor = self.enhancer_sample_rate / adaptive_sample_rate
        audio_res = audio if sample_rate == adaptive_sample_rate else self._resampler(sample_rate, adaptive_sample_rate)(audio)
        n_frames = int(audio_res.size(-1) // self.enhancer_hop_size + 1)
        f0_res = resample_f0(f0, hop_size, sample_rate, real_factor, n_frames, self.enhancer_hop_size, self.enhancer_sample_rate)
        mel = self.mel.get_mel(audio_res.contiguous())
        info = {'adaptive_sample_rate': adaptive_sample_rate, 'adaptive_factor': adaptive_factor, 'start_frame': start_frame,
                'real_silence_front': real_silence_front}
        return audio_res, mel, f0_res[:, :mel.size(-1)], info

    def finish(self, enhanced_audio, info):
        if info['adaptive_factor'] != 0:
            enhanced_audio = self._resampler(info['adaptive_sample_rate'], self.enhancer_sample_rate)(enhanced_audio)
        if info['start_frame'] > 0:
            enhanced_audio = torch.nn.functional.pad(
                enhanced_audio, (int(np.round(self.enhancer_sample_rate * info['real_silence_front'])), 0))
        return enhanced_audio, self.enhancer_sample_rate


class SolaSplicer:
    """The GUI's block splice (gui.py:338-345,408-426, without the phase vocoder): keeps the saved tail (`sola_buffer`)
    and the fade windows on the device and splices every new window in one kernel.

        out = splicer.splice(window)      # window: 1-D CUDA tensor of >= block + crossfade + search samples
        splicer.last_shift                # 0-dim int32 device tensor (no host sync)
    """

    def __init__(self, block_frame, crossfade_frame, sola_search_frame, device='cuda'):
        self.block, self.crossfade, self.search = int(block_frame), int(crossfade_frame), int(sola_search_frame)
        C = self.crossfade
        self.fade_in_window = (torch.sin(np.pi * torch.arange(0, 1, 1 / C) / 2) ** 2).to(device=device, dtype=torch.float32)   # gui.py:338-340
        self.fade_out_window = 1 - self.fade_in_window
        self.sola_buffer = torch.zeros(C, dtype=torch.float32, device=device)
        self.last_shift = torch.zeros((), dtype=torch.int32, device=device)

    def splice(self, temp_wav):
        x = _need_cuda_f32(temp_wav, 'temp_wav').contiguous()
        if x.dim() != 1 or x.numel() < self.block + self.crossfade + self.search:
            raise ValueError('the window must be 1-D with at least block + crossfade + search samples')
        out = torch.empty(self.block, dtype=torch.float32, device=x.device)
        with _OnDevice(x.device) as _st:
            _cabi.check(_cabi.lib().ddsp_b200_sola_splice(x.data_ptr(), x.numel(), self.sola_buffer.data_ptr(),
                                                          self.fade_in_window.data_ptr(), self.fade_out_window.data_ptr(), self.block,
                                                          self.crossfade, self.search, out.data_ptr(), self.last_shift.data_ptr(), _st))
        return out
