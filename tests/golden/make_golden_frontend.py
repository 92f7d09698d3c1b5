#!/usr/bin/env python
"""Golden vectors for SURVEY section 8 row (f4) -- the enhancer front-end and the GUI SOLA splice -- produced by the
REFERENCE code itself where it is importable here:

  * `nsf_hifigan.nvSTFT.STFT.get_mel`          (nvSTFT.py:65-116; librosa / soundfile are absent: stubbed, the mel basis
                                                comes from oracle.frontend_oracle.slaney_mel_basis)
  * `enhancer.Enhancer.enhance`                (enhancer.py:24-78) with the neural vocoder replaced by a recorder, so
                                                that what it RECEIVES (resampled audio, resampled f0) is captured;
                                                torchaudio (installed) does the resampling exactly as in the reference
  * gui.py:408-426 (SOLA)                      gui.py cannot be imported (sounddevice / PySimpleGUI); its tensor ops are
                                                replayed here line by line with torch on the CPU

    python tests/golden/make_golden_frontend.py        -> tests/golden/frontend.npz
"""
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get('DDSP_REFERENCE', '/root/reference')

from oracle import frontend_oracle as FO      # noqa: E402


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def main():
    stub('librosa')
    stub('librosa.util', normalize=lambda x: x)
    stub('librosa.filters', mel=lambda sr, n_fft, n_mels, fmin, fmax: FO.slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax))
    stub('soundfile')
    sys.path.insert(0, REF)
    from nsf_hifigan.nvSTFT import STFT
    import enhancer as ref_enh

    rng = np.random.default_rng(2024)
    out = {}
    # ---- mel spectrogram (44.1 kHz NSF-HiFiGAN settings: 128 mels, n_fft = win = 2048, hop 512, 40..16000 Hz)
    sr, n_mels, n_fft, win, hop, fmin, fmax = 44100, 128, 2048, 2048, 512, 40, 16000
    stft = STFT(sr, n_mels, n_fft, win, hop, fmin, fmax)
    t = np.arange(3 * 512 * 11 + 137) / sr
    y = (0.4 * np.sin(2 * np.pi * 220 * t * (1 + 0.2 * t)) + 0.05 * rng.standard_normal(t.shape)).astype(np.float32)
    y2 = (0.3 * rng.standard_normal(t.shape)).astype(np.float32)
    audio = np.stack([y, y2])
    mel = stft.get_mel(torch.from_numpy(audio)).numpy()
    out.update(mel_audio=audio, mel_ref=mel, mel_basis=FO.slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax),
               mel_params=np.array([sr, n_mels, n_fft, win, hop, fmin, fmax]))
    mel64 = stft.get_mel(torch.from_numpy(audio).double()).numpy() if False else None      # (the basis is cached as fp32)

    # ---- Enhancer.enhance: capture what the vocoder receives
    class Recorder:
        def __init__(self):
            self.calls = []

        def sample_rate(self):
            return 44100

        def hop_size(self):
            return 512

        def __call__(self, audio_res, f0_res):
            self.calls.append((audio_res.clone(), f0_res.clone()))
            return audio_res, 44100
    for tag, key, sil in (('k0', 0, 0.0), ('k3', 3, 0.0), ('k5s', 5, 0.12), ('auto', 'auto', 0.0)):
        e = ref_enh.Enhancer.__new__(ref_enh.Enhancer)
        e.device = 'cpu'
        e.enhancer = Recorder()
        e.resample_kernel = {}
        e.enhancer_sample_rate = 44100
        e.enhancer_hop_size = 512
        n_fr = 40
        f0 = (180 + 60 * np.sin(np.arange(n_fr) / 5.0)).astype(np.float32)
        if tag == 'auto':
            f0 = f0 * 5.2                       # peaks above 760 Hz -> adaptive key > 0
        a = (0.3 * np.sin(2 * np.pi * 200 * np.arange(n_fr * 512) / 44100) + 0.02 * rng.standard_normal(n_fr * 512)).astype(np.float32)
        res, sr_o = e.enhance(torch.from_numpy(a)[None], 44100, torch.from_numpy(f0.copy())[None, :, None], 512,
                              adaptive_key=key, silence_front=sil)
        audio_res, f0_res = e.enhancer.calls[0]
        out[f'enh_{tag}_audio'] = a
        out[f'enh_{tag}_f0'] = f0
        out[f'enh_{tag}_audio_res'] = audio_res.numpy()
        out[f'enh_{tag}_f0_res'] = f0_res.numpy()
        out[f'enh_{tag}_out'] = res.numpy()
        out[f'enh_{tag}_args'] = np.array([0.0 if key == 'auto' else float(key), sil, 1.0 if key == 'auto' else 0.0])

    # ---- gui.py:408-426 replayed op by op (block 0.3 s, crossfade 0.04 s, search 0.01 s at 44.1 kHz)
    block, C, S = 13230, 1764, 441
    n = block + C + S
    base = (0.5 * np.sin(2 * np.pi * 147 * np.arange(2 * n) / 44100)).astype(np.float32)
    temp_wav = torch.from_numpy(base[300:300 + n] + 0.01 * rng.standard_normal(n).astype(np.float32))
    sola_buffer = torch.from_numpy(base[300 + 173:300 + 173 + C].copy())
    fade_in = torch.sin(np.pi * torch.arange(0, 1, 1 / C) / 2) ** 2                      # gui.py:338-340
    fade_out = 1 - fade_in
    conv_input = temp_wav[None, None, :C + S]
    cor_nom = F.conv1d(conv_input, sola_buffer[None, None, :])
    cor_den = torch.sqrt(F.conv1d(conv_input ** 2, torch.ones(1, 1, C)) + 1e-8)
    shift = torch.argmax(cor_nom[0, 0] / cor_den[0, 0])
    tw = temp_wav[shift: shift + block + C].clone()
    tw[:C] *= fade_in
    tw[:C] += sola_buffer * fade_out
    out.update(sola_temp_wav=temp_wav.numpy(), sola_buffer=sola_buffer.numpy(), sola_fade_in=fade_in.numpy().astype(np.float32),
               sola_fade_out=fade_out.numpy().astype(np.float32), sola_out=tw[:-C].numpy(), sola_new_buffer=tw[-C:].numpy(),
               sola_shift=np.array([int(shift)]), sola_geom=np.array([block, C, S]))
    np.savez_compressed(os.path.join(HERE, 'frontend.npz'), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == '__main__':
    main()
