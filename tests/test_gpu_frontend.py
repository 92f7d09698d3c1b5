"""Enhancer front-end and SOLA splice kernels (csrc/frontend.cuh) against the golden vectors the
reference produced (tests/golden/frontend.npz: nvSTFT.get_mel, Enhancer.enhance, gui.py:408-426) and against the oracle."""
import os

import numpy as np
import pytest

from tests.gpu_util import HAS_CUDA, dev, torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not HAS_CUDA, reason='needs a CUDA device')]

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'frontend.npz'))


def _mel():
    from ddsp_b200.frontend import MelSTFT
    sr, n_mels, n_fft, win, hop, fmin, fmax = [int(v) for v in G['mel_params']]
    return MelSTFT(sr, n_mels, n_fft, win, hop, fmin, fmax, mel_basis=G['mel_basis'])


def test_mel_matches_nvstft_golden():
    mel = _mel().get_mel(dev(G['mel_audio'])).cpu().numpy()
    assert mel.shape == G['mel_ref'].shape
    assert np.abs(mel - G['mel_ref']).max() < 2e-4


@pytest.mark.parametrize('B,T', [(3, 44100), (1, 512 * 7 + 1), (2, 1900), (1, 700)])
def test_mel_matches_oracle(B, T):
    """Ragged lengths, and clips shorter than the padding (nvSTFT.py:97-101 switches to constant padding)."""
    from oracle import frontend_oracle as FO
    rng = np.random.default_rng(T)
    y = (0.5 * rng.standard_normal((B, T))).astype(np.float32)
    ref = FO.mel_spectrogram(y, G['mel_basis'], 2048, 2048, 512)
    got = _mel().get_mel(dev(y)).cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 2e-4


@pytest.mark.parametrize('orig,new,lpw', [(44100, 46700, 128), (44100, 52400, 128), (44100, 16000, 6), (48000, 44100, 128)])
def test_resampler_matches_oracle(orig, new, lpw):
    from oracle import frontend_oracle as FO
    from ddsp_b200.frontend import SincResampler
    x = np.random.default_rng(orig + new).standard_normal((2, 7001)).astype(np.float32)
    ref = FO.resample(x, orig, new, lpw)
    got = SincResampler(orig, new, lpw)(dev(x)).cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 2e-5


@pytest.mark.parametrize('tag', ['k0', 'k3', 'k5s', 'auto'])
def test_enhancer_front_end_matches_enhancer_py(tag):
    from oracle import frontend_oracle as FO
    from ddsp_b200.frontend import EnhancerFrontEnd
    key, sil, auto = G[f'enh_{tag}_args']
    fe = EnhancerFrontEnd(44100, 512, _mel())
    audio = dev(G[f'enh_{tag}_audio'][None])
    f0 = dev(G[f'enh_{tag}_f0'][None, :, None])
    audio_res, mel, f0_res, info = fe.prepare(audio, 44100, f0, 512, adaptive_key='auto' if auto else key, silence_front=float(sil))
    ref_audio = G[f'enh_{tag}_audio_res']
    assert tuple(audio_res.shape) == ref_audio.shape
    assert np.abs(audio_res.cpu().numpy() - ref_audio).max() < 2e-5
    ref_f0 = G[f'enh_{tag}_f0_res'][:, :f0_res.shape[1]]
    assert np.abs(f0_res.cpu().numpy() - ref_f0).max() < 1e-3
    ref_mel = FO.mel_spectrogram(ref_audio, G['mel_basis'], 2048, 2048, 512)
    assert mel.shape == ref_mel.shape and f0_res.shape[1] == mel.shape[-1]
    # compared in the linear domain: in the bands the up-sampling leaves empty (above 22.05 kHz / factor) the mel energy is
    # the fp32 noise floor of the FFT (~1e-6 of the peak), where the log amplifies differences no fp32 implementation controls
    a, b = np.exp(mel.cpu().numpy().astype(np.float64)), np.exp(ref_mel.astype(np.float64))
    assert (np.abs(a - b) <= 3e-4 * b + 2e-7).all(), np.abs(a - b).max()
    # the way back (enhancer.py:68-76) with the vocoder replaced by identity, as in the golden run
    out, sr_o = fe.finish(audio_res, info)
    ref_out = G[f'enh_{tag}_out']
    assert sr_o == 44100 and tuple(out.shape) == ref_out.shape
    assert np.abs(out.cpu().numpy() - ref_out).max() < 5e-5


def test_sola_splice_matches_gui_golden():
    from ddsp_b200.frontend import SolaSplicer
    block, C, S = [int(v) for v in G['sola_geom']]
    sp = SolaSplicer(block, C, S)
    assert np.array_equal(sp.fade_in_window.cpu().numpy(), G['sola_fade_in'])
    sp.sola_buffer.copy_(dev(G['sola_buffer']))
    out = sp.splice(dev(G['sola_temp_wav']))
    assert int(sp.last_shift.item()) == int(G['sola_shift'][0])
    assert np.array_equal(out.cpu().numpy(), G['sola_out'])
    assert np.array_equal(sp.sola_buffer.cpu().numpy(), G['sola_new_buffer'])


def test_sola_splice_random_blocks_against_oracle():
    from oracle import frontend_oracle as FO
    from ddsp_b200.frontend import SolaSplicer
    rng = np.random.default_rng(5)
    block, C, S = 4096, 600, 150
    sp = SolaSplicer(block, C, S)
    buf = np.zeros(C, np.float32)
    sig = np.sin(2 * np.pi * 97.3 * np.arange(40000) / 44100).astype(np.float32)
    for it in range(4):
        a = 1000 * it + int(rng.integers(0, 200))
        w = sig[a:a + block + C + S] + 0.05 * rng.standard_normal(block + C + S).astype(np.float32)
        ref_out, buf, ref_shift = FO.sola_splice(w, buf, sp.fade_in_window.cpu().numpy(), sp.fade_out_window.cpu().numpy(), block, C, S)
        out = sp.splice(dev(w))
        assert int(sp.last_shift.item()) == ref_shift
        assert np.abs(out.cpu().numpy() - ref_out).max() < 1e-6
        assert np.abs(sp.sola_buffer.cpu().numpy() - buf).max() < 1e-6
