"""Oracle of SURVEY section 8 row (f4) (oracle/frontend_oracle.py) against the reference itself: golden vectors made by
nsf_hifigan/nvSTFT.py, enhancer.py and the replayed gui.py:408-426 (tests/golden/make_golden_frontend.py), and against
torchaudio's own resampler where it is installed."""
import os

import numpy as np
import pytest

from oracle import frontend_oracle as FO

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'frontend.npz'))


def test_mel_spectrogram_matches_nvstft():
    sr, n_mels, n_fft, win, hop, fmin, fmax = [int(v) for v in G['mel_params']]
    assert np.array_equal(FO.slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax), G['mel_basis'])
    mel = FO.mel_spectrogram(G['mel_audio'], G['mel_basis'], n_fft, win, hop)
    assert mel.shape == G['mel_ref'].shape
    # log of fp32 magnitudes: the reference's own fp32 FFT differs from the fp64 one at the 1e-5 level near the clamp
    assert np.abs(mel - G['mel_ref']).max() < 2e-4


def test_slaney_mel_basis_properties():
    """librosa is not installed: the restated filterbank is checked through its defining properties."""
    b = FO.slaney_mel_basis(44100, 2048, 128, 40, 16000).astype(np.float64)
    freqs = np.linspace(0, 22050, 1025)
    assert b.shape == (128, 1025) and (b >= 0).all()
    assert (b[:, freqs < 40 - 22].sum(1) == 0).all() and (b[:, freqs > 16000 + 22] == 0).all()
    peaks = freqs[b.argmax(1)]
    assert (np.diff(peaks) > 0).all()                       # centre frequencies increase
    nz = [(np.flatnonzero(r)[0], np.flatnonzero(r)[-1]) for r in b]
    assert all(np.all(r[s:e + 1] > 0) for r, (s, e) in zip(b, nz))     # contiguous triangles
    # Slaney normalisation: every triangle integrates to ~1 over frequency (bin spacing sr / n_fft)
    area = b.sum(1) * (44100 / 2048)
    assert np.abs(area - 1).max() < 0.15 and np.abs(area[60:] - 1).max() < 0.02     # narrow low bands: coarse bin sampling


@pytest.mark.parametrize('orig,new,lpw', [(44100, 46700, 128), (44100, 52400, 128), (44100, 16000, 6), (48000, 44100, 128)])
def test_sinc_resampler_matches_torchaudio(orig, new, lpw):
    torch = pytest.importorskip('torch')
    ta = pytest.importorskip('torchaudio')
    r = ta.transforms.Resample(orig, new, lowpass_filter_width=lpw)
    kern, width, o, n = FO.sinc_resample_kernel(orig, new, lpw)
    assert width == r.width and kern.shape == tuple(r.kernel[:, 0].shape)
    assert np.abs(kern - r.kernel[:, 0].numpy()).max() < 1e-7
    x = np.random.default_rng(0).standard_normal((2, 3000)).astype(np.float32)
    ref = r(torch.from_numpy(x)).numpy()
    got = FO.resample(x, orig, new, lpw)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 2e-5


@pytest.mark.parametrize('tag', ['k0', 'k3', 'k5s', 'auto'])
def test_enhancer_inputs_match_enhancer_py(tag):
    key, sil, auto = G[f'enh_{tag}_args']
    audio_res, mel, f0_res, asr, start = FO.enhancer_inputs(
        G[f'enh_{tag}_audio'][None], 44100, G[f'enh_{tag}_f0'][None, :, None], 512, 'auto' if auto else key, float(sil),
        44100, 512, G['mel_basis'])
    assert audio_res.shape == G[f'enh_{tag}_audio_res'].shape
    assert np.abs(audio_res - G[f'enh_{tag}_audio_res']).max() < 2e-5
    ref_f0 = G[f'enh_{tag}_f0_res']
    assert np.abs(f0_res - ref_f0[:, :f0_res.shape[1]]).max() < 1e-3          # Hz; float32 of the same fp64 np.interp
    assert f0_res.shape[1] == mel.shape[-1]


def test_sola_splice_matches_gui():
    block, C, S = [int(v) for v in G['sola_geom']]
    out, new_buf, shift = FO.sola_splice(G['sola_temp_wav'], G['sola_buffer'], G['sola_fade_in'], G['sola_fade_out'], block, C, S)
    assert shift == int(G['sola_shift'][0])
    assert np.abs(out - G['sola_out']).max() < 1e-6
    assert np.abs(new_buf - G['sola_new_buffer']).max() < 1e-6


@pytest.mark.parametrize('orig,new,lpw', [(44100, 46700, 128), (44100, 16000, 6)])
def test_product_resample_table_equals_torchaudio(orig, new, lpw):
    """Host logic of the product (ddsp_b200.frontend.sinc_resample_table) against torchaudio's own table."""
    pytest.importorskip('torch')
    ta = pytest.importorskip('torchaudio')
    from ddsp_b200.frontend import sinc_resample_table
    r = ta.transforms.Resample(orig, new, lowpass_filter_width=lpw)
    table, width, o, n = sinc_resample_table(orig, new, lpw)
    assert width == r.width
    assert np.array_equal(table.numpy(), r.kernel[:, 0].numpy())
