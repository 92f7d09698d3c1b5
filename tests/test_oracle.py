"""Pin the CPU oracle: the reference's own known-answer tests (ddsp/core.py:54-97) and the
golden vectors produced by running the reference code (tests/golden/make_golden.py)."""
import math
import os

import numpy as np
import pytest

from oracle import ddsp_oracle as O
from ddsp_b200.synthetic import make_inputs

F32 = np.float32


# ---- the five KATs of ddsp/core.py:54-97, replayed on the oracle -----------------------
def test_fo_to_rot_dtype():                         # core.py:54-59
    fo = np.array([[1.0, 1.0, 1.0]], dtype=F32)
    assert O.fo_to_rot(fo, 1, precise=False).dtype == fo.dtype
    assert O.fo_to_rot(fo, 1, precise=True).dtype == fo.dtype


def test_fo_to_rot_stablefo():                      # core.py:62-67
    rot = O.fo_to_rot(np.array([[1.0, 1.0, 1.0]], dtype=F32), 4)
    np.testing.assert_allclose(rot, [[+0.25, +0.50, -0.25]], rtol=1e-5, atol=1e-8)


def test_fo_to_rot_fm():                            # core.py:70-76
    rot = O.fo_to_rot(np.array([[1.0, 2.0, 3.0]], dtype=F32), 4)
    np.testing.assert_allclose(rot, [[+0.25, -0.25, -0.50]], rtol=1e-5, atol=1e-8)


def test_fo_to_rot_init_phase():                    # core.py:79-87
    rot = O.fo_to_rot(np.array([[1.0, 1.0, 1.0]], dtype=F32), 4, initial_phase=np.array([math.pi], dtype=F32))
    np.testing.assert_allclose(rot, [[-0.25, 0.0, +0.25]], rtol=1e-5, atol=1e-7)


def test_fo_to_rot_fm_init_batch():                 # core.py:90-97
    fo = np.array([[1.0, 1.0, 1.0], [1.0, 2.0, 3.0]], dtype=F32)
    ip = np.array([math.pi, 0.0], dtype=F32)
    rot = O.fo_to_rot(fo, 4, initial_phase=ip, precise=True)
    np.testing.assert_allclose(rot, [[-0.25, 0.0, +0.25], [+0.25, -0.25, -0.50]], atol=1e-5)


# ---- core-level golden vectors -----------------------------------------------------------
@pytest.fixture(scope='module')
def core(golden_dir):
    return np.load(os.path.join(golden_dir, 'core.npz'))


def test_upsample_bit_exact(core):
    y = O.upsample(core['upsample_x'], 512)
    assert y.dtype == F32
    assert np.array_equal(y[:, ::37, :], core['upsample_y_sub'])
    assert np.array_equal(y.astype(np.float64).sum(axis=1), core['upsample_y_sum'])


def test_rot_precise_bit_exact(core):
    fo = O.upsample(core['upsample_x'][:, :, :1], 512)[..., 0]
    rot = O.fo_to_rot(fo, 44100, None, True)
    # sequential fp64 cumsum on both sides -> identical
    assert np.array_equal(rot[:, ::37], core['rot_precise_sub'])
    rot_ip = O.fo_to_rot(fo, 44100, core['rot_ip'], True)
    np.testing.assert_allclose(rot_ip[:, ::37], core['rot_precise_ip_sub'], atol=2e-7)


def test_nyquist_mask_bit_exact(core):
    out = O.remove_above_fmax(core['mask_amp'], core['mask_pitch'], F32(22050.0))
    assert np.array_equal(out, core['mask_out'])
    m = O.nyquist_mask(core['mask_pitch'], 128, F32(22050.0))
    assert set(np.unique(m).tolist()) == {float(F32(1e-7)), float(F32(1) + F32(1e-7))}
    # f0 = 172.265625: 128*f0 == 22050 exactly -> k=128 masked, k=127 kept
    assert m[0, 0, 127] == F32(1e-7) and m[0, 0, 126] == F32(1) + F32(1e-7)
    # f0 = 800: k=28 is the first masked harmonic (27*800=21600 < 22050 <= 28*800)
    assert m[0, 1, 26] > 0.5 and m[0, 1, 27] < 0.5
    # f0 = 0: nothing masked
    assert np.all(m[0, 3] > 0.5)


@pytest.mark.parametrize('name', ['none', 'hann'])
def test_frequency_filter_static(core, name):
    mags = core['ff_mags'].astype(np.complex128)
    ir = O.frequency_impulse_response(mags, hann_window=(name == 'hann'))
    np.testing.assert_allclose(ir, core[f'ff_{name}_ir'], atol=2e-7)   # fp32 hann window on both sides
    y = O.frequency_filter(core['ff_audio'], mags, hann_window=(name == 'hann'))
    np.testing.assert_allclose(y, core[f'ff_{name}_y'], atol=1e-6)


def test_frequency_filter_dynamic_window(core):
    mags = core['ff_mags2'].astype(np.complex128)
    hw = 1.5 * 44100 / (core['ff_f0f'] + 1e-3)
    ir = O.frequency_impulse_response(mags, True, hw)
    np.testing.assert_allclose(ir, core['ff_dyn_ir'], atol=1e-12)
    y = O.frequency_filter(core['ff_audio'], mags, True, hw)
    np.testing.assert_allclose(y, core['ff_dyn_y'], atol=1e-11)


def test_dynamic_window_quirk():
    """only x>1 is zeroed (-> weight 1), x<-1 keeps the cosine."""
    L = 1022
    mags = np.ones((1, 1, 512), dtype=np.complex128)      # ir = delta -> windowed ir shows w[L//2]
    hw = np.array([[[1.5 * 44100 / (220 + 1e-3)]]])
    x = np.arange(-(L // 2), (L + 1) // 2) / hw[0, 0, 0]
    w = (1 + np.cos(np.pi * np.where(x > 1, 0, x))) / 2
    assert abs(w[0] - 0.793) < 1e-3 and w[511] == 1.0 and np.all(w[511 + 301:] == 1.0) and w[511 + 300] < 2e-5
    ir = O.frequency_impulse_response(mags, True, hw)
    assert abs(ir[0, 0, L // 2] - 1.0) < 1e-12


def test_allpass_filter(core):
    gd = np.pi * np.tanh(core['ff_ph'])
    y = O.frequency_filter(core['ff_audio'], np.exp(1j * np.cumsum(gd, -1)), hann_window=False)
    np.testing.assert_allclose(y, core['ff_ap_y'], atol=1e-11)


def test_fft_convolve_equals_time_domain_ltv_fir(core):
    mags = core['ff_mags'].astype(np.complex128)[:1, :3]
    audio = core['ff_audio'][:1, :3 * 512]
    ir = O.frequency_impulse_response(mags, hann_window=True)
    np.testing.assert_allclose(O.fft_convolve(audio, ir), O.ltv_fir_direct(audio, ir, 512), atol=1e-11)


def test_fft_convolve_batch_mismatch():               # core.py:212-213
    with pytest.raises(ValueError):
        O.fft_convolve(np.zeros((2, 1024)), np.zeros((3, 2, 510)))


def test_windows():
    b = O.bartlett_window_periodic(1024)
    assert b[0] == 0 and b[512] == 1 and abs(b[1023] - 1 / 512) < 1e-15
    np.testing.assert_allclose(b[:512] + b[512:], 1.0)
    h = O.hann_window_periodic(1024)
    np.testing.assert_allclose(h[:512] + h[512:], 1.0, atol=1e-15)    # sqrt(hann)^2 is COLA
    assert O.get_fft_size(1024, 510, False) == 1533 and O.get_fft_size(1024, 1022, False) == 2045


# ---- model-level golden vectors ------------------------------------------------------------
SPLITS = {'combsubfast': (513, 513, 513), 'combsub': (256, 512, 256), 'sins': (128, 256, 256)}


def _load_case(golden_dir, model, tag):
    d = dict(np.load(os.path.join(golden_dir, f'{model}_{tag}.npz')))
    if 'ctrl' not in d:
        inp = make_inputs(int(d['B']), int(d['F']), sum(SPLITS[model]), seed=int(d['seed']),
                          zero_f0_fraction=float(d['zero_f0_fraction']))
        d.update(inp)
    a, b, _ = SPLITS[model]
    d['c0'], d['c1'], d['c2'] = d['ctrl'][..., :a], d['ctrl'][..., a:a + b], d['ctrl'][..., a + b:]
    return d


def _check(out, ref32, tol32=3e-6):
    """oracle (fp64 filtering on the reference's fp32 rounding points) vs the fp32 reference."""
    err = np.abs(out - ref32).max()
    assert err < tol32, err
    assert O.snr_db(ref32, out) > 90.0


def _check64(out64, ref64, tol64=3e-7):
    """oracle fed float64 f0 (no fp32 rounding points) vs the reference run in float64
    (fixtures store ref64 as float32: 6e-8 relative)."""
    err = np.abs(out64 - ref64).max()
    assert err < tol64, err


@pytest.mark.parametrize('tag', ['small', 'odd', 'gui'])
def test_combsubfast_vs_reference(golden_dir, tag):
    d = _load_case(golden_dir, 'combsubfast', tag)
    sig, pf = O.combsubfast_forward(d['c0'], d['c1'], d['c2'], d['f0_frames'], d['U'])
    _check(sig, d['signal32'])
    np.testing.assert_allclose(pf, d['phase32'][..., 0], atol=1e-6)
    sig64, _ = O.combsubfast_forward(d['c0'], d['c1'], d['c2'], d['f0_frames'].astype(np.float64), d['U'])
    _check64(sig64, d['signal64'])


@pytest.mark.parametrize('tag', ['small', 'odd', 'gui'])
def test_combsub_vs_reference(golden_dir, tag):
    d = _load_case(golden_dir, 'combsub', tag)
    sig, pf, harm, noise = O.combsub_forward(d['c0'], d['c1'], d['c2'], d['f0_frames'], d['U'])
    _check(sig, d['signal32'])
    _check(harm, d['harm32'])
    np.testing.assert_allclose(pf, d['phase32'][..., 0], atol=1e-6)
    o64 = O.combsub_forward(d['c0'], d['c1'], d['c2'], d['f0_frames'].astype(np.float64), d['U'])
    _check64(o64[0], d['signal64'])


@pytest.mark.parametrize('tag', ['small', 'odd', 'gui'])
def test_sins_vs_reference(golden_dir, tag):
    d = _load_case(golden_dir, 'sins', tag)
    sig, phase, harm, noise = O.sins_forward(d['c0'], d['c1'], d['c2'], d['f0_frames'], d['U'])
    _check(sig, d['signal32'])
    _check(harm, d['harm32'])
    o64 = O.sins_forward(d['c0'], d['c1'], d['c2'], d['f0_frames'].astype(np.float64), d['U'])
    _check64(o64[0], d['signal64'])
    ph_ref = d['phase32'][..., 0]
    step = phase.shape[1] // ph_ref.shape[1]
    np.testing.assert_allclose(phase[:, ::step], ph_ref, atol=1e-6)


@pytest.mark.parametrize('tag', ['small', 'even', 'one'])
def test_port_autograd_matches_reference_gradients(golden_dir, tag):
    """The stock-PyTorch restatement differentiates to the same control-tensor gradients as the
    reference module itself (tests/golden/make_golden_grad.py) -- it is the gradient oracle of the
    GPU backward tests at sizes without a committed fixture."""
    import torch
    from oracle import torch_port as TP
    d = dict(np.load(os.path.join(golden_dir, f'combsubfast_grad_{tag}.npz')))
    ct = torch.from_numpy(d['ctrl']).double().requires_grad_(True)
    hm, hp, nm = torch.split(ct, 513, dim=-1)
    win = torch.sqrt(torch.hann_window(1024)).double()          # the module buffer is built in fp32 (vocoder.py:434)
    sig, _ = TP.combsubfast_forward(hm, hp, nm, torch.from_numpy(d['f0_frames']).double()[..., None], win,
                                    torch.from_numpy(d['U']).double())
    (sig * torch.from_numpy(d['R']).double()).sum().backward()
    g = ct.grad.numpy()
    assert np.abs(g - d['grad64']).max() <= 2e-6 * np.abs(d['grad64']).max()
