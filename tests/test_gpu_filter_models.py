"""GPU parity of frequency_filter, CombSub (old) and Sins against the golden vectors produced by
the reference itself and against the CPU oracle.  Calls go through the C ABI."""
import os

import numpy as np
import pytest

from oracle import ddsp_oracle as O
from tests.gpu_util import HAS_CUDA, SPLITS, assert_waveform, ctrl_views, dev, torch
from ddsp_b200.synthetic import make_inputs

pytestmark = pytest.mark.gpu

if HAS_CUDA:
    from ddsp_b200 import core


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'core.npz'))


# ---- frequency_filter (core.py:331-336) ------------------------------------------------------
@pytest.mark.parametrize('name', ['none', 'hann'])
def test_frequency_filter_static_window(gold, name):
    y = core.frequency_filter(dev(gold['ff_audio']), dev(gold['ff_mags']), hann_window=(name == 'hann')).cpu().numpy()
    assert_waveform(y, gold[f'ff_{name}_y'], max_abs=2e-5, snr=90, what=f'frequency_filter/{name}')


def test_frequency_filter_dynamic_window(gold):
    f0f = gold['ff_f0f'].astype(np.float32)
    hw = 1.5 * 44100 / (torch.from_numpy(f0f) + 1e-3)
    y = core.frequency_filter(dev(gold['ff_audio']), dev(gold['ff_mags2']), True, hw.cuda()).cpu().numpy()
    assert_waveform(y, gold['ff_dyn_y'], max_abs=3e-5, snr=90, what='frequency_filter/dynamic (half_width_frames)')
    y2 = core.frequency_filter(dev(gold['ff_audio']), dev(gold['ff_mags2']), True, f0_frames=dev(f0f)).cpu().numpy()
    assert_waveform(y2, gold['ff_dyn_y'], max_abs=3e-5, snr=90, what='frequency_filter/dynamic (f0_frames)')


def test_frequency_filter_allpass_complex(gold):
    gd = np.pi * np.tanh(gold['ff_ph'])
    mags = np.exp(1j * np.cumsum(gd, -1)).astype(np.complex64)
    y = core.frequency_filter(dev(gold['ff_audio']), torch.from_numpy(mags).cuda(), hann_window=False).cpu().numpy()
    assert_waveform(y, gold['ff_ap_y'], max_abs=3e-5, snr=85, what='frequency_filter/allpass (complex magnitudes)')
    # the same filter handed over as the raw control tensor (what the synthesizer modules do)
    y2 = core.frequency_filter(dev(gold['ff_audio']), dev(gold['ff_ph'].astype(np.float32)), hann_window=False,
                               encoding=core.MAG_ALLPASS_TANH).cpu().numpy()
    assert_waveform(y2, gold['ff_ap_y'], max_abs=3e-5, snr=85, what='frequency_filter/allpass (tanh control)')


def test_frequency_filter_replays_the_reference_call_sequence(gold):
    """vocoder.py:541-542 verbatim through the mirror: `frequency_filter(x, torch.complex(src_param, zeros), hann_window=True,
    half_width_frames=1.5*sr/(f0_frames+1e-3))` with n_mag = 512 -- complex magnitudes with a zero imaginary part."""
    from ddsp_b200 import _cabi
    f0f = torch.from_numpy(gold['ff_f0f'].astype(np.float32)).cuda()
    src = dev(gold['ff_mags2'])
    y = core.frequency_filter(dev(gold['ff_audio']), torch.complex(src, torch.zeros_like(src)), hann_window=True,
                              half_width_frames=1.5 * 44100 / (f0f + 1e-3)).cpu().numpy()
    assert_waveform(y, gold['ff_dyn_y'], max_abs=3e-5, snr=90, what='frequency_filter/reference call sequence')
    with pytest.raises(_cabi.DDSPB200Error):
        core.frequency_filter(dev(gold['ff_audio']), torch.complex(src, torch.ones_like(src)), hann_window=True)


@pytest.mark.parametrize('B,F,n_mag', [(1, 1, 256), (2, 3, 256), (1, 9, 512), (3, 40, 256), (2, 33, 512)])
def test_frequency_filter_vs_oracle_ragged_runs(B, F, n_mag):
    """Frame counts that do not divide into runs evenly, single-frame clips, seams between runs."""
    rng = np.random.default_rng(F * 7 + n_mag)
    audio = (rng.random((B, F * 512)) * 2 - 1).astype(np.float32)
    mags = np.exp(0.57 * rng.standard_normal((B, F, n_mag))).astype(np.float32)
    y = core.frequency_filter(dev(audio), dev(mags), hann_window=True).cpu().numpy()
    ref = O.frequency_filter(audio, mags.astype(np.complex128), hann_window=True)
    assert_waveform(y, ref, max_abs=3e-5, snr=90, what=f'frequency_filter B={B} F={F} n_mag={n_mag}')


def test_frequency_filter_is_linear_and_deterministic():
    rng = np.random.default_rng(3)
    B, F = 4, 300
    a1 = (rng.random((B, F * 512)) * 2 - 1).astype(np.float32)
    a2 = (rng.random((B, F * 512)) * 2 - 1).astype(np.float32)
    mags = dev(np.exp(0.57 * rng.standard_normal((B, F, 256))).astype(np.float32))
    y1 = core.frequency_filter(dev(a1), mags)
    y2 = core.frequency_filter(dev(a2), mags)
    y12 = core.frequency_filter(dev(a1 + a2), mags)
    assert (y1 + y2 - y12).abs().max().item() < 2e-5
    assert torch.equal(y1, core.frequency_filter(dev(a1), mags))        # atomics at run seams: two addends -> bitwise stable


@pytest.mark.parametrize('n_mag', [256, 512])
def test_frequency_filter_audio_eight_byte_aligned(n_mag):
    """The convolution kernels stage a frame's input samples by bulk copies of the 16-byte-aligned window around them:
    audio that is only 8-byte aligned (the ABI's requirement) takes the skewed path and must give the same bits."""
    rng = np.random.default_rng(11)
    B, F = 3, 37
    a = (rng.random((B, F * 512)) * 2 - 1).astype(np.float32)
    mags = dev(np.exp(0.57 * rng.standard_normal((B, F, n_mag))).astype(np.float32))
    aligned = dev(a)
    flat = torch.zeros(B * F * 512 + 2, device='cuda')
    skewed = flat[2:].view(B, F * 512)                      # contiguous, data pointer 8 bytes past a 16-byte boundary
    skewed.copy_(aligned)
    assert aligned.data_ptr() % 16 == 0 and skewed.data_ptr() % 16 == 8 and skewed.is_contiguous()
    assert torch.equal(core.frequency_filter(aligned, mags), core.frequency_filter(skewed, mags))


def test_frequency_filter_batch_mismatch_raises():
    with pytest.raises(ValueError):
        core.frequency_filter(torch.zeros(2, 1024).cuda(), torch.zeros(3, 2, 256).cuda())


# ---- CombSub (old) and Sins -------------------------------------------------------------------
def _load(golden_dir, model, tag):
    d = dict(np.load(os.path.join(golden_dir, f'{model}_{tag}.npz')))
    if 'ctrl' not in d:
        d.update(make_inputs(int(d['B']), int(d['F']), sum(SPLITS[model]), seed=int(d['seed']),
                             zero_f0_fraction=float(d['zero_f0_fraction'])))
    return d


def run_combsub(ctrl, f0_frames, U, seed=0):
    gd, hm, nm = ctrl_views(ctrl, 'combsub')
    f0 = dev(f0_frames)[..., None]
    pf, prefix, _ = core.phase_stage(f0, 512, 44100)
    sig, harm, noise = core.combsub_stage(gd, hm, nm, f0, prefix, 512, 44100, noise_u=None if U is None else dev(U), seed=seed)
    torch.cuda.synchronize()
    return sig.cpu().numpy(), pf.cpu().numpy(), harm.cpu().numpy(), noise.cpu().numpy()


def run_sins(ctrl, f0_frames, U, seed=0):
    am, gd, nm = ctrl_views(ctrl, 'sins')
    f0 = dev(f0_frames)[..., None]
    pf, _, phase = core.phase_stage(f0, 512, 44100, full_rate=True)
    sig, harm, noise = core.sins_stage(am, gd, nm, f0, phase, 512, 44100, noise_u=None if U is None else dev(U), seed=seed)
    torch.cuda.synchronize()
    return sig.cpu().numpy(), phase.cpu().numpy(), harm.cpu().numpy(), noise.cpu().numpy()


@pytest.mark.parametrize('tag', ['small', 'odd', 'gui'])
def test_combsub_vs_reference_golden(golden_dir, tag):
    d = _load(golden_dir, 'combsub', tag)
    sig, pf, harm, noise = run_combsub(d['ctrl'], d['f0_frames'], d['U'])
    err, s = assert_waveform(sig, d['signal32'], what=f'combsub/{tag} signal vs reference fp32')
    assert err < 5e-5 and s > 75, (err, s)
    assert_waveform(harm, d['harm32'], what=f'combsub/{tag} harmonic')
    if 'noise32' in d:
        assert_waveform(noise, d['noise32'], what=f'combsub/{tag} noise')
    assert np.abs(pf - d['phase32'][..., 0]).max() < 1e-6


@pytest.mark.parametrize('tag', ['small', 'odd', 'gui'])
def test_sins_vs_reference_golden(golden_dir, tag):
    d = _load(golden_dir, 'sins', tag)
    sig, phase, harm, noise = run_sins(d['ctrl'], d['f0_frames'], d['U'])
    err, s = assert_waveform(sig, d['signal32'], what=f'sins/{tag} signal vs reference fp32')
    assert err < 5e-5 and s > 75, (err, s)
    assert_waveform(harm, d['harm32'], what=f'sins/{tag} harmonic')
    ph_ref = d['phase32'][..., 0]
    step = phase.shape[1] // ph_ref.shape[1]
    dphi = np.abs(phase[:, ::step].astype(np.float64) - ph_ref)
    assert np.minimum(dphi, np.abs(dphi - 2 * np.pi)).max() < 1e-6


@pytest.mark.parametrize('B,F', [(1, 1), (3, 50), (2, 129)])
def test_combsub_vs_oracle(B, F):
    d = make_inputs(B, F, 1024, seed=300 + F, zero_f0_fraction=0.1)
    sig, pf, harm, noise = run_combsub(d['ctrl'], d['f0_frames'], d['U'])
    ref = O.combsub_forward(d['ctrl'][..., :256], d['ctrl'][..., 256:768], d['ctrl'][..., 768:], d['f0_frames'], d['U'])
    assert_waveform(sig, ref[0], max_abs=5e-5, snr=75, what=f'combsub B={B} F={F}')
    assert_waveform(noise, ref[3], max_abs=2e-5, snr=80, what=f'combsub noise B={B} F={F}')


@pytest.mark.parametrize('B,F', [(1, 1), (3, 50), (2, 129)])
def test_sins_vs_oracle(B, F):
    d = make_inputs(B, F, 640, seed=400 + F, zero_f0_fraction=0.1)
    sig, phase, harm, noise = run_sins(d['ctrl'], d['f0_frames'], d['U'])
    ref = O.sins_forward(d['ctrl'][..., :128], d['ctrl'][..., 128:384], d['ctrl'][..., 384:], d['f0_frames'], d['U'])
    assert_waveform(sig, ref[0], max_abs=5e-5, snr=75, what=f'sins B={B} F={F}')


def test_sins_nyquist_mask_in_the_oscillator():
    """f0 = 173 Hz: harmonic 127 (21971 Hz) is below sr/2 and kept (factor 1+1e-7), harmonic 128
    (22144 Hz) is masked to 1e-7 (not 0) -- core.py:27.  Drive one harmonic at a time, unit amplitude."""
    B, F = 1, 6
    f0 = np.full((B, F), 173.0, np.float32)
    ctrl = np.full((B, F, 640), -30.0, np.float32)
    ctrl[..., 128:384] = 0.0                      # group delay control 0 -> identity all-pass
    for k, expect in ((126, 1.0), (127, 1e-7)):
        c = ctrl.copy()
        c[..., k] = np.log(128.0)                 # amplitude exp(c)/128 = 1 before the mask
        sig, phase, harm, noise = run_sins(c, f0, np.full((B, F * 512), 0.5, np.float32))
        peak = np.abs(harm[:, 1024:-1024]).max()
        assert abs(peak / expect - 1.0) < 2e-2, (k, peak, expect)


@pytest.mark.parametrize('model', ['combsub', 'sins'])
def test_in_kernel_noise_matches_host_restatement(model):
    """Without an injected U the noise filter draws the same (seed, clip, hop, lane) stream as CombSubFast:
    injecting the host restatement of that stream reproduces all three outputs bit for bit."""
    from tests.gpu_util import in_kernel_noise
    B, F = 2, 11
    d = make_inputs(B, F, sum(SPLITS[model]), seed=77, noise=False)
    run = run_combsub if model == 'combsub' else run_sins
    seed = 31337
    a = run(d['ctrl'], d['f0_frames'], None, seed=seed)
    b = run(d['ctrl'], d['f0_frames'], in_kernel_noise(seed, B, F))
    for x, y in zip((a[0], a[2], a[3]), (b[0], b[2], b[3])):
        assert np.array_equal(x, y)
    assert np.abs(a[3]).max() > 0


def test_models_batch_invariance():
    d = make_inputs(5, 120, 1024, seed=21, zero_f0_fraction=0.05)
    sig_all, *_ = run_combsub(d['ctrl'], d['f0_frames'], d['U'])
    one, *_ = run_combsub(d['ctrl'][2:3], d['f0_frames'][2:3], d['U'][2:3])
    assert np.array_equal(sig_all[2], one[0])


def test_drop_in_modules_forward_signature():
    """The modules keep the reference's forward signature / return tuple (vocoder.py:381,437,504)."""
    from ddsp_b200 import vocoder as V

    class FixedCtrl(torch.nn.Module):
        def __init__(self, names, ctrl):
            super().__init__()
            self.names, self.ctrl = names, ctrl

        def forward(self, units, f0, phase, volume, spk_id, spk_mix_dict=None):
            assert phase.shape == f0.shape[:2]
            return dict(zip(self.names, torch.split(self.ctrl, self.sizes, dim=-1)))

    B, F = 2, 16
    units = torch.zeros(B, F, 4).cuda()
    vol = torch.zeros(B, F).cuda()
    spk = torch.ones(B, 1, dtype=torch.long).cuda()
    for cls, args, names, sizes in [
            (V.CombSubFast, (44100, 512), ['harmonic_magnitude', 'harmonic_phase', 'noise_magnitude'], [513] * 3),
            (V.CombSub, (44100, 512, 256, 512, 256), ['group_delay', 'harmonic_magnitude', 'noise_magnitude'], [256, 512, 256]),
            (V.Sins, (44100, 512, 128, 256, 256), ['amplitudes', 'group_delay', 'noise_magnitude'], [128, 256, 256])]:
        d = make_inputs(B, F, sum(sizes), seed=7)
        fc = FixedCtrl(names, dev(d['ctrl']))
        fc.sizes = sizes
        model = cls(*args, unit2ctrl=fc).cuda().eval()
        assert {'sampling_rate', 'block_size'} <= set(model.state_dict().keys())
        with torch.no_grad():
            signal, phase, (harm, noise) = model(units, dev(d['f0_frames'])[..., None], vol, spk)
        assert signal.shape == (B, F * 512)
        assert phase.shape == ((B, F * 512, 1) if cls is V.Sins else (B, F, 1))
        assert harm.shape == noise.shape == signal.shape
        signal *= 0.5        # callers mutate the result in place (main.py:159, gui.py:127)


@pytest.mark.parametrize('model', ['combsub', 'sins'])
def test_headline_shape_properties(model):
    """B=64 x 10 s (F=862), the headline batch: a clip re-synthesised alone agrees to the last ulp or
    two (clips are independent; the 4-frame overlap-add is associated per run, and the run length
    depends on the batch size, so unlike CombSubFast the match is not bitwise), repeated runs are
    bitwise identical, and one clip matches the oracle."""
    B, F = 64, 862
    run = run_combsub if model == 'combsub' else run_sins
    d = make_inputs(B, F, sum(SPLITS[model]), seed=1234, zero_f0_fraction=0.02)
    sig, _, harm, noise = run(d['ctrl'], d['f0_frames'], d['U'])
    assert np.all(np.isfinite(sig))
    sig2, *_ = run(d['ctrl'], d['f0_frames'], d['U'])
    assert np.array_equal(sig, sig2)                       # deterministic (two-addend atomics at run seams)
    for b in (0, 63):
        one, *_ = run(d['ctrl'][b:b + 1], d['f0_frames'][b:b + 1], d['U'][b:b + 1])
        assert np.abs(sig[b] - one[0]).max() < 2e-6, b
    b = 17
    a, c, e = SPLITS[model]
    fwd = O.combsub_forward if model == 'combsub' else O.sins_forward
    ref = fwd(d['ctrl'][b:b + 1, :, :a], d['ctrl'][b:b + 1, :, a:a + c], d['ctrl'][b:b + 1, :, a + c:],
              d['f0_frames'][b:b + 1], d['U'][b:b + 1])
    err, s = assert_waveform(sig[b:b + 1], ref[0], what=f'{model} headline clip 17')
    assert err < 5e-5 and s > 75, (err, s)
