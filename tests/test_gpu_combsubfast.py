"""GPU parity of CombSubFast stage A+B (C ABI -> sm_100a kernels) against the golden vectors
produced by the reference itself and against the CPU oracle."""
import os

import numpy as np
import pytest

from oracle import ddsp_oracle as O
from tests.gpu_util import HAS_CUDA, assert_waveform, ctrl_views, dev, in_kernel_noise, snr_db, torch
from ddsp_b200.synthetic import make_inputs

pytestmark = pytest.mark.gpu

if HAS_CUDA:
    from ddsp_b200 import core


def run_gpu(ctrl, f0_frames, U=None, seed=0, initial_phase=None, window=None):
    hm, hp, nm = ctrl_views(ctrl, 'combsubfast')
    f0 = dev(f0_frames)[..., None]
    ip = None if initial_phase is None else dev(initial_phase)
    pf, prefix, _ = core.phase_stage(f0, 512, 44100, ip, True)
    sig = core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, ip, noise_u=None if U is None else dev(U),
                                 seed=seed, window=window)
    torch.cuda.synchronize()
    return sig.cpu().numpy(), pf.cpu().numpy()


def torch_window():
    return torch.sqrt(torch.hann_window(1024)).cuda()


@pytest.mark.parametrize('tag', ['small', 'odd', 'gui'])
def test_vs_reference_golden(golden_dir, tag):
    d = dict(np.load(os.path.join(golden_dir, f'combsubfast_{tag}.npz')))
    if 'ctrl' not in d:
        d.update(make_inputs(int(d['B']), int(d['F']), 1539, seed=int(d['seed']),
                             zero_f0_fraction=float(d['zero_f0_fraction'])))
    sig, pf = run_gpu(d['ctrl'], d['f0_frames'], d['U'], window=torch_window())
    err, s = assert_waveform(sig, d['signal32'], what=f'combsubfast/{tag} vs reference fp32')
    assert err < 2e-5 and s > 80, (err, s)            # far inside the north-star budget
    assert_waveform(sig, d['signal64'], what=f'combsubfast/{tag} vs reference fp64')
    assert np.abs(pf - d['phase32'][..., 0]).max() < 1e-6


@pytest.mark.parametrize('B,F,zf', [(1, 1, 0.0), (1, 2, 0.0), (3, 50, 0.2), (2, 129, 0.0), (5, 64, 0.1)])
def test_vs_oracle(B, F, zf):
    d = make_inputs(B, F, 1539, seed=100 + F, zero_f0_fraction=zf)
    sig, pf = run_gpu(d['ctrl'], d['f0_frames'], d['U'])
    ref, pf_ref = O.combsubfast_forward(d['ctrl'][..., :513], d['ctrl'][..., 513:1026], d['ctrl'][..., 1026:],
                                        d['f0_frames'], d['U'])
    err, s = assert_waveform(sig, ref, what=f'combsubfast B={B} F={F}')
    assert err < 2e-5 and s > 80, (err, s)
    assert np.abs(pf - pf_ref).max() < 1e-6


def test_initial_phase():
    d = make_inputs(2, 20, 1539, seed=5)
    ip = np.array([1.0, -2.5], np.float32)
    sig, pf = run_gpu(d['ctrl'], d['f0_frames'], d['U'], initial_phase=ip)
    ref, pf_ref = O.combsubfast_forward(d['ctrl'][..., :513], d['ctrl'][..., 513:1026], d['ctrl'][..., 1026:],
                                        d['f0_frames'], d['U'], initial_phase=ip)
    assert_waveform(sig, ref, max_abs=2e-5, snr=80, what='initial_phase')


def test_contiguous_and_strided_controls_agree():
    d = make_inputs(2, 33, 1539, seed=6)
    sig_views, _ = run_gpu(d['ctrl'], d['f0_frames'], d['U'])
    f0 = dev(d['f0_frames'])[..., None]
    pf, prefix, _ = core.phase_stage(f0, 512, 44100)
    parts = [dev(d['ctrl'][..., :513]), dev(d['ctrl'][..., 513:1026]), dev(d['ctrl'][..., 1026:])]   # 3 separate tensors
    sig_sep = core.combsubfast_stage(*parts, f0, prefix, 512, 44100, noise_u=dev(d['U'])).cpu().numpy()
    assert np.array_equal(sig_views, sig_sep)


def test_control_rows_at_every_alignment():
    """The kernel copies a filter row as the 16-byte-aligned window around it (cp.async.bulk): three separately
    allocated control tensors whose rows start 4, 8 and 12 bytes past a 16-byte boundary (row length 2052 B, so every
    row of each tensor has a different skew as well) must give the bits of the packed (B,F,1539) layout."""
    d = make_inputs(3, 41, 1539, seed=17, zero_f0_fraction=0.1)
    packed = dev(d['ctrl'])
    f0 = dev(d['f0_frames'])[..., None]
    U = dev(d['U'])
    pf, prefix, _ = core.phase_stage(f0, 512, 44100)
    ref = core.combsubfast_stage(*torch.split(packed, 513, dim=-1), f0, prefix, 512, 44100, noise_u=U)
    views = []
    for k, off in enumerate((1, 2, 3)):
        flat = torch.zeros(3 * 41 * 513 + 4, device='cuda')
        v = flat[off:off + 3 * 41 * 513].view(3, 41, 513)
        v.copy_(packed[..., 513 * k:513 * (k + 1)])
        assert v.data_ptr() % 16 == 4 * off and v.is_contiguous()
        views.append(v)
    assert torch.equal(ref, core.combsubfast_stage(*views, f0, prefix, 512, 44100, noise_u=U))


def test_batch_and_run_partition_invariance():
    """A clip's waveform must not depend on what else is in the batch nor on how frame pairs are
    partitioned into warp runs (different B -> different run length): bitwise identical."""
    d = make_inputs(6, 200, 1539, seed=8, zero_f0_fraction=0.05)
    sig_all, _ = run_gpu(d['ctrl'], d['f0_frames'], d['U'])
    for b in (0, 3, 5):
        sig_one, _ = run_gpu(d['ctrl'][b:b + 1], d['f0_frames'][b:b + 1], d['U'][b:b + 1])
        assert np.array_equal(sig_all[b], sig_one[0]), b


def test_in_kernel_noise_is_deterministic_and_uniform():
    d = make_inputs(2, 40, 1539, seed=9, noise=False)
    ctrl = d['ctrl'].copy()
    ctrl[..., :513] = -60.0          # harmonic branch off -> signal is the filtered noise alone
    a, _ = run_gpu(ctrl, d['f0_frames'], None, seed=1234)
    b, _ = run_gpu(ctrl, d['f0_frames'], None, seed=1234)
    c, _ = run_gpu(ctrl, d['f0_frames'], None, seed=1235)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    # same statistics as an injected U (flat filter -> variance of (2U-1)/128 = 1/(3*128^2))
    ctrl[..., 1026:] = 0.0
    n, _ = run_gpu(ctrl, d['f0_frames'], None, seed=77)
    var = n[:, 1024:-1024].var()
    assert abs(var * 3 * 128 ** 2 - 1.0) < 0.05
    assert abs(n.mean()) < 1e-4


def test_noise_generator_matches_host_restatement():
    """The counter-based generator is integer work: bit-exact against a numpy restatement."""
    B, F = 2, 9
    d = make_inputs(B, F, 1539, seed=10, noise=False)
    seed = 4242
    U = in_kernel_noise(seed, B, F)
    a, _ = run_gpu(d['ctrl'], d['f0_frames'], None, seed=seed)
    b_, _ = run_gpu(d['ctrl'], d['f0_frames'], U)
    assert np.array_equal(a, b_)


def test_module_window_buffer_vs_exact_window():
    d = make_inputs(1, 30, 1539, seed=11)
    a, _ = run_gpu(d['ctrl'], d['f0_frames'], d['U'], window=torch_window())
    b, _ = run_gpu(d['ctrl'], d['f0_frames'], d['U'], window=None)
    assert np.abs(a - b).max() < 5e-6


def test_headline_shape_properties():
    """Config (2): B=64 x 10 s (F=862).  Size-independent properties instead of a full oracle run:
    (i) a few clips re-synthesised alone are bitwise identical; (ii) those clips match the oracle."""
    B, F = 64, 862
    d = make_inputs(B, F, 1539, seed=1234)
    sig, pf = run_gpu(d['ctrl'], d['f0_frames'], d['U'])
    assert np.all(np.isfinite(sig))
    for b in (0, 31, 63):
        one, _ = run_gpu(d['ctrl'][b:b + 1], d['f0_frames'][b:b + 1], d['U'][b:b + 1])
        assert np.array_equal(sig[b], one[0])
    b = 17
    ref, pf_ref = O.combsubfast_forward(d['ctrl'][b:b + 1, :, :513], d['ctrl'][b:b + 1, :, 513:1026],
                                        d['ctrl'][b:b + 1, :, 1026:], d['f0_frames'][b:b + 1], d['U'][b:b + 1])
    err, s = assert_waveform(sig[b:b + 1], ref, what='headline clip 17')
    assert err < 5e-5 and s > 75, (err, s)
    assert np.abs(pf[b:b + 1] - pf_ref).max() < 1e-6


def test_rejects_cpu_tensors_and_bad_hop():
    from ddsp_b200 import _cabi
    with pytest.raises(_cabi.DDSPB200Error):
        core.phase_stage(torch.zeros(1, 4, 1), 512, 44100)
    with pytest.raises(_cabi.DDSPB200Error):
        core.phase_stage(torch.zeros(1, 4, 1).cuda(), 256, 44100)


def test_vs_stock_pytorch_ops_on_the_same_gpu():
    """fp32 reference of the same op sequence on the same device: the reference's forward restated
    with torch CUDA ops (oracle/torch_port.py; bit-identical to the reference on CPU)."""
    from oracle import torch_port as T
    d = make_inputs(4, 200, 1539, seed=31, zero_f0_fraction=0.05)
    sig, pf = run_gpu(d['ctrl'], d['f0_frames'], d['U'], window=torch_window())
    hm, hp, nm = ctrl_views(d['ctrl'], 'combsubfast')
    with torch.no_grad():
        ref, pf_ref = T.combsubfast_forward(hm, hp, nm, dev(d['f0_frames'])[..., None], torch_window(), noise_u=dev(d['U']))
    err, s = assert_waveform(sig, ref.cpu().numpy(), what='combsubfast vs torch CUDA ops')
    assert err < 3e-5 and s > 80, (err, s)
    dp = np.abs(pf.astype(np.float64) - pf_ref.cpu().numpy())
    assert np.minimum(dp, np.abs(dp - 2 * np.pi)).max() < 1e-6


def test_long_form_five_minutes():
    """Config (4): 5-minute clips (F=25840, T=13.2 M samples) stress the phase accumulation and the
    overlap-add length.  One low-pitched (69-98 Hz) clip and one with unvoiced (f0=0) frames.

    Arbiter: the oracle with an exact cumulative sum.  The reference's CPU path accumulates
    `cumsum(f0/sr)` sequentially in fp64 and drifts by ~1e-5 rotations over 13 M samples; where f0
    interpolates to 0 that drift is multiplied by sr/(f0+1e-3) inside the sinc argument
    (vocoder.py:459), so the *sequential* CPU phase and any scan-based phase (torch's CUDA cumsum,
    these kernels) disagree there by construction.  The phase itself is checked against torch's own
    CUDA ops (the reference's stage A on this GPU)."""
    F = 25840
    d = make_inputs(2, F, 1539, seed=78, zero_f0_fraction=0.02)
    t = np.arange(F) * (512 / 44100)
    d['f0_frames'][0] = (82.0 * 2 ** (0.25 * np.sin(2 * np.pi * t / 7.0))).astype(np.float32)    # 69..98 Hz
    sig, pf = run_gpu(d['ctrl'], d['f0_frames'], d['U'])
    assert np.all(np.isfinite(sig))
    # stage A against the reference's op sequence on this GPU (core.py:7-21,40-49 with torch CUDA ops)
    f0 = dev(d['f0_frames'])[..., None]
    _, _, full = core.phase_stage(f0, 512, 44100, full_rate=True)
    xp = f0.permute(0, 2, 1)
    up = torch.nn.functional.interpolate(torch.cat((xp, xp[:, :, -1:]), 2), size=F * 512 + 1, mode='linear',
                                         align_corners=True)[:, 0, :-1]
    rot = torch.cumsum(up.double() / 44100, axis=1)
    rot = (rot - torch.round(rot)).float()
    dphi = (full.double() - (2 * np.pi * rot).double()).abs()
    dphi = torch.minimum(dphi, (dphi - 2 * np.pi).abs())
    assert dphi.max().item() < 2e-6, dphi.max().item()
    for b in range(2):
        ref, pf_ref = O.combsubfast_forward(d['ctrl'][b:b + 1, :, :513], d['ctrl'][b:b + 1, :, 513:1026],
                                            d['ctrl'][b:b + 1, :, 1026:], d['f0_frames'][b:b + 1], d['U'][b:b + 1],
                                            exact_cumsum=True)
        err, s = assert_waveform(sig[b:b + 1], ref, what=f'5-min clip {b}')
        err_tail = np.abs(sig[b, -2_000_000:] - ref[0, -2_000_000:]).max()
        assert err_tail <= 1e-4, err_tail
        dp = np.abs(pf[b].astype(np.float64) - pf_ref[0])
        assert np.minimum(dp, np.abs(dp - 2 * np.pi)).max() < 2e-6
        print(f'5-min clip {b}: max-abs {err:.2e} snr {s:.1f} dB tail {err_tail:.2e}')
    # the fully voiced clip also agrees with the sequential-cumsum (reference CPU) oracle
    ref_seq, _ = O.combsubfast_forward(d['ctrl'][:1, :, :513], d['ctrl'][:1, :, 513:1026], d['ctrl'][:1, :, 1026:],
                                       d['f0_frames'][:1], d['U'][:1])
    assert_waveform(sig[:1], ref_seq, what='5-min voiced clip vs sequential-cumsum oracle')
