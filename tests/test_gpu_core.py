"""GPU parity of the standalone core ops and stage A against the oracle, the reference's KATs
(ddsp/core.py:54-97) and the golden vectors.  Calls go through the C ABI (ctypes)."""
import math
import os

import numpy as np
import pytest

from oracle import ddsp_oracle as O
from tests.gpu_util import HAS_CUDA, dev, torch

pytestmark = pytest.mark.gpu
F32 = np.float32

if HAS_CUDA:
    from ddsp_b200 import core


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'core.npz'))


def test_upsample_bit_exact_vs_oracle_and_golden(gold):
    x = gold['upsample_x']
    y = core.upsample(dev(x), 512).cpu().numpy()
    assert np.array_equal(y, O.upsample(x, 512))
    assert np.array_equal(y[:, ::37, :], gold['upsample_y_sub'])


def test_upsample_bit_exact_vs_torch_cuda():
    """The op this replaces on the GPU: F.interpolate(linear, align_corners=True) (core.py:17)."""
    g = torch.Generator(device='cpu').manual_seed(3)
    for (B, F, C, factor) in [(2, 37, 3, 512), (1, 300, 1, 512), (2, 11, 4, 100), (1, 5, 1, 7)]:
        x = (torch.rand(B, F, C, generator=g) * 800).cuda()
        xp = x.permute(0, 2, 1)
        ref = torch.nn.functional.interpolate(torch.cat((xp, xp[:, :, -1:]), 2), size=F * factor + 1,
                                              mode='linear', align_corners=True)[:, :, :-1].permute(0, 2, 1)
        got = core.upsample(x, factor)
        assert torch.equal(got, ref.contiguous()), (B, F, C, factor)


def test_upsample_strided_input():
    x = (torch.rand(2, 9, 6) * 800).cuda()
    v = x[:, :, 1::2]
    assert torch.equal(core.upsample(v, 512), core.upsample(v.contiguous(), 512))


# ---- the five reference KATs (core.py:54-97) ---------------------------------------------
def test_kat_dtype():
    fo = torch.tensor([[1.0, 1.0, 1.0]]).cuda()
    assert core.fo_to_rot(fo, 1, precise=False).dtype == fo.dtype
    assert core.fo_to_rot(fo, 1, precise=True).dtype == fo.dtype


@pytest.mark.parametrize('precise', [False, True])
def test_kat_stablefo(precise):
    rot = core.fo_to_rot(torch.tensor([[1.0, 1.0, 1.0]]).cuda(), 4, None, precise).cpu()
    assert torch.allclose(torch.tensor([[+0.25, +0.50, -0.25]]), rot)


@pytest.mark.parametrize('precise', [False, True])
def test_kat_fm(precise):
    rot = core.fo_to_rot(torch.tensor([[1.0, 2.0, 3.0]]).cuda(), 4, None, precise).cpu()
    assert torch.allclose(torch.tensor([[+0.25, -0.25, -0.50]]), rot)


@pytest.mark.parametrize('precise', [False, True])
def test_kat_init_phase(precise):
    rot = core.fo_to_rot(torch.tensor([[1.0, 1.0, 1.0]]).cuda(), 4, torch.tensor([math.pi]).cuda(), precise).cpu()
    assert torch.allclose(torch.tensor([[-0.25, 0.0, +0.25]]), rot, atol=1e-7)


def test_kat_fm_init_batch():
    fo = torch.tensor([[1.0, 1.0, 1.0], [1.0, 2.0, 3.0]]).cuda()
    ip = torch.tensor([math.pi, 0.0]).cuda()
    rot = core.fo_to_rot(fo, 4, ip, True).cpu()
    assert torch.allclose(torch.tensor([[-0.25, 0.0, +0.25], [+0.25, -0.25, -0.50]]), rot, atol=1e-5)


def test_fo_to_rot_precise_vs_oracle(gold):
    fo = O.upsample(gold['upsample_x'][:, :, :1], 512)[..., 0]
    rot = core.fo_to_rot(dev(fo), 44100, None, True).cpu().numpy()
    ref = O.fo_to_rot(fo, 44100, None, True)
    # block-parallel fp64 scan vs sequential fp64 cumsum: differs only where the wrapped value sits
    # within 1e-9 of a rounding boundary of the fp32 cast
    assert np.abs(rot - ref).max() <= 6e-8
    assert np.mean(rot == ref) > 0.999
    assert np.abs(rot[:, ::37] - gold['rot_precise_sub']).max() <= 6e-8
    ip = gold['rot_ip']
    rot_ip = core.fo_to_rot(dev(fo), 44100, dev(ip), True).cpu().numpy()
    assert np.abs(rot_ip[:, ::37] - gold['rot_precise_ip_sub']).max() <= 2e-7


def test_fo_to_rot_ragged_lengths():
    for T in (1, 7, 2047, 2048, 2049, 70001):
        fo = (torch.rand(2, T) * 700 + 65).cuda()
        rot = core.fo_to_rot(fo, 44100, None, True).cpu().numpy()
        ref = O.fo_to_rot(fo.cpu().numpy(), 44100, None, True)
        assert np.abs(rot - ref).max() <= 6e-8, T


def test_remove_above_fmax_bit_exact(gold):
    out = core.remove_above_fmax(dev(gold['mask_amp']), dev(gold['mask_pitch']), 22050.0).cpu().numpy()
    assert np.array_equal(out, gold['mask_out'])
    assert np.array_equal(out, O.remove_above_fmax(gold['mask_amp'], gold['mask_pitch'], F32(22050.0)))


def test_remove_above_fmax_mask_values():
    """harmonic-index / Nyquist mask must be bit-exact (north star): probe every (f0, k)."""
    rng = np.random.default_rng(5)
    pitch = np.concatenate([rng.uniform(0, 900, 4000), [172.265625, 344.53125, 0.0, 22050.0, 22049.998]]).astype(F32)
    pitch = pitch.reshape(1, -1, 1)
    amp = np.ones((1, pitch.shape[1], 128), F32)
    out = core.remove_above_fmax(dev(amp), dev(pitch), 22050.0).cpu().numpy()
    assert np.array_equal(out, O.nyquist_mask(pitch, 128, F32(22050.0)))


def _wrapped_err(a, b):
    """Phase error modulo one rotation: when the fp64 rotation count lands within an ulp of k+0.5
    (e.g. a constant 65 Hz contour reaches exactly 6.5 rotations at sample 4410) the half-to-even
    wrap may fall on either side, +pi or -pi; both are the same phase."""
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return np.minimum(d, np.abs(d - 2 * np.pi)).max()


def test_phase_stage_vs_oracle():
    from ddsp_b200.synthetic import make_f0
    rng = np.random.default_rng(21)
    for (B, F) in [(1, 1), (2, 7), (3, 130), (2, 862), (70, 40)]:
        f0 = make_f0(B, F, rng, zero_f0_fraction=0.1)
        ip = rng.uniform(-3, 3, B).astype(F32)
        for init in (None, ip):
            pf, prefix, full = core.phase_stage(dev(f0)[..., None], 512, 44100,
                                                None if init is None else dev(init), True, full_rate=True)
            f0_up, rot, pf_ref = O.stage_a(f0, 44100, 512, init, True)
            # prefix = exclusive per-hop sums of the upsampled fp32 f0 (exact in fp64)
            hop_sums = f0_up.astype(np.float64).reshape(B, F, 512).sum(-1)
            ref_prefix = np.cumsum(hop_sums, 1) - hop_sums
            if init is not None:      # the initial phase rides in the prefix, in Hz*samples
                ref_prefix = ref_prefix + (init.astype(np.float64) / 2 / np.pi * 44100)[:, None]
            np.testing.assert_allclose(prefix.cpu().numpy(), ref_prefix, rtol=1e-14, atol=1e-9)
            assert _wrapped_err(pf.cpu().numpy(), pf_ref) <= 4e-7, (B, F)
            ref_full = (F32(2 * np.pi) * rot).astype(F32)
            assert _wrapped_err(full.cpu().numpy(), ref_full) <= 4e-7, (B, F)
            assert np.mean(full.cpu().numpy() == ref_full) > 0.995


def test_phase_stage_two_kernel_path_equals_fused():
    """B*F large enough for the spread (two-launch) path vs tiny-B fused path: same numbers."""
    from ddsp_b200.synthetic import make_f0
    rng = np.random.default_rng(22)
    f0 = make_f0(2, 3000, rng)
    pf2, pre2, _ = core.phase_stage(dev(f0), 512, 44100)          # B=2, F=3000 -> spread path
    ref = O.stage_a(f0, 44100, 512)[2]
    assert _wrapped_err(pf2.cpu().numpy(), ref) <= 4e-7


def test_apply_frame_mask_bit_exact():
    """signal *= upsample(mask) (main.py:116,159) fused: bit-identical to the two-op reference form."""
    g = torch.Generator().manual_seed(9)
    B, F = 3, 41
    sig = torch.randn(B, F * 512, generator=g).cuda()
    mask = (torch.rand(B, F, generator=g) > 0.4).float().cuda()
    ref = sig * core.upsample(mask[..., None], 512)[..., 0]
    out = core.apply_frame_mask_(sig.clone(), mask[..., None])
    assert torch.equal(out, ref)
    xp = mask[:, None, :]
    up = torch.nn.functional.interpolate(torch.cat((xp, xp[:, :, -1:]), 2), size=F * 512 + 1, mode='linear',
                                         align_corners=True)[:, 0, :-1]
    assert torch.equal(out, sig * up)
