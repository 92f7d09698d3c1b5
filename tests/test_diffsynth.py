"""`ddsp_b200.diffsynth` (the torch-op restatement behind the backward of the Sins / CombSub-old drop-ins) and
`ddsp_b200.loss` against the REFERENCE: forward values and autograd gradients recorded from the unmodified reference
modules (tests/golden/make_golden_grad.py filter_models -> {combsub,sins}_grad_small.npz), on CPU in fp64."""
import math
import os

import numpy as np
import pytest
import torch

from ddsp_b200 import diffsynth as D

SPLITS = {'combsub': (256, 512, 256), 'sins': (128, 256, 256)}


def stage_a(f0_frames, hop=512, sr=44100):
    """upsample + fp64 cumsum + wrap (core.py:7-51) in torch, CPU."""
    f0 = D.upsample(f0_frames.reshape(f0_frames.shape[0], -1, 1), hop)[..., 0]
    rot = torch.cumsum(f0.double() / sr, dim=1)
    rot = rot - torch.round(rot)
    return rot.to(f0_frames.dtype)


@pytest.mark.parametrize('model', ['combsub', 'sins'])
def test_forward_and_gradients_match_the_reference(golden_dir, model):
    g = dict(np.load(os.path.join(golden_dir, f'{model}_grad_small.npz')))
    ctrl = torch.from_numpy(g['ctrl']).double().requires_grad_(True)
    c0, c1, c2 = torch.split(ctrl, list(SPLITS[model]), dim=-1)
    f0 = torch.from_numpy(g['f0_frames']).double()
    U = torch.from_numpy(g['U']).double()
    rot = stage_a(f0)
    if model == 'combsub':
        sig, harm, noise = D.combsub_stage(c0, c1, c2, f0, rot, U, 512, 44100)
    else:
        sig, harm, noise = D.sins_stage(c0, c1, c2, f0, 2 * math.pi * rot, U, 512, 44100)
    assert np.abs(sig.detach().numpy() - g['signal64']).max() < 2e-6        # fixture stored as fp32
    (sig * torch.from_numpy(g['R']).double()).sum().backward()
    ref = g['grad64']
    err = np.abs(ctrl.grad.numpy() - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err


def test_frequency_filter_batch_mismatch_raises():
    with pytest.raises(ValueError):
        D.ltv_fir(torch.zeros(2, 1024), torch.zeros(3, 2, 510), 512)


def test_sss_loss_known_values():
    """Identical signals: converge term 0, log term 0; scaled copy: closed form."""
    from ddsp_b200.loss import SSSLoss, RSSLoss
    torch.manual_seed(0)
    x = torch.randn(2, 8000)
    L = SSSLoss(256)
    assert float(L(x, x)) == 0.0
    # x_pred = 2 x: ||S - 2S|| / ||S + 2S|| = 1/3 ; |log S - log 2S| = log 2 (eps negligible)
    v = float(SSSLoss(256, eps=0.0)(x, 2 * x))
    assert abs(v - (1 / 3 + math.log(2))) < 1e-5
    torch.manual_seed(3)
    r = RSSLoss(128, 512, 3, device='cpu')
    a = float(r(x, 0.5 * x))
    assert abs(a - (1 / 3 + math.log(2))) < 1e-4          # scale invariance of both terms' structure: same value at any n_fft
