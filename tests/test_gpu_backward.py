"""GPU parity of the CombSubFast stage-B gradient (C ABI -> sm_100a kernel) against gradients that
autograd produced through the reference module (tests/golden/combsubfast_grad_*.npz) and through the
stock-PyTorch restatement in oracle/torch_port.py (fp64, CPU)."""
import os

import numpy as np
import pytest

from tests.gpu_util import HAS_CUDA, ctrl_views, dev, torch
from ddsp_b200.synthetic import make_inputs

pytestmark = pytest.mark.gpu

if HAS_CUDA:
    from ddsp_b200 import core
    from ddsp_b200.vocoder import CombSubFast

GRAD_REL_TOL = 1e-4        # max |g - g_ref| relative to max |g_ref| of the same tensor


def run_backward(ctrl, f0_frames, R, U=None, seed=0, window=None, initial_phase=None):
    hm, hp, nm = ctrl_views(ctrl, 'combsubfast')
    f0 = dev(f0_frames)[..., None]
    ip = None if initial_phase is None else dev(initial_phase)
    _, prefix, _ = core.phase_stage(f0, 512, 44100, ip, True)
    g = core.combsubfast_backward_stage(dev(R), hm, hp, nm, f0, prefix, 512, 44100,
                                        noise_u=None if U is None else dev(U), seed=seed, window=window)
    torch.cuda.synchronize()
    return np.concatenate([t.cpu().numpy() for t in g], axis=-1)


def port_grad(ctrl, f0_frames, U, R, initial_phase=None):
    """fp64 autograd through the op-for-op PyTorch restatement, on CPU."""
    from oracle import torch_port as TP
    ct = torch.from_numpy(ctrl).double().requires_grad_(True)
    hm, hp, nm = torch.split(ct, 513, dim=-1)
    win = torch.sqrt(torch.hann_window(1024, dtype=torch.float64))
    ip = None if initial_phase is None else torch.from_numpy(initial_phase).double()
    sig, _ = TP.combsubfast_forward(hm, hp, nm, torch.from_numpy(f0_frames).double()[..., None], win,
                                    torch.from_numpy(U).double(), initial_phase=ip)
    (sig * torch.from_numpy(R).double()).sum().backward()
    return ct.grad.numpy()


def assert_grad(g, ref, what=''):
    worst = 0.0
    for i, name in enumerate(('harmonic_magnitude', 'harmonic_phase', 'noise_magnitude')):
        a, b = g[..., 513 * i:513 * (i + 1)], ref[..., 513 * i:513 * (i + 1)]
        scale = np.abs(b).max() + 1e-30
        err = np.abs(a - b).max() / scale
        worst = max(worst, err)
        assert err <= GRAD_REL_TOL, f'{what} d/d{name}: rel err {err:.3e} (scale {scale:.3e})'
    return worst


@pytest.mark.parametrize('tag', ['small', 'even', 'one'])
def test_vs_reference_golden_gradients(golden_dir, tag):
    d = dict(np.load(os.path.join(golden_dir, f'combsubfast_grad_{tag}.npz')))
    g = run_backward(d['ctrl'], d['f0_frames'], d['R'], d['U'], window=torch.sqrt(torch.hann_window(1024)).cuda())
    worst = assert_grad(g, d['grad64'], f'{tag} vs reference fp64')
    assert_grad(g, d['grad32'], f'{tag} vs reference fp32')
    # as close to the fp64 arbiter as the reference's own fp32 autograd is (within a small factor)
    ref_gap = np.abs(d['grad32'] - d['grad64']).max() / np.abs(d['grad64']).max()
    assert worst < max(10 * ref_gap, 2e-5), (worst, ref_gap)


@pytest.mark.parametrize('B,F,zf', [(1, 2, 0.0), (3, 50, 0.2), (2, 129, 0.0), (5, 64, 0.1), (1, 301, 0.0)])
def test_vs_port_autograd(B, F, zf):
    d = make_inputs(B, F, 1539, seed=300 + F, zero_f0_fraction=zf)
    R = np.random.default_rng(F).standard_normal((B, F * 512)).astype(np.float32)
    g = run_backward(d['ctrl'], d['f0_frames'], R, d['U'])
    assert_grad(g, port_grad(d['ctrl'], d['f0_frames'], d['U'], R), f'B={B} F={F}')


def test_initial_phase_and_strided_gradient_layout():
    d = make_inputs(2, 20, 1539, seed=31)
    ip = np.array([1.0, -2.5], np.float32)
    R = np.random.default_rng(1).standard_normal((2, 20 * 512)).astype(np.float32)
    g = run_backward(d['ctrl'], d['f0_frames'], R, d['U'], initial_phase=ip)
    assert_grad(g, port_grad(d['ctrl'], d['f0_frames'], d['U'], R, ip), 'initial_phase')


def test_in_kernel_noise_gradient_matches_forward_stream():
    """Without an injected U the backward pass must regenerate the noise the forward pass drew from
    (seed, clip, hop): d/d noise_magnitude then agrees with a central finite difference of the forward."""
    B, F = 2, 16
    d = make_inputs(B, F, 1539, seed=32, noise=False)
    R = np.random.default_rng(2).standard_normal((B, F * 512)).astype(np.float32)
    seed = 777
    g = run_backward(d['ctrl'], d['f0_frames'], R, None, seed=seed)
    assert np.array_equal(g, run_backward(d['ctrl'], d['f0_frames'], R, None, seed=seed))     # deterministic
    assert not np.array_equal(g, run_backward(d['ctrl'], d['f0_frames'], R, None, seed=seed + 1))
    # the same stream injected as a tensor gives the same gradient bit for bit, and that one is tied to autograd
    from tests.gpu_util import in_kernel_noise
    U = in_kernel_noise(seed, B, F)
    assert np.array_equal(g, run_backward(d['ctrl'], d['f0_frames'], R, U))
    assert_grad(g, port_grad(d['ctrl'], d['f0_frames'], U, R), 'in-kernel noise')

    def loss(ctrl):
        hm, hp, nm = ctrl_views(ctrl, 'combsubfast')
        f0 = dev(d['f0_frames'])[..., None]
        _, prefix, _ = core.phase_stage(f0, 512, 44100, None, True)
        sig = core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, None, seed=seed)
        return float((sig.double() * dev(R).double()).sum())

    rng = np.random.default_rng(3)
    for _ in range(3):
        delta = rng.standard_normal(d['ctrl'].shape).astype(np.float32)
        eps = 1e-2
        fd = (loss(d['ctrl'] + eps * delta) - loss(d['ctrl'] - eps * delta)) / (2 * eps)
        an = float((g.astype(np.float64) * delta).sum())
        assert abs(fd - an) <= 5e-3 * max(abs(an), 1.0), (fd, an)


def test_linearity_and_batch_invariance_headline_shape():
    """B=64 x 10 s: the gradient is linear in dL/dsignal and clips are independent."""
    B, F = 64, 862
    d = make_inputs(B, F, 1539, seed=33, noise=False)
    rng = np.random.default_rng(4)
    R1 = rng.standard_normal((B, F * 512)).astype(np.float32)
    R2 = rng.standard_normal((B, F * 512)).astype(np.float32)
    g1 = run_backward(d['ctrl'], d['f0_frames'], R1, None, seed=5)
    g2 = run_backward(d['ctrl'], d['f0_frames'], R2, None, seed=5)
    g12 = run_backward(d['ctrl'], d['f0_frames'], R1 + R2, None, seed=5)
    assert np.isfinite(g12).all()
    scale = np.abs(g12).max()
    assert np.abs(g12 - (g1 + g2)).max() <= 2e-5 * scale
    # clip 0 alone (different run partition; the in-kernel noise stream is keyed by the clip index,
    # so only clip 0 keeps its key) gives the same gradient up to fp32 rounding
    solo0 = run_backward(d['ctrl'][0:1], d['f0_frames'][0:1], R1[0:1], None, seed=5)
    assert np.abs(solo0[0] - g1[0]).max() <= 2e-6 * scale


def test_module_trains_like_the_port():
    """End to end: CombSubFast(..., infer=False) under autograd gives parameter gradients that match
    the same control network followed by the stock-PyTorch restatement of the synthesizer."""
    from oracle import torch_port as TP
    torch.manual_seed(0)
    model = CombSubFast(44100, 512, n_unit=16, n_spk=2).cuda()
    B, F = 2, 40
    d = make_inputs(B, F, 1539, seed=34)
    units = torch.randn(B, F, 16, device='cuda')
    f0 = dev(d['f0_frames'])[..., None]
    vol = torch.rand(B, F, device='cuda')
    spk = torch.ones(B, 1, dtype=torch.long, device='cuda')
    U, R = dev(d['U']), torch.randn(B, F * 512, device='cuda')

    model.zero_grad()
    signal, _, (s_h, s_n) = model(units, f0, vol, spk, infer=False, noise_u=U)
    assert signal.requires_grad and s_h is signal
    (signal * R).sum().backward()
    ours = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    assert len(ours) > 10

    model.zero_grad()
    pf, _, _ = core.phase_stage(f0, 512, 44100, None, True)
    ctrls = model.unit2ctrl(units, f0, pf, vol, spk)
    sig, _ = TP.combsubfast_forward(ctrls['harmonic_magnitude'], ctrls['harmonic_phase'], ctrls['noise_magnitude'],
                                    f0, model.window, U)
    (sig * R).sum().backward()
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        scale = p.grad.abs().max().item() + 1e-12
        assert (ours[k] - p.grad).abs().max().item() <= 2e-3 * scale, k

    # an optimiser step on our gradients lowers the loss of a tiny fitting problem
    target = torch.randn(B, F * 512, device='cuda') * 0.05
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        out, _, _ = model(units, f0, vol, spk, infer=False, noise_u=U)
        loss = ((out - target) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses


@pytest.mark.parametrize('model', ['combsub', 'sins'])
def test_filter_models_train(golden_dir, model):
    """solver.py:111-113 with the Sins / CombSub-old drop-ins: forward through the kernels, gradients of the control
    tensors from ddsp_b200.diffsynth -- checked against the gradients autograd gave through the reference module."""
    from ddsp_b200 import vocoder
    g = dict(np.load(os.path.join(golden_dir, f'{model}_grad_small.npz')))
    splits = {'combsub': (256, 512, 256), 'sins': (128, 256, 256)}[model]
    names = {'combsub': ('group_delay', 'harmonic_magnitude', 'noise_magnitude'),
             'sins': ('amplitudes', 'group_delay', 'noise_magnitude')}[model]
    ctrl = dev(g['ctrl']).requires_grad_(True)

    class Fixed(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.dummy = torch.nn.Parameter(torch.zeros(1))

        def forward(self, *a, **k):
            return dict(zip(names, torch.split(ctrl, list(splits), dim=-1)))
    if model == 'combsub':
        m = vocoder.CombSub(44100, 512, 256, 512, 256, n_unit=4, unit2ctrl=Fixed()).cuda()
    else:
        m = vocoder.Sins(44100, 512, 128, 256, 256, n_unit=4, unit2ctrl=Fixed()).cuda()
    B, Fr = g['f0_frames'].shape
    sig, _, (harm, noise) = m(torch.zeros(B, Fr, 4, device='cuda'), dev(g['f0_frames'])[..., None], torch.zeros(B, Fr, device='cuda'),
                              torch.ones(B, 1, dtype=torch.long, device='cuda'), infer=False, noise_u=dev(g['U']))
    assert sig.requires_grad
    assert np.abs(sig.detach().cpu().numpy() - g['signal64']).max() < 5e-5
    (sig * dev(g['R'])).sum().backward()
    ref = g['grad64']
    err = np.abs(ctrl.grad.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err < 1e-4, err


def test_filter_model_with_control_network_takes_an_optimiser_step():
    from ddsp_b200.vocoder import Sins
    from ddsp_b200.loss import RSSLoss
    torch.manual_seed(0)
    m = Sins(44100, 512, 128, 256, 256, n_unit=8).cuda()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    loss_fn = RSSLoss(256, 1024, 2, device='cuda')
    units = torch.randn(2, 12, 8, device='cuda')
    f0 = torch.full((2, 12, 1), 220.0, device='cuda')
    vol = torch.rand(2, 12, device='cuda')
    spk = torch.ones(2, 1, dtype=torch.long, device='cuda')
    target = torch.randn(2, 12 * 512, device='cuda') * 0.05
    sig, _, _ = m(units, f0, vol, spk, infer=False)
    loss = loss_fn(sig, target)
    opt.zero_grad()
    loss.backward()
    n_grad = sum(1 for p in m.parameters() if p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0)
    assert n_grad > 50
    opt.step()


def test_backward_error_cases():
    d = make_inputs(1, 4, 1539, seed=35)
    hm, hp, nm = ctrl_views(d['ctrl'], 'combsubfast')
    f0 = dev(d['f0_frames'])[..., None]
    _, prefix, _ = core.phase_stage(f0, 512, 44100, None, True)
    with pytest.raises(ValueError):
        core.combsubfast_backward_stage(torch.zeros(1, 100, device='cuda'), hm, hp, nm, f0, prefix, 512, 44100)
    with pytest.raises(ValueError):
        core.combsubfast_backward_stage(torch.zeros(1, 2048, device='cuda'), hm[..., :500], hp[..., :500],
                                        nm[..., :500], f0, prefix, 512, 44100)
