"""Fused stages of the control network (csrc/control.cuh) against the plain PyTorch ops they replace
(ddsp_b200/control.py, itself checked against the reference modules in tests/test_control.py)."""
import numpy as np
import pytest

from tests.gpu_util import HAS_CUDA, torch

pytestmark = pytest.mark.gpu

if HAS_CUDA:
    import torch.nn.functional as F
    from ddsp_b200 import core
    from ddsp_b200.control import Unit2Control, _softmax_features


@pytest.mark.parametrize('B,N,H', [(1, 1, 8), (2, 37, 8), (3, 130, 4)])
@pytest.mark.parametrize('is_query', [True, False])
def test_performer_features(B, N, H, is_query):
    torch.manual_seed(B * 100 + N)
    x = torch.randn(B, N, H * 64, device='cuda') * 1.5
    proj = torch.randn(266, 64, device='cuda')
    ref = _softmax_features(x.view(B, N, H, 64).transpose(1, 2), proj, is_query)        # (B,H,N,266)
    dash = torch.matmul((64 ** -0.25 * x).view(-1, 64), proj.t())
    out = core.performer_features(dash, x, H, is_query)
    assert out.shape == ref.shape
    assert torch.isfinite(out).all()
    assert ((out - ref).abs() <= 1e-4 * ref.abs() + 1e-9).all(), (out - ref).abs().max().item()   # exp() of args up to ~20


@pytest.mark.parametrize('B,N,H', [(1, 1, 8), (2, 37, 8), (3, 131, 4), (64, 100, 8)])
@pytest.mark.parametrize('is_query', [True, False])
def test_performer_project_features(B, N, H, is_query):
    torch.manual_seed(B * 100 + N + 1)
    x = torch.randn(B, N, H * 64, device='cuda') * 1.5
    proj = torch.randn(266, 64, device='cuda')
    ref = _softmax_features(x.view(B, N, H, 64).transpose(1, 2).double(), proj.double(), is_query)
    out = core.performer_project_features(x, proj, H, is_query)
    assert out.shape == ref.shape and torch.isfinite(out).all()
    assert ((out.double() - ref).abs() <= 1e-4 * ref.abs() + 1e-9).all(), (out.double() - ref).abs().max().item()


def test_bias_folded_into_the_fused_stages():
    """x_bias / u_bias: the bias of the producing Linear added on load equals adding it beforehand."""
    torch.manual_seed(3)
    x = torch.randn(2, 50, 8 * 64, device='cuda')
    bias = torch.randn(8 * 64, device='cuda') * 0.3
    proj = torch.randn(266, 64, device='cuda')
    for is_query in (True, False):
        a = core.performer_project_features(x, proj, 8, is_query, x_bias=bias)
        b = core.performer_project_features(x + bias, proj, 8, is_query)
        assert ((a - b).abs() <= 1e-5 * b.abs() + 1e-9).all()
    u = torch.randn(2, 70, 1024, device='cuda')
    ub = torch.randn(1024, device='cuda') * 0.3
    w = torch.randn(512, 1, 31, device='cuda') * 0.2
    db = torch.randn(512, device='cuda')
    assert torch.equal(core.glu_dwconv_silu(u, w, db, u_bias=ub), core.glu_dwconv_silu(u + ub, w, db))


@pytest.mark.parametrize('B,T,C', [(1, 5, 512), (2, 64, 512), (2, 131, 512), (1, 300, 96)])
def test_glu_dwconv_silu(B, T, C):
    torch.manual_seed(T)
    u = torch.randn(B, T, 2 * C, device='cuda')
    w = torch.randn(C, 1, 31, device='cuda') * 0.2
    b = torch.randn(C, device='cuda')
    ref = F.silu(F.conv1d(F.glu(u.transpose(1, 2).double(), dim=1), w.double(), b.double(), padding=15, groups=C))
    out = core.glu_dwconv_silu(u, w, b)
    assert out.shape == (B, T, C)
    assert (out.double() - ref.transpose(1, 2)).abs().max().item() < 5e-6


# (16, 520) takes the fused projection; (24, 862): 10-second clips, every tensor-core kernel runs several tiles per CTA
@pytest.mark.parametrize('B,F_', [(1, 9), (2, 130), (3, 77), (16, 520), (24, 862)])
def test_unit2control_fused_matches_plain_ops(B, F_):
    """no_grad on CUDA takes the fused stages; under enable_grad the same module runs stock ops."""
    torch.manual_seed(7)
    splits = {'harmonic_magnitude': 513, 'harmonic_phase': 513, 'noise_magnitude': 513}
    net = Unit2Control(32, 3, splits).cuda().eval()
    units = torch.randn(B, F_, 32, device='cuda')
    f0 = torch.rand(B, F_, 1, device='cuda') * 500 + 80
    ph = (torch.rand(B, F_, device='cuda') - 0.5) * 6
    vol = torch.rand(B, F_, device='cuda')
    spk = torch.full((B, 1), 2, dtype=torch.long, device='cuda')
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the stock 1x1 convs would otherwise run in TF32
    try:
        with torch.enable_grad():
            ref = net(units, f0, ph, vol, spk)
        with torch.no_grad():
            out = net(units, f0, ph, vol, spk)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    for k in splits:
        assert out[k].shape == ref[k].shape and not out[k].requires_grad
        assert (out[k] - ref[k].detach()).abs().max().item() < 5e-5, k
    # views of one tensor, as the reference's split_to_dict emits
    assert out['harmonic_phase'].data_ptr() == out['harmonic_magnitude'].data_ptr() + 513 * 4


@pytest.mark.parametrize('B,N,H', [(1, 1, 8), (1, 9, 8), (1, 16, 8), (1, 17, 8), (2, 26, 8), (1, 130, 8), (3, 37, 4), (2, 862, 8)])
def test_performer_attention_fused(B, N, H):
    """Fused attention (one kernel up to 16 frames, three tiled kernels beyond) against the plain PyTorch
    formulation (pcmer.py:69-78,124-160); the tiled path sums its partial contexts in a fixed order."""
    from ddsp_b200.control import _FastAttention
    torch.manual_seed(N)
    q, k, v = (torch.randn(B, N, H * 64, device='cuda') for _ in range(3))
    qb, kb, vb = (torch.randn(H * 64, device='cuda') * 0.2 for _ in range(3))
    fa = _FastAttention(64).cuda()
    split = lambda t: t.view(B, N, H, 64).transpose(1, 2).double()      # noqa: E731
    fa64 = _FastAttention(64).cuda().double()
    fa64.projection_matrix.copy_(fa.projection_matrix.double())
    ref = fa64(split(q + qb), split(k + kb), split(v + vb)).transpose(1, 2).reshape(B, N, H * 64)
    out = core.performer_attention(q, k, v, fa.projection_matrix, H, qb, kb, vb)
    assert out.shape == ref.shape and torch.isfinite(out).all()
    assert (out.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-6
    out_nb = core.performer_attention(q + qb, k + kb, v + vb, fa.projection_matrix, H)
    assert (out_nb - out).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-6
    assert torch.equal(out, core.performer_attention(q, k, v, fa.projection_matrix, H, qb, kb, vb))      # deterministic


@pytest.mark.parametrize('N', [9, 40])
def test_performer_attention_takes_slices_of_a_merged_projection(N):
    torch.manual_seed(N)
    B, H = 2, 8
    proj = torch.randn(266, 64, device='cuda')
    qkv = torch.randn(B, N, 3 * H * 64, device='cuda')
    q, k, v = qkv.chunk(3, dim=-1)                               # frame stride 1536, no copies
    a = core.performer_attention(q, k, v, proj, H)
    b = core.performer_attention(q.contiguous(), k.contiguous(), v.contiguous(), proj, H)
    assert torch.equal(a, b)


def test_merged_qkv_weight_follows_weight_updates():
    from ddsp_b200.control import _SelfAttention
    att = _SelfAttention(256, 8).cuda()
    w0 = att._merged_qkv_weight()
    assert w0.shape == (1536, 256) and att._merged_qkv_weight() is w0
    with torch.no_grad():
        att.to_k.weight.add_(1.0)
    w1 = att._merged_qkv_weight()
    assert w1 is not w0 and torch.equal(w1[512:1024], att.to_k.weight)
    assert '_qkv_cache' not in att.state_dict()


@pytest.mark.parametrize('B,N,n_unit', [(1, 300, 256), (3, 101, 256), (2, 131, 768), (5, 64, 32)])
def test_prenet_on_tensor_cores_matches_stock_ops(B, N, n_unit):
    """unit_prenet as two overlapping-row GEMMs + GroupNorm/LeakyReLU kernel (channels last) against the stock
    Transpose - Conv1d - GroupNorm - LeakyReLU - Conv1d - Transpose in fp64 (unit2control.py:38-45)."""
    torch.manual_seed(B * 1000 + N)
    net = Unit2Control(n_unit, 1, {'a': 8}).cuda().eval()
    with torch.no_grad():
        net.unit_prenet[2].weight.uniform_(0.5, 1.5)
        net.unit_prenet[2].bias.uniform_(-0.5, 0.5)
    units = torch.randn(B, N, n_unit, device='cuda')
    units[:, :, ::7] += 3.0                                   # non-zero group means
    with torch.no_grad():
        out = net._prenet_tc(units)
        ref = net.unit_prenet.double()(units.double())
        net.float()
    assert out.shape == (B, N, 256) and out.stride(2) == 1
    assert (out.double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())


def test_embed_sum_with_fused_layer_norm():
    torch.manual_seed(12)
    net = Unit2Control(32, 4, {'a': 8}).cuda().eval()
    B, N = 3, 97
    xp = torch.randn(B, N + 2, 256, device='cuda')
    x = xp[:, :N]                                              # the strided view the pre-net hands over
    f0 = torch.rand(B, N, 1, device='cuda') * 500 + 80
    f0[1, 5:9] = 0.0
    ph = (torch.rand(B, N, device='cuda') - 0.5) * 6
    vol = torch.rand(B, N, device='cuda')
    ln = net.dec_post[0].net[0].norm
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.5, 0.5)
        for spk in (net.spk_embed.weight[1:2], net.spk_embed.weight[:3].unsqueeze(1)):
            a = core.embed_sum(x, f0, ph, vol, net.f0_embed, net.phase_embed, net.volume_embed, spk)
            b, bn = core.embed_sum_ln(x, f0, ph, vol, net.f0_embed, net.phase_embed, net.volume_embed, spk, ln)
            assert torch.equal(a, b)
            ref = torch.nn.functional.layer_norm(a.double(), (256,), ln.weight.double(), ln.bias.double(), ln.eps)
            assert (bn.double() - ref).abs().max().item() < 5e-6


def test_embed_sum_and_speaker_mix():
    torch.manual_seed(9)
    net = Unit2Control(16, 3, {'a': 513, 'b': 513, 'c': 513}).cuda().eval()
    B, N = 2, 33
    units = torch.randn(B, N, 16, device='cuda')
    f0 = torch.rand(B, N, 1, device='cuda') * 500 + 80
    f0[0, 3:6] = 0.0                                             # unvoiced frames: log(1 + 0)
    ph = (torch.rand(B, N, device='cuda') - 0.5) * 6
    vol = torch.rand(B, N, device='cuda')
    spk = torch.tensor([[1], [3]], device='cuda')
    for mix in (None, {1: 0.25, 3: 0.75}):
        with torch.enable_grad():
            ref = net(units, f0, ph, vol, spk, spk_mix_dict=mix)
        with torch.no_grad():
            out = net(units, f0, ph, vol, spk, spk_mix_dict=mix)
        for k in ref:
            assert (out[k] - ref[k].detach()).abs().max().item() < 1e-4, (k, mix)


def test_fused_forward_is_graph_capturable():
    torch.manual_seed(8)
    net = Unit2Control(16, 1, {'a': 513, 'b': 513, 'c': 513}).cuda().eval()
    args = (torch.randn(1, 26, 16, device='cuda'), torch.rand(1, 26, 1, device='cuda') * 300 + 100,
            torch.rand(1, 26, device='cuda'), torch.rand(1, 26, device='cuda'),
            torch.ones(1, 1, dtype=torch.long, device='cuda'))
    with torch.no_grad():
        eager = net(*args)['a'].clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            net(*args)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                out = net(*args)
        g.replay()
        torch.cuda.synchronize()
    assert np.allclose(out['a'].cpu().numpy(), eager.cpu().numpy(), atol=1e-6)
