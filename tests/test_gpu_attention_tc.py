"""Second-generation tensor-core path of the control network (csrc/gemm_attn.cuh): linear_ex (TMA-store epilogue,
pre-split weights, fused LayerNorm) and the Performer attention as GEMMs, against fp64 restatements of
ddsp/pcmer.py:69-78, :124-160, :191-251 in plain torch ops."""
import math

import pytest

from tests.gpu_util import HAS_CUDA, torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not HAS_CUDA, reason='needs a CUDA device')]


@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (300, 256, 256), (1000, 1539, 256), (517, 256, 512), (129, 1024, 256),
                                   (5, 16, 32), (862 * 2 + 7, 512, 256)])
@pytest.mark.parametrize('presplit', [False, True])
@pytest.mark.parametrize('with_bias,with_res', [(False, False), (True, True)])
def test_linear_ex_matches_fp64(M, N, K, presplit, with_bias, with_res):
    from ddsp_b200 import core
    g = torch.Generator(device='cuda').manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, device='cuda', generator=g)
    w = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g) if with_bias else None
    r = torch.randn(M, N, device='cuda', generator=g) if (with_res and N % 32 == 0) else None
    ldc = (N + 3) // 4 * 4
    buf = torch.full((M, ldc), 7.0, device='cuda')
    if presplit:
        hi, lo = core.split_tf32(w)
        assert torch.equal(hi + lo, w)
        y = core.linear_ex(x, hi, b, r, out=buf[:, :N], weight_lo=lo)
    else:
        y = core.linear_ex(x, w, b, r, out=buf[:, :N])
    ref = x.double() @ w.double().t()
    if b is not None:
        ref = ref + b.double()
    if r is not None:
        ref = ref + r.double()
    scale = ref.abs().max().item()
    err = (y.double() - ref).abs().max().item()
    assert err <= 6e-6 * scale, (err, scale)
    if ldc > N:
        # the TMA store clips at the tensor edge in 16-byte granules: the row padding up to the next multiple of four
        # floats is either left alone or overwritten with the (zero) accumulator of the out-of-range weight rows
        pad = buf[:, N:]
        assert ((pad == 7.0) | (pad == 0.0)).all()


@pytest.mark.parametrize('M', [64, 1000])
def test_linear_ex_fused_layer_norm(M):
    from ddsp_b200 import core
    g = torch.Generator(device='cuda').manual_seed(M)
    x = torch.randn(M, 512, device='cuda', generator=g)
    w = torch.randn(256, 512, device='cuda', generator=g) / 512 ** 0.5
    b = torch.randn(256, device='cuda', generator=g)
    r = torch.randn(M, 256, device='cuda', generator=g) * 3 + 1.5
    gamma = torch.randn(256, device='cuda', generator=g)
    beta = torch.randn(256, device='cuda', generator=g)
    y, yn = core.linear_ex(x, w, b, r, ln=(gamma, beta, 1e-5))
    ref = x.double() @ w.double().t() + b.double() + r.double()
    refn = torch.nn.functional.layer_norm(ref, (256,), gamma.double(), beta.double(), 1e-5)
    # values up to ~13; the fp32 TMEM accumulator rounds once per MMA (3 x K/8 accumulations): a few 1e-6 relative
    assert (y.double() - ref).abs().max().item() < 4e-5
    assert (yn.double() - refn).abs().max().item() < 2e-5
    # in place on the residual stream, as the control network calls it
    r2 = r.clone()
    y2, yn2 = core.linear_ex(x, w, b, r2, out=r2, ln=(gamma, beta, 1e-5))
    assert torch.equal(y2, y) and torch.equal(yn2, yn)


def _favor_reference(x, wq, wk, wv, bq, bk, bv, proj, heads, eps=1e-4):
    """pcmer.py:191-251 (to_q/k/v, softmax_kernel features, linear_attention) in fp64."""
    x = x.double()
    b, n, _ = x.shape
    split = lambda t: t.view(b, n, heads, 64).transpose(1, 2)      # noqa: E731
    q = split(x @ wq.double().t() + bq.double())
    k = split(x @ wk.double().t() + bk.double())
    v = split(x @ wv.double().t() + bv.double())
    P = proj.double()
    scale = 64 ** -0.25
    ratio = P.shape[0] ** -0.5

    def feat(t, is_q):
        dash = torch.einsum('bhnd,jd->bhnj', scale * t, P)
        diag = (t * t).sum(-1, keepdim=True) * (0.5 * scale * scale)
        if is_q:
            return ratio * (torch.exp(dash - diag - dash.amax(dim=-1, keepdim=True)) + eps)
        return ratio * torch.exp(dash - diag + eps)
    qf, kf = feat(q, True), feat(k, False)
    d_inv = 1.0 / (torch.einsum('bhnj,bhj->bhn', qf, kf.sum(dim=-2)) + 1e-8)
    ctx = torch.einsum('bhnj,bhne->bhje', kf, v)
    out = torch.einsum('bhje,bhnj,bhn->bhne', ctx, qf, d_inv)
    return out.transpose(1, 2).reshape(b, n, heads * 64)


# (12, 862): 672 feature tiles = 4..5 per CTA, every slot of the accumulator ring and both barrier parities come round
@pytest.mark.parametrize('B,F', [(1, 300), (3, 862), (2, 129), (5, 37), (12, 862)])
@pytest.mark.parametrize('presplit', [False, True])
def test_favor_attention_matches_fp64(B, F, presplit):
    from ddsp_b200 import core
    from ddsp_b200.control import _orthogonal_gaussian_features
    H = 8
    g = torch.Generator(device='cuda').manual_seed(B * 1000 + F)
    torch.manual_seed(B * 1000 + F)
    x = torch.randn(B, F, 256, device='cuda', generator=g)
    ws = [torch.randn(512, 256, device='cuda', generator=g) / 16 for _ in range(3)]
    bs = [torch.randn(512, device='cuda', generator=g) * 0.1 for _ in range(3)]
    proj = _orthogonal_gaussian_features(266, 64).cuda()
    w = torch.cat(ws).contiguous()
    bias = torch.cat(bs).contiguous()
    ps = (64 ** -0.25 * proj).contiguous()
    for rep in range(2):        # the second call reuses the cached workspace (pad rows must still be zero)
        if presplit:
            hi, lo = core.split_tf32(w)
            out = core.favor_attention(x, hi, lo, bias, ps, H)
        else:
            out = core.favor_attention(x, w, None, bias, ps, H)
    ref = _favor_reference(x, *ws, *bs, proj, H)
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 3e-5 * max(scale, 1.0), (err, scale)


def test_favor_attention_large_arguments():
    """Queries / keys with |x| ~ 6 (dash up to ~ +-20): the stabilised query features and the exp of the keys stay
    finite and accurate."""
    from ddsp_b200 import core
    from ddsp_b200.control import _orthogonal_gaussian_features
    torch.manual_seed(3)
    B, F, H = 2, 200, 8
    x = torch.randn(B, F, 256, device='cuda')
    ws = [torch.randn(512, 256, device='cuda') / 16 * s for s in (0.8, 0.8, 1.0)]
    bs = [torch.zeros(512, device='cuda') for _ in range(3)]
    proj = _orthogonal_gaussian_features(266, 64).cuda()
    out = core.favor_attention(x, torch.cat(ws).contiguous(), None, torch.cat(bs), (64 ** -0.25 * proj).contiguous(), H)
    ref = _favor_reference(x, *ws, *bs, proj, H)
    assert torch.isfinite(out).all()
    rel = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    assert rel < 1e-4, rel


@pytest.mark.parametrize('B,T', [(2, 300), (1, 33), (3, 862)])
def test_linear_glu_and_dwconv_silu_match_the_module(B, T):
    """pcmer.py:52-55: Conv1d(256 -> 1024, 1) -> GLU -> depthwise Conv1d(k=31, same) -> SiLU, channels-last."""
    from ddsp_b200 import core
    torch.manual_seed(B * 100 + T)
    x = torch.randn(B, T, 256, device='cuda')
    pw1 = torch.nn.Conv1d(256, 1024, 1).cuda()
    dw = torch.nn.Conv1d(512, 512, 31, padding='same', groups=512).cuda()
    with torch.no_grad():
        ref_u = torch.nn.functional.linear(x.double(), pw1.weight.double().squeeze(-1), pw1.bias.double())
        ref_g = torch.nn.functional.glu(ref_u, dim=-1).transpose(1, 2)
        ref = torch.nn.functional.silu(torch.nn.functional.conv1d(ref_g, dw.weight.double(), dw.bias.double(), padding=15, groups=512))
        w_il, b_il = core.glu_interleave(pw1.weight.detach(), pw1.bias.detach())
        hi, lo = core.split_tf32(w_il)
        for g in (core.linear_glu(x, w_il, b_il), core.linear_glu(x, hi, b_il, weight_lo=lo)):
            assert (g.double() - ref_g.transpose(1, 2)).abs().max().item() < 2e-5
        y = core.dwconv_silu(g, dw.weight, dw.bias)
    assert (y.double() - ref.transpose(1, 2)).abs().max().item() < 3e-5
