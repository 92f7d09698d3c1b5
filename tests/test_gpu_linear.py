"""Tensor-core Linear (csrc/gemm_tc.cuh, 3xTF32 on tcgen05) against an fp64 reference of the same op."""
import numpy as np
import pytest

from tests.gpu_util import HAS_CUDA, torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not HAS_CUDA, reason='needs a CUDA device')]


def _ref(x, w, b, r):
    y = x.double() @ w.double().t()
    if b is not None:
        y = y + b.double()
    if r is not None:
        y = y + r.double()
    return y


@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (300, 256, 256), (1000, 1539, 256), (862 * 3, 1536, 256),
                                   (517, 256, 512), (129, 1024, 256), (64, 128, 768), (5, 16, 32)])
@pytest.mark.parametrize('with_bias,with_res', [(False, False), (True, True)])
def test_linear_matches_fp64(M, N, K, with_bias, with_res):
    from ddsp_b200 import core
    g = torch.Generator(device='cuda').manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, device='cuda', generator=g)
    w = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g) if with_bias else None
    r = torch.randn(M, N, device='cuda', generator=g) if with_res else None
    y = core.linear(x, w, b, r)
    ref = _ref(x, w, b, r)
    scale = ref.abs().max().item()
    err = (y.double() - ref).abs().max().item()
    # fp32 GEMM (cuBLAS, no TF32) on the same data for comparison of the error level
    y32 = torch.nn.functional.linear(x, w, b) + (r if r is not None else 0)
    err32 = (y32.double() - ref).abs().max().item()
    assert err <= 4e-6 * scale + 4 * err32, (err, err32, scale)


def test_linear_strided_views_and_padded_output():
    """A is a column slice of a wider tensor (row stride > K); the output is a padded buffer (ldc = 1540) as the control
    network allocates it for the 1539-wide projection; the residual aliases the output."""
    from ddsp_b200 import core
    g = torch.Generator(device='cuda').manual_seed(5)
    big = torch.randn(700, 1024, device='cuda', generator=g)
    x = big[:, 256:512]
    w = torch.randn(1539, 256, device='cuda', generator=g) / 16
    buf = torch.zeros(700, 1540, device='cuda')
    out = buf[:, :1539]
    core.linear(x, w, out=out)
    ref = x.double() @ w.double().t()
    assert (out.double() - ref).abs().max().item() < 2e-5
    assert buf[:, 1539].abs().max().item() == 0.0
    acc = torch.randn(700, 256, device='cuda', generator=g)
    w2 = torch.randn(256, 256, device='cuda', generator=g) / 16
    expect = acc.double() + x.double() @ w2.double().t()
    core.linear(x, w2, residual=acc, out=acc)
    assert (acc.double() - expect).abs().max().item() < 2e-5


def test_linear_batched_leading_dims_and_errors():
    from ddsp_b200 import core
    x = torch.randn(3, 50, 256, device='cuda')
    w = torch.randn(512, 256, device='cuda') / 16
    b = torch.randn(512, device='cuda')
    y = core.linear(x, w, b)
    assert y.shape == (3, 50, 512)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    assert (y.double() - ref).abs().max().item() < 2e-5
    with pytest.raises(ValueError):
        core.linear(torch.randn(4, 100, device='cuda'), w)
