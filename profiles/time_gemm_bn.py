"""GEMM tile-width experiment for the N = 256 layers of the control network (run on the GPU box):
DDSP_B200_GEMM_BN=128 forces 128-column tiles (3 pipeline stages instead of 2) on calls without a fused LayerNorm."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200 import core
torch.manual_seed(0)
M = 64 * 862
def timeit(fn, name, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get('DDSP_B200_GEMM_BN', 'auto'), name, round(e0.elapsed_time(e1) / n * 1e3, 1), 'us', flush=True)
for K in (512, 768):
    x = torch.randn(M, K, device='cuda'); w = torch.randn(256, K, device='cuda') * K ** -0.5; b = torch.randn(256, device='cuda')
    hi, lo = core.split_tf32(w)
    res = torch.randn(M, 256, device='cuda'); out = torch.empty(M, 256, device='cuda')
    g = torch.ones(256, device='cuda'); be = torch.zeros(256, device='cuda')
    timeit(lambda: core.linear_ex(x, hi, b, out=out, weight_lo=lo), f'K={K} N=256 bias')
    timeit(lambda: core.linear_ex(x, hi, b, residual=res, out=out, weight_lo=lo), f'K={K} N=256 bias+residual')
    timeit(lambda: core.linear_ex(x, hi, b, residual=res, out=out, weight_lo=lo, ln=(g, be, 1e-5)), f'K={K} N=256 bias+residual+LN')
    timeit(lambda: torch.nn.functional.layer_norm(out, (256,), g, be), 'torch layer_norm alone')
