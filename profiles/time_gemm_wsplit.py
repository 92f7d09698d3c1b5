"""Is the N = 256 GEMM bound by the bytes its CTAs pull from L2?  (run on the GPU box)
Pre-split weights cost (16 + 64) KB per k-block and CTA, raw weights split in the kernel (16 + 32) KB."""
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
    print(name, round(e0.elapsed_time(e1) / n * 1e3, 1), 'us', flush=True)
for K, N in ((512, 256), (768, 256), (256, 1024), (256, 1539)):
    x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') * K ** -0.5; b = torch.randn(N, device='cuda')
    hi, lo = core.split_tf32(w)
    out = torch.empty(M, N, device='cuda')
    timeit(lambda: core.linear_ex(x, hi, b, out=out, weight_lo=lo), f'K={K} N={N} pre-split W')
    timeit(lambda: core.linear_ex(x, w, b, out=out), f'K={K} N={N} raw W (split in kernel)')
