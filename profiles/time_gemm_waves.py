"""Is the residual + LayerNorm epilogue of the N = 256 GEMM a per-tile cost or a tail?  (run on the GPU box)
M = w * 148 * 128 rows = exactly w tiles per CTA; time against w for the three epilogue forms."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200 import core
torch.manual_seed(0)
K = 512
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
w_ = torch.randn(256, K, device='cuda') * K ** -0.5; b = torch.randn(256, device='cuda')
hi, lo = core.split_tf32(w_)
g = torch.ones(256, device='cuda'); be = torch.zeros(256, device='cuda')
for waves in (1, 2, 3, 4, 6):
    M = waves * 148 * 128
    x = torch.randn(M, K, device='cuda'); res = torch.randn(M, 256, device='cuda'); out = torch.empty(M, 256, device='cuda')
    t0 = timeit(lambda: core.linear_ex(x, hi, b, out=out, weight_lo=lo))
    t1 = timeit(lambda: core.linear_ex(x, hi, b, residual=res, out=out, weight_lo=lo))
    t2 = timeit(lambda: core.linear_ex(x, hi, b, residual=res, out=out, weight_lo=lo, ln=(g, be, 1e-5)))
    print(f'tiles per CTA {waves}: bias {t0:6.1f} us, + residual {t1:6.1f} us, + LayerNorm {t2:6.1f} us', flush=True)
