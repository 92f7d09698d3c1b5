"""Times of the tensor-core Linear at the control network's shapes (M = 64 x 862 rows)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddsp_b200 import core
M = 55168
rows = []
for name, N, K in (('qkv-like', 1536, 256), ('to_out', 256, 512), ('pw1-like', 1024, 256), ('final', 1539, 256)):
    x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') / K ** 0.5
    hi, lo = core.split_tf32(w)
    out = torch.empty(M, (N + 3) // 4 * 4, device='cuda')[:, :N]
    for tag, fn in (('presplit', lambda: core.linear_ex(x, hi, out=out, weight_lo=lo)), ('insplit', lambda: core.linear_ex(x, w, out=out))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        rows.append({'layer': name, 'N': N, 'K': K, 'mode': tag, 'ms': ms, 'tflops_fp32_equiv': 2 * M * N * K / ms / 1e9})
print(json.dumps({'bn': os.environ.get('DDSP_B200_GEMM_BN', 'auto'), 'rows': rows}))
