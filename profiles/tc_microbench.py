#!/usr/bin/env python
"""Tensor-pipe measurements with the 3xTF32 tcgen05 kernel (csrc/gemm_tc.cuh):
 (1) the Linear shapes of the control network at the headline batch against cuBLAS fp32 and cuBLAS TF32;
 (2) a DFT-32 pass as a GEMM (N = K = 64: a batch of 32-point complex transforms as [re|im] x [[C,S],[-S,C]]) with
     the operands resident on chip -- the go / no-go number for a tensor-core FFT in the CombSubFast kernel
     (VERDICT r1 item 1c): 27 648 frame pairs x 3 FFT-1024 x 2 passes x 32 columns = 5.3 M rows of 64.
Prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ddsp_b200 import core


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {'linear': [], 'dft_pass': []}
M = 64 * 862
for (name, N, K) in [('qkv merged', 1536, 256), ('to_out', 256, 512), ('pw1', 1024, 256), ('pw2', 256, 512), ('final', 1539, 256)]:
    x = torch.randn(M, K, device='cuda')
    w = torch.randn(N, K, device='cuda') / K ** 0.5
    b = torch.randn(N, device='cuda')
    y = torch.empty(M, N, device='cuda')
    t_tc = timeit(lambda: core.linear(x, w, b, out=y))
    torch.backends.cuda.matmul.allow_tf32 = False
    t_f32 = timeit(lambda: torch.nn.functional.linear(x, w, b))
    torch.backends.cuda.matmul.allow_tf32 = True
    t_tf32 = timeit(lambda: torch.nn.functional.linear(x, w, b))
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = x[:4096].double() @ w.double().t() + b.double()
    e_tc = (core.linear(x[:4096], w, b).double() - ref).abs().max().item()
    e_f32 = (torch.nn.functional.linear(x[:4096], w, b).double() - ref).abs().max().item()
    torch.backends.cuda.matmul.allow_tf32 = True
    e_tf32 = (torch.nn.functional.linear(x[:4096], w, b).double() - ref).abs().max().item()
    torch.backends.cuda.matmul.allow_tf32 = False
    fl = 2.0 * M * N * K
    out['linear'].append({'layer': name, 'M': M, 'N': N, 'K': K, 'ms_3xtf32_tcgen05': t_tc, 'ms_cublas_fp32': t_f32,
                          'ms_cublas_tf32': t_tf32, 'tflops_fp32_equiv': fl / t_tc / 1e9,
                          'tensor_tflops_tf32_issued': 3 * fl / t_tc / 1e9,
                          'max_abs_err': {'3xtf32': e_tc, 'cublas_fp32': e_f32, 'cublas_tf32': e_tf32}})
# DFT-32 pass on resident operands: rows per launch = virtual_tiles * 128
rows = 27648 * 3 * 2 * 32
for bn in (64,):
    tiles = rows // 128
    t = timeit(lambda: core.tc_microbench(64, 64, bn, tiles), n=10)
    out['dft_pass'].append({'block_n': bn, 'rows': rows, 'ms_all_6_passes_of_all_pairs': t,
                            'note': '3xTF32 MMAs + TMA from L2 + split + TMEM drain, no HBM traffic, no twiddles / transposes'})
# the same DFT work as ONE wide GEMM tile shape (N=256 = four 64-column DFT blocks side by side is not a valid DFT,
# but shows the pipe rate at full MMA width for the same MAC count per row: K = 64)
t = timeit(lambda: core.tc_microbench(256, 64, 256, rows // 128 // 4), n=10)
out['dft_pass'].append({'block_n': 256, 'rows_equiv': rows, 'ms_same_macs_at_full_width': t})
print(json.dumps(out))
