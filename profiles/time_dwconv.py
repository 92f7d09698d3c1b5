#!/usr/bin/env python
"""Time dwconv_silu_kernel (depthwise Conv1d(31) + SiLU of the conformer module) at the headline batch for every
library in lib/variants/ (DDSP_B200_LIB), and check each variant against the first."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, 'ddsp-svc-official_b200', 'lib', 'variants')
code = r'''
import sys, os, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
from ddsp_b200 import core
torch.manual_seed(0)
B, N = 64, 862
g = torch.randn(B, N, 512, device='cuda'); w = torch.randn(512, 1, 31, device='cuda') * 0.2; bias = torch.randn(512, device='cuda')
out = core.dwconv_silu(g, w, bias)
ref_path = sys.argv[2]
res = {'variant': sys.argv[1]}
o = out[:4].cpu().numpy()
if os.path.exists(ref_path): res['max_abs_vs_first'] = float(np.abs(o - np.load(ref_path)).max())
else: np.save(ref_path, o)
# bytes in flight exceed L2 between repeats: 113 MB in + 113 MB out per call
for _ in range(3): core.dwconv_silu(g, w, bias)
torch.cuda.synchronize()
ts = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): core.dwconv_silu(g, w, bias)
    e1.record(); torch.cuda.synchronize()
    ts.append(round(e0.elapsed_time(e1) / 20 * 1e3, 1))
res['us'] = ts
print(json.dumps(res))
'''
ref = '/tmp/dwconv_ref.npy'
if os.path.exists(ref): os.remove(ref)
for n in sorted(f[:-3] for f in os.listdir(VAR) if f.endswith('.so')):
    env = dict(os.environ, DDSP_B200_LIB=os.path.join(VAR, n + '.so'))
    r = subprocess.run([sys.executable, '-c', code, n, ref], capture_output=True, text=True, env=env, cwd=ROOT)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else 'FAILED ' + n + ' ' + r.stderr[-600:], flush=True)
