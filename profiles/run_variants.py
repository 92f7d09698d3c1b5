#!/usr/bin/env python
"""Time the CombSubFast stage-B kernel of every library in ddsp-svc-official_b200/lib/variants/ on the
headline shape (B=64 x 10 s) and compare each variant's output with the first one (injected noise and
in-kernel noise).  One subprocess per variant (DDSP_B200_LIB).  Prints one JSON line per variant."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, 'ddsp-svc-official_b200', 'lib', 'variants')
code = r'''
import sys, os, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
from ddsp_b200 import core
from ddsp_b200.synthetic import make_inputs
name, ref_path = sys.argv[1], sys.argv[2]
res = {'variant': name}
# correctness: small shape with unvoiced frames, injected and in-kernel noise
d = make_inputs(6, 301, 1539, seed=5, zero_f0_fraction=0.1)
ctrl = torch.from_numpy(d['ctrl']).cuda(); hm, hp, nm = torch.split(ctrl, 513, dim=-1)
f0 = torch.from_numpy(d['f0_frames']).cuda()[..., None]
pf, prefix, _ = core.phase_stage(f0, 512, 44100)
a = core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, noise_u=torch.from_numpy(d['U']).cuda()).cpu().numpy()
b = core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, seed=3).cpu().numpy()
if os.path.exists(ref_path):
    r = np.load(ref_path)
    res['max_abs_vs_first_injected'] = float(np.abs(a - r['a']).max())
    res['max_abs_vs_first_inkernel'] = float(np.abs(b - r['b']).max())
else:
    np.savez(ref_path, a=a, b=b)
try:
    from oracle import ddsp_oracle as O
    ref, _ = O.combsubfast_forward(d['ctrl'][:2, :, :513], d['ctrl'][:2, :, 513:1026], d['ctrl'][:2, :, 1026:], d['f0_frames'][:2], d['U'][:2])
    res['max_abs_vs_oracle'] = float(np.abs(a[:2] - ref).max())
except Exception as e:
    res['oracle_error'] = repr(e)[:200]
# timing: headline shape
B, F = 64, 862
d = make_inputs(B, F, 1539, seed=1234, noise=False)
ctrl = torch.from_numpy(d['ctrl']).cuda(); hm, hp, nm = torch.split(ctrl, 513, dim=-1)
f0 = torch.from_numpy(d['f0_frames']).cuda()[..., None]
win = torch.sqrt(torch.hann_window(1024)).cuda()
pf, prefix, _ = core.phase_stage(f0, 512, 44100)
for i in range(5):
    core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, seed=i, window=win)
torch.cuda.synchronize()
ts = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, seed=i, window=win)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 50 * 1e3)
res['stage_b_us'] = [round(t, 1) for t in ts]
print(json.dumps(res))
'''
ref = '/tmp/variant_ref.npz'
if os.path.exists(ref):
    os.remove(ref)
names = sys.argv[1:] or sorted(f[:-3] for f in os.listdir(VAR) if f.endswith('.so'))
for n in names:
    env = dict(os.environ, DDSP_B200_LIB=os.path.join(VAR, n + '.so'))
    r = subprocess.run([sys.executable, '-c', code, n, ref], capture_output=True, text=True, env=env, cwd=ROOT)
    print(r.stdout.strip() or ('FAILED ' + n + ' ' + r.stderr[-600:]), flush=True)
