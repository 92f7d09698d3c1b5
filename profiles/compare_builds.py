"""Runs the headline CombSubFast launch with two builds of the library (DDSP_B200_LIB) in subprocesses
and reports whether the outputs are bit-identical (used when a kernel edit is meant to be a pure
re-scheduling).  usage: python profiles/compare_builds.py libA.so libB.so"""
import hashlib, os, subprocess, sys
code = r'''
import sys, os, hashlib, torch
sys.path.insert(0, os.getcwd())
from ddsp_b200 import core
from ddsp_b200.synthetic import make_inputs
d = make_inputs(8, 301, 1539, seed=5, zero_f0_fraction=0.1)
ctrl = torch.from_numpy(d['ctrl']).cuda(); hm, hp, nm = torch.split(ctrl, 513, dim=-1)
f0 = torch.from_numpy(d['f0_frames']).cuda()[..., None]
pf, prefix, _ = core.phase_stage(f0, 512, 44100)
a = core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, seed=3)
b = core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, noise_u=torch.from_numpy(d['U']).cuda())
torch.cuda.synchronize()
print(' '.join(hashlib.sha1(t.cpu().numpy().tobytes()).hexdigest()[:16] for t in (a, b, pf, prefix)))
'''
outs = []
for lib in sys.argv[1:3]:
    env = dict(os.environ, DDSP_B200_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env)
    print(lib, r.stdout.strip() or r.stderr[-300:])
    outs.append(r.stdout.strip())
print('bit-identical' if outs[0] == outs[1] and outs[0] else 'DIFFERENT')
