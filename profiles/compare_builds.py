"""Runs CombSubFast, CombSub-old and Sins launches with two builds of the library (DDSP_B200_LIB) in subprocesses
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
outs = [a, b, pf, prefix]
for model, K in (('combsub', 1024), ('sins', 640)):
    dm = make_inputs(8, 301, K, seed=6, zero_f0_fraction=0.1)
    cm = torch.from_numpy(dm['ctrl']).cuda()
    fm = torch.from_numpy(dm['f0_frames']).cuda()[..., None]
    U = torch.from_numpy(dm['U']).cuda()
    pfm, prem, phm = core.phase_stage(fm, 512, 44100, full_rate=(model == 'sins'))
    if model == 'combsub':
        gd, hm2, nm2 = torch.split(cm, [256, 512, 256], dim=-1)
        outs += list(core.combsub_stage(gd, hm2, nm2, fm, prem, 512, 44100, noise_u=U))
        outs += list(core.combsub_stage(gd, hm2, nm2, fm, prem, 512, 44100, seed=9))
    else:
        am, gd, nm2 = torch.split(cm, [128, 256, 256], dim=-1)
        outs += list(core.sins_stage(am, gd, nm2, fm, phm, 512, 44100, noise_u=U))
torch.cuda.synchronize()
print(' '.join(hashlib.sha1(t.cpu().numpy().tobytes()).hexdigest()[:16] for t in outs))
'''
outs = []
for lib in sys.argv[1:3]:
    env = dict(os.environ, DDSP_B200_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env)
    print(lib, r.stdout.strip() or r.stderr[-300:])
    outs.append(r.stdout.strip())
print('bit-identical' if outs[0] == outs[1] and outs[0] else 'DIFFERENT')
