#!/usr/bin/env python
"""A few launches of one synthesizer's stage A + stage B at the headline shape, for ncu:
    python profiles/prof_stage.py [combsubfast|combsub|sins] [iters]
Also prints the source hash of csrc/ so that a capture can be tied to the build it profiled."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ddsp_b200 import core
from ddsp_b200.synthetic import make_inputs
from profiles.sass_sections import source_hash
model = sys.argv[1] if len(sys.argv) > 1 else 'combsubfast'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
B, F = 64, 862
K = {'combsubfast': (513, 513, 513), 'combsub': (256, 512, 256), 'sins': (128, 256, 256)}[model]
d = make_inputs(B, F, sum(K), seed=1234, noise=False)
ctrl = torch.from_numpy(d['ctrl']).cuda()
c0, c1, c2 = torch.split(ctrl, list(K), dim=-1)
f0 = torch.from_numpy(d['f0_frames']).cuda()[..., None]
win = torch.sqrt(torch.hann_window(1024)).cuda()
for i in range(iters):
    pf, prefix, phase = core.phase_stage(f0, 512, 44100, full_rate=(model == 'sins'))
    if model == 'combsubfast':
        core.combsubfast_stage(c0, c1, c2, f0, prefix, 512, 44100, seed=i, window=win)
    elif model == 'combsub':
        core.combsub_stage(c0, c1, c2, f0, prefix, 512, 44100, seed=i)
    else:
        core.sins_stage(c0, c1, c2, f0, phase, 512, 44100, seed=i)
torch.cuda.synchronize()
print('source_hash', source_hash(model))
