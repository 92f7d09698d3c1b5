"""Smallest end-to-end invocation of every kernel family, for compute-sanitizer runs
(memcheck / racecheck / synccheck one tool per call):
    compute-sanitizer --tool memcheck python profiles/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200 import core                      # noqa: E402
from ddsp_b200.synthetic import make_inputs     # noqa: E402

B, F = 2, 9
for model, splits in (('combsubfast', (513, 513, 513)), ('combsub', (256, 512, 256)), ('sins', (128, 256, 256))):
    d = make_inputs(B, F, sum(splits), seed=5, zero_f0_fraction=0.2)
    c0, c1, c2 = torch.split(torch.from_numpy(d['ctrl']).cuda(), list(splits), dim=-1)
    f0 = torch.from_numpy(d['f0_frames']).cuda()[..., None]
    U = torch.from_numpy(d['U']).cuda()
    pf, prefix, phase = core.phase_stage(f0, 512, 44100, full_rate=(model == 'sins'))
    for u in (U, None):
        if model == 'combsubfast':
            out = core.combsubfast_stage(c0, c1, c2, f0, prefix, 512, 44100, noise_u=u, seed=7)
        elif model == 'combsub':
            out = core.combsub_stage(c0, c1, c2, f0, prefix, 512, 44100, noise_u=u, seed=7)[0]
        else:
            out = core.sins_stage(c0, c1, c2, f0, phase, 512, 44100, noise_u=u, seed=7)[0]
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
# longer clip so that runs have seams (atomics) and several frames per warp
d = make_inputs(3, 300, 1539, seed=6)
c0, c1, c2 = torch.split(torch.from_numpy(d['ctrl']).cuda(), [513] * 3, dim=-1)
f0 = torch.from_numpy(d['f0_frames']).cuda()[..., None]
pf, prefix, _ = core.phase_stage(f0, 512, 44100)
out = core.combsubfast_stage(c0, c1, c2, f0, prefix, 512, 44100, seed=1)
rot = core.fo_to_rot(torch.rand(2, 5000).cuda() * 500, 44100, None, True)
up = core.upsample(torch.rand(2, 7, 3).cuda(), 512)
torch.cuda.synchronize()
print('sanitize_small ok')
