"""Times the CombSubFast forward and gradient kernels at the headline shape (run on the GPU box)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200 import core
from ddsp_b200.synthetic import make_inputs
B, F = 64, 862
d = make_inputs(B, F, 1539, seed=1, noise=False)
ctrl = torch.from_numpy(d['ctrl']).cuda(); hm, hp, nm = torch.split(ctrl, 513, dim=-1)
f0 = torch.from_numpy(d['f0_frames']).cuda()[..., None]
_, prefix, _ = core.phase_stage(f0, 512, 44100, None, True)
R = torch.randn(B, F * 512, device='cuda')
for fn, name in [(lambda: core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, None, seed=3), 'forward'),
                 (lambda: core.combsubfast_backward_stage(R, hm, hp, nm, f0, prefix, 512, 44100, seed=3), 'backward')]:
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1) / 50, 4), 'ms')
