import os, sys, time, torch, cProfile, pstats, io
sys.path.insert(0, os.getcwd())
from ddsp_b200 import core
from ddsp_b200.synthetic import make_inputs
import numpy as np
F=26
d = make_inputs(1, F, 1539, seed=1, noise=False)
ctrl = torch.from_numpy(d['ctrl']).cuda(); hm, hp, nm = torch.split(ctrl, 513, dim=-1)
f0 = torch.from_numpy(d['f0_frames']).cuda()[..., None]
win = torch.sqrt(torch.hann_window(1024)).cuda()
def call(i):
    pf, prefix, _ = core.phase_stage(f0, 512, 44100)
    return core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, seed=i, window=win)
for i in range(200): call(i)
torch.cuda.synchronize()
t=time.perf_counter()
for i in range(2000): call(i)
torch.cuda.synchronize()
print('wall per call us', (time.perf_counter()-t)/2000*1e6)
ops = torch.ops.ddsp_b200
t=time.perf_counter()
for i in range(2000):
    pf, prefix, _ = ops.phase(f0, 512, 44100.0, None, True, False, None)
torch.cuda.synchronize(); print('ops.phase us', (time.perf_counter()-t)/2000*1e6)
t=time.perf_counter()
for i in range(2000):
    ops.combsubfast(hm, hp, nm, f0, prefix, 512, 44100.0, None, i, win, None, 0, None)
torch.cuda.synchronize(); print('ops.combsubfast us', (time.perf_counter()-t)/2000*1e6)
t=time.perf_counter()
for i in range(2000):
    torch.empty((1, F*512), device='cuda')
print('torch.empty us', (time.perf_counter()-t)/2000*1e6)
pr=cProfile.Profile(); pr.enable()
for i in range(2000): call(i)
pr.disable(); torch.cuda.synchronize()
s=io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(12); print(s.getvalue()[:2500])
