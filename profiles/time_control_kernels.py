"""Times the control-network kernels at the headline batch (run on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200 import core
torch.manual_seed(0)
B, N, H = 64, 862, 8
x = torch.randn(B, N, H * 64, device='cuda')
proj = torch.randn(266, 64, device='cuda')
u = torch.randn(B, N, 1024, device='cuda'); w = torch.randn(512, 1, 31, device='cuda'); bias = torch.randn(512, device='cuda')
def timeit(fn, name, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1) / n, 4), 'ms')
timeit(lambda: core.performer_project_features(x, proj, H, True), 'project_features(query)')
timeit(lambda: core.performer_project_features(x, proj, H, False), 'project_features(key)')
timeit(lambda: core.performer_features(torch.matmul((64 ** -0.25 * x).view(-1, 64), proj.t()), x, H, True), 'mm + features(query)')
timeit(lambda: core.glu_dwconv_silu(u, w, bias), 'glu_dwconv_silu')
