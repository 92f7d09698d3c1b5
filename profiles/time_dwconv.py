import os, sys, torch
sys.path.insert(0, os.getcwd())
from ddsp_b200 import core
torch.manual_seed(0)
g = torch.randn(64, 862, 512, device='cuda'); w = torch.randn(512, 1, 31, device='cuda') * 0.1; b = torch.randn(512, device='cuda')
def timeit(fn, name, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1) / n * 1e3, 1), 'us')
timeit(lambda: core.dwconv_silu(g, w, b), 'dwconv_silu 64x862x512')
ref = torch.nn.functional.silu(torch.nn.functional.conv1d(g.transpose(1, 2).double(), w.double(), b.double(), padding=15, groups=512)).transpose(1, 2)
print('max err', (core.dwconv_silu(g, w, b).double() - ref).abs().max().item())
g2 = torch.randn(3, 77, 384, device='cuda'); w2 = torch.randn(384, 1, 31, device='cuda') * 0.1; b2 = torch.randn(384, device='cuda')
ref2 = torch.nn.functional.silu(torch.nn.functional.conv1d(g2.transpose(1, 2).double(), w2.double(), b2.double(), padding=15, groups=384)).transpose(1, 2)
print('max err generic', (core.dwconv_silu(g2, w2, b2).double() - ref2).abs().max().item())
