"""A few forwards of the tcgen05 control network at the headline batch, for ncu captures (run on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200.control import Unit2Control
torch.manual_seed(0)
B, F = 64, 862
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
net = Unit2Control(256, 1, {'harmonic_magnitude': 513, 'harmonic_phase': 513, 'noise_magnitude': 513}).cuda().eval()
units = torch.randn(B, F, 256, device='cuda'); f0 = torch.rand(B, F, 1, device='cuda') * 300 + 100
ph = torch.rand(B, F, device='cuda'); vol = torch.rand(B, F, device='cuda'); spk = torch.ones(B, 1, dtype=torch.long, device='cuda')
with torch.no_grad():
    for _ in range(n):
        net(units, f0, ph, vol, spk)
    torch.cuda.synchronize()
print('done')
