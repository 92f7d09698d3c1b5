"""torch.profiler breakdown of the PyTorch control network (ddsp_b200.control.Unit2Control) at the
headline batch -- where the 26 ms of the full forward go (run on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200.control import Unit2Control
torch.manual_seed(0)
B, F = 64, 862
net = Unit2Control(256, 100, {'harmonic_magnitude': 513, 'harmonic_phase': 513, 'noise_magnitude': 513}).cuda().eval()
units = torch.randn(B, F, 256, device='cuda'); f0 = torch.rand(B, F, 1, device='cuda') * 300 + 100
ph = torch.rand(B, F, device='cuda'); vol = torch.rand(B, F, device='cuda'); spk = torch.ones(B, 1, dtype=torch.long, device='cuda')
with torch.no_grad():
    for _ in range(3): net(units, f0, ph, vol, spk)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for _ in range(3): net(units, f0, ph, vol, spk)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=70))
