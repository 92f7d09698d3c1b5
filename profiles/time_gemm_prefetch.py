"""L2-prefetch-distance experiment for the batched GEMMs of the control network (run on the GPU box):
DDSP_B200_GEMM_PREFETCH=D makes the TMA producer of every gemm3x launch prefetch the DRAM-streamed operand tiles D
k-blocks ahead into L2 (0 = off).  Prints the per-kernel averages of three control-network forwards at 64 x 10 s."""
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): net(units, f0, ph, vol, spk)
    e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / 10
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(3): net(units, f0, ph, vol, spk)
        torch.cuda.synchronize()
tag = os.environ.get('DDSP_B200_GEMM_PREFETCH', 'default')
print(f'prefetch={tag} control forward {total:.3f} ms')
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:11]:
    print(f'  prefetch={tag} {ev.key[:60]:60s} n={ev.count:3d} avg {ev.device_time_total / ev.count:8.1f} us')
