"""Lists the CUDA kernels of one full forward (CombSubFast + PyTorch control network) at a streaming
block size -- the launch count is what bounds the graph-replay latency there (run on the GPU box)."""
import collections, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200 import vocoder
torch.manual_seed(0)
m = vocoder.CombSubFast(44100, 512, n_unit=256, n_spk=4).cuda().eval()
B, F = 1, 26
args = (torch.randn(B, F, 256, device='cuda'), torch.rand(B, F, 1, device='cuda') * 300 + 100, torch.rand(B, F, device='cuda'),
        torch.ones(B, 1, dtype=torch.long, device='cuda'))
with torch.no_grad():
    for _ in range(3): m(*args)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        m(*args)
        torch.cuda.synchronize()
cnt = collections.Counter()
tot = 0.0
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        cnt[e.name[:70]] += 1
        tot += e.device_time
print('kernels', sum(cnt.values()), 'device time us', round(tot, 1))
for k, v in cnt.most_common(40):
    print(f'{v:4d}  {k}')
