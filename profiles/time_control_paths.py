"""Full-forward device time for mid-sized calls with and without the fused attention kernels
(ddsp_b200.control._ATTENTION_KERNEL_MAX_FRAMES) -- picks the dispatch threshold (run on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200 import vocoder, control
torch.manual_seed(0)
m = vocoder.CombSubFast(44100, 512, n_unit=256, n_spk=4).cuda().eval()
for B, F in [(1, 130), (1, 431), (1, 862), (4, 862), (8, 862), (1, 2000)]:
    args = (torch.randn(B, F, 256, device='cuda'), torch.rand(B, F, 1, device='cuda') * 300 + 100,
            torch.rand(B, F, device='cuda'), torch.ones(B, 1, dtype=torch.long, device='cuda'))
    res = []
    for thr in (4096, 0):
        control._ATTENTION_KERNEL_MAX_FRAMES = thr
        with torch.no_grad():
            for _ in range(3): m(*args)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                m(*args)
                with torch.cuda.graph(g, stream=s):
                    m(*args)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20): g.replay()
            e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 20)
    print(f'B={B} F={F}: fused attention {res[0]:.3f} ms, multi-kernel path {res[1]:.3f} ms')
