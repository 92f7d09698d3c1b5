"""Where does the e2e_forward leg lose time against the device-timed forward?  (scratch probe, GPU box)"""
import os, sys, time, torch, contextlib, io
sys.path.insert(0, '/root/repo')
from ddsp_b200 import vocoder
from ddsp_b200.control import Unit2Control
dev = torch.device('cuda:0')
B, F, SR, HOP = 64, 862, 44100, 512
torch.manual_seed(1234)
splits = {'harmonic_magnitude': 513, 'harmonic_phase': 513, 'noise_magnitude': 513}
with contextlib.redirect_stdout(io.StringIO()):
    net = vocoder.CombSubFast(SR, HOP, n_unit=256, n_spk=1, unit2ctrl=Unit2Control(256, 1, splits)).to(dev).eval()
h_units = torch.randn(B, F, 256).pin_memory(); h_vol = torch.rand(B, F).pin_memory(); h_f0 = (torch.rand(B, F) * 300 + 100).pin_memory()
h_out = [torch.empty(B, F * HOP).pin_memory() for _ in range(2)]
spk = torch.ones(B, 1, dtype=torch.long, device=dev)
d_units = [torch.empty((B, F, 256), device=dev) for _ in range(2)]
d_vol = [torch.empty((B, F), device=dev) for _ in range(2)]
d_f0 = [torch.empty((B, F), device=dev) for _ in range(2)]
s_in, s_cp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def run(n, copies_in=True, copies_out=True):
    ev_cp = [None] * n
    for i in range(n):
        k = i & 1
        with torch.cuda.stream(s_in):
            if i >= 2: s_in.wait_event(ev_cp[i - 2])
            if copies_in:
                d_units[k].copy_(h_units, non_blocking=True); d_f0[k].copy_(h_f0, non_blocking=True); d_vol[k].copy_(h_vol, non_blocking=True)
            ev_in = torch.cuda.Event(); ev_in.record()
        with torch.cuda.stream(s_cp):
            s_cp.wait_event(ev_in)
            with torch.no_grad():
                sig = net(d_units[k], d_f0[k][..., None], d_vol[k], spk)[0]
            ev_cp[i] = torch.cuda.Event(); ev_cp[i].record()
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cp[i])
            if copies_out: h_out[k].copy_(sig, non_blocking=True)
            sig.record_stream(s_out)
def timed(n, **kw):
    run(6, **kw); torch.cuda.synchronize()
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(s_in): x0.record()
    run(n, **kw)
    t_enq = time.perf_counter() - t0
    with torch.cuda.stream(s_out):
        s_out.wait_stream(s_in); s_out.wait_stream(s_cp); x1.record()
    torch.cuda.synchronize()
    return x0.elapsed_time(x1) / n, t_enq / n * 1e3
for n in (20, 40):
    for kw in ({}, {'copies_in': False}, {'copies_out': False}, {'copies_in': False, 'copies_out': False}):
        ms, enq = timed(n, **kw)
        print(f'steps {n} {kw}: {ms:.3f} ms/step (host enqueue {enq:.3f} ms/step)')
