import ctypes, numpy as np, torch, time
torch.cuda.init()
rt = ctypes.CDLL('libcudart.so.12')
n = 339834880
def alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags))
    assert rc == 0, rc
    a = np.ctypeslib.as_array((ctypes.c_float * (n // 4)).from_address(p.value))
    return torch.from_numpy(a)
dst = torch.empty(n // 4, device='cuda')
def bw(h, name):
    h[:] = 1.0
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): dst.copy_(h, non_blocking=True)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): dst.copy_(h, non_blocking=True)
        e1.record(s); s.synchronize()
    print(name, 'pinned', h.is_pinned(), round(n * 10 / e0.elapsed_time(e1) / 1e6, 2), 'GB/s', flush=True)
bw(torch.empty(n // 4).pin_memory(), 'torch pin_memory')
bw(alloc(0), 'cudaHostAlloc default')
bw(alloc(4), 'cudaHostAlloc write-combined')
bw(alloc(1), 'cudaHostAlloc portable')
# D2H
h = torch.empty(112984064 // 4).pin_memory(); src = torch.empty(112984064 // 4, device='cuda')
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): h.copy_(src, non_blocking=True)
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(10): h.copy_(src, non_blocking=True)
    e1.record(s); s.synchronize()
print('D2H', round(112984064 * 10 / e0.elapsed_time(e1) / 1e6, 2), 'GB/s')
