"""Probe: does the H2D/D2H bandwidth of pinned host buffers depend on the CPU set the allocating thread runs on?
Prints the GPU's NUMA node, this process's affinity and copy bandwidths per candidate CPU set."""
import glob
import os
import time

import torch


def cpulist(s):
    out = set()
    for part in s.strip().split(','):
        if not part:
            continue
        a, _, b = part.partition('-')
        out.update(range(int(a), int(b or a) + 1))
    return out


def bw(nbytes=256 << 20, reps=6):
    h = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    h.fill_(1.0)
    d = torch.empty(nbytes // 4, dtype=torch.float32, device='cuda')
    res = {}
    for name, (src, dst) in {'h2d': (h, d), 'd2h': (d, h)}.items():
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        res[name] = round(nbytes * reps / (time.perf_counter() - t0) / 1e9, 1)
    return res


def main():
    torch.cuda.init()
    prop = torch.cuda.get_device_properties(0)
    bus = f'{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0'
    base = f'/sys/bus/pci/devices/{bus}'
    node = open(base + '/numa_node').read().strip() if os.path.exists(base + '/numa_node') else '?'
    local = open(base + '/local_cpulist').read().strip() if os.path.exists(base + '/local_cpulist') else ''
    aff = os.sched_getaffinity(0)
    print('gpu', bus, 'numa_node', node, 'local_cpulist', local)
    print('affinity', sorted(aff))
    for n in sorted(glob.glob('/sys/devices/system/node/node*/cpulist')):
        cp = cpulist(open(n).read())
        print(n.split('/')[-2], 'cpus in affinity:', sorted(cp & aff))
    print('all  ', bw())
    for n in sorted(glob.glob('/sys/devices/system/node/node*/cpulist')):
        cp = cpulist(open(n).read()) & aff
        if cp:
            os.sched_setaffinity(0, cp)
            print(n.split('/')[-2], bw())
            os.sched_setaffinity(0, aff)
    if local:
        cp = cpulist(local) & aff
        print('gpu-local cpus in affinity:', sorted(cp))


if __name__ == '__main__':
    main()
