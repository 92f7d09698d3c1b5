"""Clip sharding across GPUs: clips (batch rows) are independent on this path,
so rank r simply takes clips r, r+world, ... and no collective touches the data.  The only
communication is the benchmark's timing reduction."""
import torch
import torch.distributed as dist


def shard_clips(n_clips, rank, world):
    """Indices of the clips rank `rank` synthesises (round-robin, as SURVEY §8e: clip i -> GPU i mod G)."""
    return range(rank, n_clips, world)


def reduce_timing(elapsed_ms, samples, device):
    """(max elapsed over ranks, total samples over ranks); identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed_ms), int(samples)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    n = torch.tensor([int(samples)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())


# ---- host placement of a rank's pinned buffers ---------------------------------------------------------------------
# Every rank of an 8-GPU box moves its own clips over its own PCIe link, but the pages of its pinned host buffers land
# on whatever NUMA node the allocating thread happened to run on; buffers on the far socket make every H2D / D2H copy
# cross the inter-socket link.  `near_gpu(device)` is a context manager that confines the calling thread to the CPUs
# next to the GPU (sysfs `local_cpulist` of its PCI function) and prefers that node for new pages while the pinned
# buffers are allocated; it restores the previous affinity / policy on exit and is a no-op where sysfs has no answer.
import contextlib
import ctypes
import os

_MPOL_DEFAULT, _MPOL_PREFERRED = 0, 1
_SYS_SET_MEMPOLICY = 238          # x86_64


def parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the kernel's cpulist format)."""
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_locality(pci_bus_id, sysfs='/sys/bus/pci/devices'):
    """(numa_node or None, local cpus or []) of the PCI function `pci_bus_id` ('0000:1b:00.0')."""
    node, cpus = None, []
    base = os.path.join(sysfs, pci_bus_id.lower())
    try:
        with open(os.path.join(base, 'numa_node')) as f:
            n = int(f.read().strip())
        node = n if n >= 0 else None
    except (OSError, ValueError):
        pass
    try:
        with open(os.path.join(base, 'local_cpulist')) as f:
            cpus = parse_cpulist(f.read())
    except (OSError, ValueError):
        pass
    return node, cpus


def _set_mempolicy(mode, node):
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        if node is None:
            return libc.syscall(_SYS_SET_MEMPOLICY, _MPOL_DEFAULT, None, 0) == 0
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        return libc.syscall(_SYS_SET_MEMPOLICY, mode, mask, 16 * 64 + 1) == 0
    except Exception:      # pragma: no cover
        return False


def node_of_address(ptr):
    """NUMA node that holds the page at host address `ptr` (move_pages(2) as a query), or None."""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        page = ctypes.c_void_p(ptr & ~4095)
        status = ctypes.c_int(-1)
        rc = libc.syscall(279, 0, ctypes.c_ulong(1), ctypes.byref(page), None, ctypes.byref(status), 0)
        return int(status.value) if rc == 0 and status.value >= 0 else None
    except Exception:      # pragma: no cover
        return None


@contextlib.contextmanager
def near_gpu(device=None, pci_bus_id=None, sysfs='/sys/bus/pci/devices'):
    """Allocate (and first-touch) host buffers inside this block so that their pages sit next to the GPU.
    Yields {'node', 'cpus', 'bound'} for the record."""
    if pci_bus_id is None:
        p = torch.cuda.get_device_properties(device)
        pci_bus_id = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    node, cpus = gpu_locality(pci_bus_id, sysfs)
    if os.environ.get('DDSP_B200_NO_NUMA') == '1':          # experiments: leave the placement to the scheduler
        node, cpus = None, []
    before = os.sched_getaffinity(0)
    target = before & set(cpus)
    info = {'node': node, 'cpus': len(target), 'bound': False}
    policy = False
    try:
        if target and target != before:
            os.sched_setaffinity(0, target)
            info['bound'] = True
        if node is not None:
            policy = _set_mempolicy(_MPOL_PREFERRED, node)
            info['bound'] = info['bound'] or policy
        yield info
    finally:
        if policy:
            _set_mempolicy(_MPOL_DEFAULT, None)
        if info['bound'] and target and target != before:
            os.sched_setaffinity(0, before)
