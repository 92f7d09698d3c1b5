"""Multi-GPU plumbing on CPU: clips are sharded per rank with no data-path collective, and the
benchmark's reduction (max elapsed over ranks, sum of samples) is exercised with a world-size-2
gloo group.  The per-rank "synthesis" here is the CPU oracle -- this tests the host logic only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddsp_b200 import sharding


def test_shard_clips_partition():
    for n_clips, world in [(256, 8), (4, 8), (7, 2), (1, 1), (64, 3)]:
        seen = []
        for r in range(world):
            idx = sharding.shard_clips(n_clips, r, world)
            seen += list(idx)
            assert len(idx) in (n_clips // world, n_clips // world + 1) or len(idx) == 0 or n_clips < world
        assert sorted(seen) == list(range(n_clips))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import ddsp_oracle as O
    from ddsp_b200.synthetic import make_inputs
    d = make_inputs(4, 6, 1539, seed=5)
    mine = sharding.shard_clips(4, rank, world)
    sig, _ = O.combsubfast_forward(d['ctrl'][mine, :, :513], d['ctrl'][mine, :, 513:1026], d['ctrl'][mine, :, 1026:],
                                   d['f0_frames'][mine], d['U'][mine])
    elapsed_ms = 10.0 + 5.0 * rank
    total_ms, total_samples = sharding.reduce_timing(elapsed_ms, sig.size, device='cpu')
    gathered = [None] * world
    dist.all_gather_object(gathered, (list(mine), np.asarray(sig)))
    if rank == 0:
        q.put((total_ms, total_samples, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_process():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total_ms, total_samples, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total_ms == 15.0 and total_samples == 4 * 6 * 512       # max over ranks, sum of samples
    from oracle import ddsp_oracle as O
    from ddsp_b200.synthetic import make_inputs
    d = make_inputs(4, 6, 1539, seed=5)
    ref, _ = O.combsubfast_forward(d['ctrl'][..., :513], d['ctrl'][..., 513:1026], d['ctrl'][..., 1026:], d['f0_frames'], d['U'])
    out = np.zeros_like(ref)
    for idx, sig in gathered:
        out[idx] = sig
    assert np.array_equal(out, ref)          # sharding clips changes nothing: clips are independent


def test_host_placement_helpers(tmp_path):
    """sharding.near_gpu: cpulist parsing, sysfs lookup, affinity restored on exit, no-op without an answer."""
    import os
    assert sharding.parse_cpulist('0-3,8,10-11\n') == [0, 1, 2, 3, 8, 10, 11]
    assert sharding.parse_cpulist('') == []
    before = os.sched_getaffinity(0)
    dev = tmp_path / '0000:1b:00.0'
    dev.mkdir()
    first = min(before)
    (dev / 'numa_node').write_text('-1\n')
    (dev / 'local_cpulist').write_text('%d\n' % first)
    assert sharding.gpu_locality('0000:1B:00.0', str(tmp_path)) == (None, [first])
    with sharding.near_gpu(pci_bus_id='0000:1b:00.0', sysfs=str(tmp_path)) as info:
        assert os.sched_getaffinity(0) == {first}
        assert info['cpus'] == 1 and info['bound'] == (before != {first})
    assert os.sched_getaffinity(0) == before
    with sharding.near_gpu(pci_bus_id='0000:ff:00.0', sysfs=str(tmp_path)) as info:      # unknown device: nothing changes
        assert os.sched_getaffinity(0) == before and not info['bound']
    buf = np.ones(1 << 16, np.float32)
    assert sharding.node_of_address(buf.ctypes.data) in (None, 0, 1, 2, 3, 4, 5, 6, 7)
