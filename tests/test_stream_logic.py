"""Host logic of ddsp_b200.streaming on CPU: the window / tail / emit bookkeeping is run against a CPU
stand-in of the two C-ABI stages built from the oracle (test infrastructure only -- the product path has
no CPU route; here `core`'s stage functions are monkeypatched), and the concatenated blocks must equal the
oracle's one-shot synthesis of all frames."""
import numpy as np
import pytest
import torch

from oracle import ddsp_oracle as O
from ddsp_b200 import core, streaming
from ddsp_b200.synthetic import make_inputs

SR, HOP = 44100, 512


def fake_phase_stage_stream(f0_frames, block_size, sampling_rate, carry=None, initial_phase=None):
    f0 = f0_frames.numpy().astype(np.float32)
    B, F = f0.shape
    up = O.upsample(f0[..., None], HOP)[..., 0].astype(np.float64)                 # fp32 values, summed in fp64
    totals = up.reshape(B, F, HOP).sum(-1)
    start = np.zeros(B) if carry is None else carry.numpy().astype(np.float64)
    if carry is None and initial_phase is not None:
        start = np.asarray(initial_phase, np.float64) / 2 / np.pi * SR
    prefix = start[:, None] + np.concatenate([np.zeros((B, 1)), np.cumsum(totals, 1)[:, :-1]], 1)
    c = (prefix + f0) / SR
    phase = (np.float32(2 * np.pi) * (c - np.rint(c)).astype(np.float32)).astype(np.float32)
    return torch.from_numpy(phase), torch.from_numpy(prefix)


def fake_combsubfast_stage(hm, hp, nm, f0_frames, prefix, block_size, sampling_rate, initial_phase=None, noise_u=None,
                           seed=0, window=None, out=None, hop_offset=None):
    start = prefix[:, 0].numpy()
    ip = 2 * np.pi * (start / SR - np.rint(start / SR))                               # the carried phase, in radians
    sig, _ = O.combsubfast_forward(hm.numpy(), hp.numpy(), nm.numpy(), f0_frames.numpy(), noise_u.numpy(),
                                   initial_phase=ip, exact_cumsum=True)
    return torch.from_numpy(np.ascontiguousarray(sig.astype(np.float32)))


@pytest.fixture
def cpu_backend(monkeypatch):
    monkeypatch.setattr(core, '_need_cuda_f32', lambda t, name: t)
    monkeypatch.setattr(core, 'phase_stage_stream', fake_phase_stage_stream)
    monkeypatch.setattr(core, 'combsubfast_stage', fake_combsubfast_stage)


@pytest.mark.parametrize('blocks', [[1, 1, 1, 1, 1, 1, 1], [2, 5, 1, 9, 70, 3], [26, 9, 9]])
def test_blocks_equal_one_shot_on_the_oracle(cpu_backend, blocks):
    F = sum(blocks)
    d = make_inputs(2, F, 1539, seed=31 + F, zero_f0_fraction=0.1)
    ctrl = torch.from_numpy(d['ctrl'])
    hm, hp, nm = torch.split(ctrl, 513, dim=-1)
    f0 = torch.from_numpy(d['f0_frames'])
    U = torch.from_numpy(d['U'])
    ref, pf_ref = O.combsubfast_forward(d['ctrl'][..., :513], d['ctrl'][..., 513:1026], d['ctrl'][..., 1026:],
                                        d['f0_frames'], d['U'], exact_cumsum=True)
    s = streaming.CombSubFastStream(HOP, SR)
    outs, phases, a = [], [], 0
    for k in blocks:
        b = a + k
        phases.append(s.begin(f0[:, a:b]))
        outs.append(s.finish(hm[:, a:b], hp[:, a:b], nm[:, a:b], noise_u=U[:, a * HOP:b * HOP]))
        assert outs[-1].shape == (2, HOP * (max(0, b - streaming.LATENCY) - max(0, a - streaming.LATENCY)))
        assert s.frames_pushed == b and s.hops_emitted == max(0, b - streaming.LATENCY)
        a = b
    outs.append(s.flush())
    out = torch.cat(outs, dim=1).numpy()
    assert out.shape == ref.shape
    assert np.abs(torch.cat(phases, dim=1).numpy() - pf_ref).max() < 2e-6
    assert np.abs(out - ref).max() < 2e-6, np.abs(out - ref).max()
    assert s.frames_pushed == 0 and s._f0_buf is None


def test_buffer_wraps_and_rejects_misuse(cpu_backend):
    d = make_inputs(1, 300, 1539, seed=3)
    hm, hp, nm = torch.split(torch.from_numpy(d['ctrl']), 513, dim=-1)
    f0 = torch.from_numpy(d['f0_frames'])
    U = torch.from_numpy(d['U'])
    ref, _ = O.combsubfast_forward(d['ctrl'][..., :513], d['ctrl'][..., 513:1026], d['ctrl'][..., 1026:],
                                   d['f0_frames'], d['U'], exact_cumsum=True)
    s = streaming.CombSubFastStream(HOP, SR)
    outs = []
    for a in range(0, 300, 10):                      # capacity is 64 frames: the tail moves to the front several times
        outs.append(s.push(hm[:, a:a + 10], hp[:, a:a + 10], nm[:, a:a + 10], f0[:, a:a + 10], noise_u=U[:, a * HOP:(a + 10) * HOP]))
        assert s._f0_buf.shape[1] == 64
    with pytest.raises(ValueError):
        s.begin(torch.zeros(2, 4))                   # clip count changed mid-stream
    with pytest.raises(ValueError):
        s.begin(f0[:, :0])
    outs.append(s.flush())
    assert np.abs(torch.cat(outs, dim=1).numpy() - ref).max() < 2e-6
    with pytest.raises(RuntimeError):
        s.finish(hm[:, :1], hp[:, :1], nm[:, :1])
