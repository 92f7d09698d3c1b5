"""Streaming CombSubFast (gui.py:373-388): blocks that carry the phase, three frames of
context and the noise hop index must reproduce ONE call over the concatenated frames."""
import numpy as np
import pytest

from tests.gpu_util import HAS_CUDA, ctrl_views, dev, torch
from ddsp_b200.synthetic import make_inputs

pytestmark = pytest.mark.gpu

if HAS_CUDA:
    from ddsp_b200 import core
    from ddsp_b200.streaming import CONTEXT, LATENCY, CombSubFastStream, StreamingCombSubFast


PAIRING_ULPS = 5e-7      # signals are O(0.1..1): a couple of fp32 ulps


def one_shot(d, U=None, seed=0, initial_phase=None):
    hm, hp, nm = ctrl_views(d['ctrl'], 'combsubfast')
    f0 = dev(d['f0_frames'])
    pf, prefix, _ = core.phase_stage(f0, 512, 44100, initial_phase, True)
    sig = core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, initial_phase, noise_u=U, seed=seed)
    return sig, pf


def stream_blocks(d, blocks, U=None, seed=0, initial_phase=None):
    hm, hp, nm = ctrl_views(d['ctrl'], 'combsubfast')
    f0 = dev(d['f0_frames'])
    s = CombSubFastStream(512, 44100, seed=seed, initial_phase=initial_phase)
    outs, phases, a = [], [], 0
    for k in blocks:
        b = a + k
        phases.append(s.begin(f0[:, a:b]))
        outs.append(s.finish(hm[:, a:b], hp[:, a:b], nm[:, a:b], noise_u=None if U is None else U[:, a * 512:b * 512]))
        assert outs[-1].shape[1] == 512 * (max(0, b - LATENCY) - max(0, a - LATENCY))
        assert s.frames_pushed == b and s.hops_emitted == max(0, b - LATENCY)
        a = b
    outs.append(s.flush())
    assert s.frames_pushed == 0
    return torch.cat(outs, dim=1), torch.cat(phases, dim=1)


@pytest.mark.parametrize('blocks', [[40], [1] * 12, [2, 1, 5, 9, 3, 20], [9] * 6, [26, 9, 9, 130, 1, 1]])
def test_stream_equals_one_call_injected_noise(blocks):
    F = sum(blocks)
    d = make_inputs(3, F, 1539, seed=7 + F)                       # voiced throughout: hop totals sum exactly in fp64
    U = dev(d['U'])
    ref, pf_ref = one_shot(d, U)
    out, pf = stream_blocks(d, blocks, U)
    assert out.shape == ref.shape
    assert torch.equal(pf, pf_ref)
    if len(blocks) == 1:
        assert torch.equal(out, ref), float((out - ref).abs().max())
    # the inverse FFT handles frames in pairs (V = Y_m + j*Y_m+1) and which frames share a pair depends on the
    # parity of the block's first frame: blocks agree with the one-shot call to the last ulp or two, not bitwise
    assert float((out - ref).abs().max()) <= PAIRING_ULPS


def test_stream_with_unvoiced_frames_and_initial_phase():
    # f0 ramps through 0 produce hop totals below 2^-24 granularity: the fp64 scan may associate differently
    # between the two paths, so the comparison allows rounding of the phase (<= 1e-9 rotations)
    d = make_inputs(4, 61, 1539, seed=3, zero_f0_fraction=0.25)
    U = dev(d['U'])
    ip = dev(np.array([0.5, -3.0, 2.0, 0.0], np.float32))
    ref, pf_ref = one_shot(d, U, initial_phase=ip)
    out, pf = stream_blocks(d, [5, 9, 1, 1, 30, 15], U, initial_phase=ip)
    assert float((pf - pf_ref).abs().max()) <= 1e-6
    assert float((out - ref).abs().max()) <= 2e-6


def test_stream_equals_one_call_in_kernel_noise():
    d = make_inputs(2, 48, 1539, seed=11)
    ref, _ = one_shot(d, None, seed=1234)
    out, _ = stream_blocks(d, [9, 9, 9, 9, 12], None, seed=1234)
    assert float((out - ref).abs().max()) <= PAIRING_ULPS
    same, _ = stream_blocks(d, [48], None, seed=1234)
    assert torch.equal(same, ref)
    other, _ = stream_blocks(d, [48], None, seed=99)
    assert not torch.equal(other, ref)


def test_push_operator_equals_begin_finish():
    """`push` runs the whole block in one operator of the extension host (csf_stream_push): bitwise the same stream as
    begin() + finish(), for split views of one tensor, separately stored rows and (B,k,1) f0."""
    d = make_inputs(3, 57, 1539, seed=21, zero_f0_fraction=0.1)
    hm, hp, nm = ctrl_views(d['ctrl'], 'combsubfast')
    f0 = dev(d['f0_frames'])
    ip = dev(np.array([0.3, -1.0, 2.0], np.float32))
    for separate in (False, True):
        sa = CombSubFastStream(512, 44100, seed=5, initial_phase=ip)
        sb = CombSubFastStream(512, 44100, seed=5, initial_phase=ip)
        a = 0
        for k in (9, 1, 26, 2, 19):
            b = a + k
            rows = [t[:, a:b] for t in (hm, hp, nm)]
            if separate:
                rows = [r.contiguous() for r in rows]
            out_a = sa.push(*rows, f0[:, a:b, None] if separate else f0[:, a:b])
            sb.begin(f0[:, a:b])
            out_b = sb.finish(*rows)
            assert torch.equal(out_a, out_b)
            assert (sa.frames_pushed, sa.hops_emitted, sa._t, sa._end) == (sb.frames_pushed, sb.hops_emitted, sb._t, sb._end)
            assert torch.equal(sa._carry, sb._carry)
            a = b
        assert torch.equal(sa.flush(), sb.flush())


def test_fused_stage_a_b_operator_equals_the_two_stages():
    d = make_inputs(2, 33, 1539, seed=8, zero_f0_fraction=0.2)
    hm, hp, nm = ctrl_views(d['ctrl'], 'combsubfast')
    f0 = dev(d['f0_frames'])
    ip = dev(np.array([1.5, -0.25], np.float32))
    for U in (None, dev(d['U'])):
        pf, prefix, _ = core.phase_stage(f0, 512, 44100, ip)
        ref = core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, noise_u=U, seed=77)
        sig, pf2, prefix2 = core.combsubfast_synth(hm, hp, nm, f0, 512, 44100, initial_phase=ip, noise_u=U, seed=77)
        assert torch.equal(sig, ref) and torch.equal(pf2, pf) and torch.equal(prefix2, prefix)


def test_carry_fold_keeps_the_phase_after_hours_of_audio():
    # a carry beyond 2^40 Hz*samples is folded modulo sr in the kernel; the phase only depends on carry mod sr
    f0 = torch.full((1, 8), 441.0, device='cuda')
    big = torch.tensor([44100.0 * 2 ** 26 + 11025.0], dtype=torch.float64, device='cuda')      # = 0.25 rotations
    small = torch.tensor([11025.0], dtype=torch.float64, device='cuda')
    pa, _ = core.phase_stage_stream(f0, 512, 44100, carry=big)
    pb, _ = core.phase_stage_stream(f0, 512, 44100, carry=small)
    assert float((pa - pb).abs().max()) <= 1e-6
    assert abs(float(pb[0, 0]) - 2 * np.pi * 0.26) < 1e-6        # inclusive cumsum: (11025 + 441) / 44100 rotations


def test_stream_argument_errors():
    s = CombSubFastStream()
    with pytest.raises(RuntimeError):
        s.finish(None, None, None)
    with pytest.raises(RuntimeError):
        s.flush()
    d = make_inputs(1, 4, 1539, seed=1)
    hm, hp, nm = ctrl_views(d['ctrl'], 'combsubfast')
    s.begin(dev(d['f0_frames']))
    with pytest.raises(ValueError):
        s.finish(hm[:, :3], hp[:, :3], nm[:, :3])
    with pytest.raises(ValueError):
        core.phase_stage_stream(dev(d['f0_frames']), 512, 44100, carry=torch.zeros(1, device='cuda'))   # fp32 carry


class FrameLocalCtrl(torch.nn.Module):
    """A control network without temporal context (rows depend on the frame's own inputs, phase included):
    with it the module-level stream must equal the one-shot module forward."""

    def __init__(self, n_unit):
        super().__init__()
        g = torch.Generator().manual_seed(5)
        self.w = torch.nn.Parameter(0.2 * torch.randn(n_unit, 1539, generator=g))

    def forward(self, units, f0, phase, volume, spk_id=None, spk_mix_dict=None):
        rows = units @ self.w + 0.3 * torch.sin(phase).reshape(phase.shape[0], -1, 1) + 0.1 * volume
        return dict(zip(('harmonic_magnitude', 'harmonic_phase', 'noise_magnitude'), torch.split(rows, 513, dim=-1)))


def test_module_level_stream_with_context_frames():
    from ddsp_b200.vocoder import CombSubFast
    B, F, n_unit = 2, 57, 16
    d = make_inputs(B, F, 1539, seed=21)
    model = CombSubFast(44100, 512, n_unit=n_unit, unit2ctrl=FrameLocalCtrl(n_unit)).cuda().eval()
    g = torch.Generator().manual_seed(9)
    units = torch.randn(B, F, n_unit, generator=g).cuda()
    volume = torch.rand(B, F, 1, generator=g).cuda()
    f0 = dev(d['f0_frames'])[..., None]
    U = dev(d['U'])
    with torch.no_grad():
        ref, ph_ref, _ = model(units, f0, volume, None, noise_u=U)
    s = StreamingCombSubFast(model, history=16)
    outs, phs, a = [], [], 0
    for k, c in [(9, 0), (9, 9), (1, 3), (20, 16), (18, 5)]:
        b = a + k
        audio, ph = s.push(units[:, a - c:b], f0[:, a - c:b], volume[:, a - c:b], None, n_context=c,
                           noise_u=U[:, a * 512:b * 512])
        outs.append(audio)
        phs.append(ph)
        a = b
    with pytest.raises(ValueError):
        s.push(units[:, :20], f0[:, :20], volume[:, :20], None, n_context=17)        # more than `history`
    outs.append(s.flush())
    out = torch.cat(outs, dim=1)
    assert torch.equal(torch.cat(phs, dim=1), ph_ref)
    # rows of repeated context frames are recomputed by the (frame-local) network: identical values
    assert float((out - ref).abs().max()) <= 1e-6
    assert CONTEXT == 3 and LATENCY == 2


def test_stream_block_under_cuda_graph():
    # one steady-state block (begin + finish) captured and replayed: what a low-latency caller would do
    d = make_inputs(1, 30, 1539, seed=2)
    hm, hp, nm = ctrl_views(d['ctrl'], 'combsubfast')
    f0 = dev(d['f0_frames'])
    s = CombSubFastStream(seed=5)
    s.push(hm[:, :12], hp[:, :12], nm[:, :12], f0[:, :12])
    eager = CombSubFastStream(seed=5)
    eager.push(hm[:, :12], hp[:, :12], nm[:, :12], f0[:, :12])
    want = eager.push(hm[:, 12:21], hp[:, 12:21], nm[:, 12:21], f0[:, 12:21])
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            got = s.push(hm[:, 12:21], hp[:, 12:21], nm[:, 12:21], f0[:, 12:21])
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(got, want)


# ---------------------------------------------------------------------------------------------------------------------
# frequency_filter models (Sins, CombSub old): carried phase + re-synthesised context frames
SPLITS = {'combsub': (256, 512, 256), 'sins': (128, 256, 256)}


def _filter_one_shot(model, d, U, seed):
    a, b, c = SPLITS[model]
    ctrl = dev(d['ctrl'])
    c0, c1, c2 = torch.split(ctrl, [a, b, c], dim=-1)
    f0 = dev(d['f0_frames'])
    _, prefix, phase = core.phase_stage(f0, 512, 44100, None, True, full_rate=(model == 'sins'))
    if model == 'sins':
        return core.sins_stage(c0, c1, c2, f0, phase, 512, 44100, noise_u=U, seed=seed)
    return core.combsub_stage(c0, c1, c2, f0, prefix, 512, 44100, noise_u=U, seed=seed)


@pytest.mark.parametrize('model', ['sins', 'combsub'])
@pytest.mark.parametrize('blocks', [[30], [1] * 10, [9] * 5, [2, 1, 7, 26, 3]])
@pytest.mark.parametrize('inject', [True, False])
def test_filter_model_stream_equals_one_call(model, blocks, inject):
    from ddsp_b200.streaming import FilterModelStream
    F = sum(blocks)
    a, b, c = SPLITS[model]
    d = make_inputs(2, F, a + b + c, seed=11 + F)
    U = dev(d['U']) if inject else None
    ref = _filter_one_shot(model, d, U, seed=5)
    ctrl = dev(d['ctrl'])
    c0, c1, c2 = torch.split(ctrl, [a, b, c], dim=-1)
    f0 = dev(d['f0_frames'])
    s = FilterModelStream(model, 512, 44100, seed=5)
    outs, pos = [], 0
    for k in blocks:
        e = pos + k
        outs.append(s.push(c0[:, pos:e], c1[:, pos:e], c2[:, pos:e], f0[:, pos:e],
                           noise_u=None if U is None else U[:, pos * 512:e * 512]))
        assert s.frames_pushed == e and s.hops_emitted == max(0, e - s.right)
        assert outs[-1][0].shape[1] == 512 * (max(0, e - s.right) - max(0, pos - s.right))
        pos = e
    outs.append(s.flush())
    for i, name in enumerate(('signal', 'harmonic', 'noise')):
        got = torch.cat([o[i] for o in outs], dim=1)
        assert got.shape == ref[i].shape
        err = float((got - ref[i]).abs().max())
        # the overlap-add order inside a call follows its run partition: a few fp32 ulps of O(1) signals
        assert err <= 3e-6, (name, err)


def test_apply_volume_mask_matches_the_callers_numpy_code():
    """main.py:112-116,159: threshold -> edge padding by 4 -> 9-frame maximum -> upsample -> multiply."""
    from oracle import ddsp_oracle as O
    rng = np.random.default_rng(3)
    B, F = 3, 57
    vol = (10.0 ** rng.uniform(-5, -1, size=(B, F))).astype(np.float32)
    vol[0, :6] = 1e-6
    vol[1, -3:] = 1e-6
    vol[2, 20:45] = 1e-7
    sig = rng.standard_normal((B, F * 512)).astype(np.float32)
    thr_db = -60.0
    expect = np.empty_like(sig)
    for b in range(B):
        mask = (vol[b] > 10 ** (float(thr_db) / 20)).astype('float')
        mask = np.pad(mask, (4, 4), constant_values=(mask[0], mask[-1]))
        mask = np.array([np.max(mask[n: n + 9]) for n in range(len(mask) - 8)])
        up = O.upsample(mask[None, :, None].astype(np.float32), 512)[0, :, 0]
        expect[b] = sig[b] * up
    got = core.apply_volume_mask_(dev(sig.copy()), dev(vol), thr_db)
    assert np.array_equal(got.cpu().numpy(), expect)
