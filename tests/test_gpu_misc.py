"""Boundary behaviour on the GPU: errors, CUDA-graph capture, non-default streams / threads
(gui.py calls forward from a PortAudio callback thread), checkpoint keys."""
import os
import threading

import numpy as np
import pytest

from tests.gpu_util import HAS_CUDA, ctrl_views, dev, torch
from ddsp_b200.synthetic import make_inputs

pytestmark = pytest.mark.gpu

if HAS_CUDA:
    from ddsp_b200 import _cabi, core, vocoder


def _device_inputs(d):
    hm, hp, nm = ctrl_views(d['ctrl'], 'combsubfast')
    return hm, hp, nm, dev(d['f0_frames'])[..., None]


def _synth(d, seed=3, inputs=None):
    hm, hp, nm, f0 = inputs if inputs is not None else _device_inputs(d)
    pf, prefix, _ = core.phase_stage(f0, 512, 44100)
    return core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, seed=seed)


def test_shape_and_dtype_errors():
    d = make_inputs(2, 8, 1539, seed=1)
    hm, hp, nm = ctrl_views(d['ctrl'], 'combsubfast')
    f0 = dev(d['f0_frames'])[..., None]
    pf, prefix, _ = core.phase_stage(f0, 512, 44100)
    with pytest.raises(ValueError):
        core.combsubfast_stage(hm[..., :512], hp[..., :512], nm[..., :512], f0, prefix, 512, 44100)
    with pytest.raises(TypeError):
        core.phase_stage(f0.double(), 512, 44100)
    with pytest.raises(ValueError):
        core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, noise_u=torch.zeros(2, 5).cuda())
    with pytest.raises(_cabi.DDSPB200Error):                     # F*512 must stay below 2^24 (fp32-exact sample index)
        core.phase_stage(torch.zeros(1, 40000, 1).cuda(), 512, 44100)


def test_cuda_graph_capture_and_replay():
    d = make_inputs(1, 26, 1539, seed=2)
    inp = _device_inputs(d)            # host->device copies are not capturable: do them first
    ref = _synth(d, inputs=inp).clone()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        _synth(d, inputs=inp)          # warm-up: the one-time table initialisation must not be captured
        with torch.cuda.graph(g, stream=s):
            out = _synth(d, inputs=inp)
    torch.cuda.synchronize()
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)


def test_forward_from_worker_threads_on_their_own_streams():
    d = make_inputs(2, 40, 1539, seed=4)
    ref = _synth(d).cpu()
    results, errors = {}, []

    def work(i):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(5):
                    out = _synth(d)
                st.synchronize()
                results[i] = out.cpu()
        except Exception as e:        # pragma: no cover
            errors.append(e)
    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(4):
        assert torch.equal(results[i], ref)


def test_state_dict_keys_and_load_model(tmp_path):
    class Ctrl(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4, 1539)

        def forward(self, units, f0, phase, volume, spk_id, spk_mix_dict=None):
            e = self.lin(units)
            return dict(zip(['harmonic_magnitude', 'harmonic_phase', 'noise_magnitude'], torch.split(e, [513] * 3, dim=-1)))

    m = vocoder.CombSubFast(44100, 512, n_unit=4, n_spk=1, unit2ctrl=Ctrl())
    keys = set(m.state_dict().keys())
    assert {'sampling_rate', 'block_size', 'window', 'unit2ctrl.lin.weight', 'unit2ctrl.lin.bias'} == keys
    assert m.state_dict()['window'].shape == (1024,) and m.state_dict()['sampling_rate'].dtype == torch.int64
    # outside no_grad CombSubFast records its hand-written backward (tests/test_gpu_backward.py)
    m = m.cuda()
    units = torch.randn(1, 6, 4).cuda()
    f0 = torch.full((1, 6, 1), 220.0).cuda()
    vol = torch.zeros(1, 6).cuda()
    spk = torch.ones(1, 1, dtype=torch.long).cuda()
    sig, _, _ = m(units, f0, vol, spk)
    assert sig.requires_grad
    sig.square().mean().backward()
    assert m.unit2ctrl.lin.weight.grad is not None and torch.isfinite(m.unit2ctrl.lin.weight.grad).all()
    with torch.no_grad():
        sig, ph, (h, n) = m(units, f0, vol, spk)
    assert sig.shape == (1, 6 * 512) and h is sig and n is sig          # vocoder.py:492 returns the same tensor thrice
    with pytest.raises(ValueError):
        cfg = tmp_path / 'config.yaml'
        cfg.write_text('data:\n  sampling_rate: 44100\n  block_size: 512\n  encoder_out_channels: 4\nmodel:\n  type: Nope\n  n_spk: 1\n  c: false\n')
        vocoder.load_model(str(tmp_path / 'model_0.pt'))


def test_full_module_forward_with_builtin_control_network():
    """CombSubFast with the PyTorch control network of ddsp_b200.control: the synthesizer output must
    equal the stock-PyTorch op sequence fed with the same control tensors (captured by a hook)."""
    from oracle import torch_port as T
    torch.manual_seed(5)
    model = vocoder.CombSubFast(44100, 512, n_unit=32, n_spk=3).cuda().eval()
    assert type(model.unit2ctrl).__module__.endswith('control')
    B, F = 2, 40
    units = torch.randn(B, F, 32).cuda()
    f0 = (torch.rand(B, F, 1) * 300 + 100).cuda()
    vol = torch.rand(B, F).cuda()
    spk = torch.tensor([[1], [3]]).cuda()
    U = torch.rand(B, F * 512).cuda()
    cap = {}
    model.unit2ctrl.register_forward_hook(lambda m, i, o: cap.update(o))
    with torch.no_grad():
        sig, pf, _ = model(units, f0, vol, spk, noise_u=U)
        ref, pf_ref = T.combsubfast_forward(cap['harmonic_magnitude'], cap['harmonic_phase'], cap['noise_magnitude'],
                                            f0, model.window, noise_u=U)
    assert (sig - ref).abs().max().item() < 3e-5
    assert (pf[..., 0] - pf_ref).abs().max().item() < 1e-6


def test_first_call_under_graph_capture_is_refused_cleanly():
    """The lazy per-device table setup synchronises; under capture the library must say so instead of
    corrupting the capture (fresh process: the tables of this process are already initialised)."""
    import subprocess
    import sys
    code = (
        "import sys, torch; sys.path.insert(0, %r)\n"
        "from ddsp_b200 import core, _cabi\n"
        "f0 = torch.full((1, 8, 1), 220.0, device='cuda'); ctrl = torch.zeros(1, 8, 1539, device='cuda')\n"
        "hm, hp, nm = torch.split(ctrl, 513, dim=-1)\n"
        "_, prefix, _ = core.phase_stage(f0, 512, 44100)\n"
        "g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())\n"
        "try:\n"
        "    with torch.cuda.stream(s):\n"
        "        with torch.cuda.graph(g, stream=s):\n"
        "            core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100)\n"
        "    print('NO ERROR')\n"
        "except _cabi.DDSPB200Error as e:\n"
        "    print('REFUSED', e)\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert 'REFUSED' in out.stdout and 'warm-up' in out.stdout, (out.stdout, out.stderr[-500:])
