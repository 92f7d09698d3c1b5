"""GraphedForward: the whole module forward (control network + synthesizer) replayed as one CUDA graph per
input shape -- the low-latency form of the GUI callback (gui.py:125-127, BASELINE config 5)."""
import pytest

from tests.gpu_util import HAS_CUDA, torch

pytestmark = pytest.mark.gpu

if HAS_CUDA:
    from ddsp_b200 import vocoder


class Silenced(torch.nn.Module if torch is not None else object):
    """Control network wrapper that turns the noise branch off (exp(-80)/128 flushes to 0), so that outputs
    can be compared exactly although every call draws different noise."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, *a, **k):
        out = dict(self.inner(*a, **k))
        out['noise_magnitude'] = out['noise_magnitude'] * 0 - 80.0
        return out


def _inputs(B, F, n_unit, seed):
    g = torch.Generator().manual_seed(seed)
    units = torch.randn(B, F, n_unit, generator=g).cuda()
    f0 = (torch.rand(B, F, 1, generator=g) * 300 + 100).cuda()
    vol = torch.rand(B, F, generator=g).cuda()
    spk = torch.ones(B, 1, dtype=torch.int64).cuda()
    return units, f0, vol, spk


def test_graphed_combsubfast_equals_eager_and_caches_per_shape():
    torch.manual_seed(3)
    model = vocoder.CombSubFast(44100, 512, n_unit=32, n_spk=2).cuda().eval()
    model.unit2ctrl = Silenced(model.unit2ctrl)
    fast = vocoder.GraphedForward(model)
    for F, seed in [(9, 1), (26, 2), (9, 3), (26, 4)]:
        units, f0, vol, spk = _inputs(1, F, 32, seed)
        with torch.no_grad():
            ref, ph_ref, _ = model(units, f0, vol, spk)
        sig, ph, (h, n) = fast(units, f0, vol, spk)
        assert h is sig and n is sig
        assert torch.equal(ph, ph_ref)
        # cuBLAS may choose another algorithm for the control network's GEMMs under capture (workspace policy):
        # the control rows agree to ~1e-6, not bitwise
        assert float((sig - ref).abs().max()) < 5e-6
    assert len(fast._graphs) == 2
    # outputs are clones: a later call must not change an earlier result
    units, f0, vol, spk = _inputs(1, 9, 32, 7)
    a, _, _ = fast(units, f0, vol, spk)
    keep = a.clone()
    fast(*_inputs(1, 9, 32, 8))
    assert torch.equal(a, keep)
    raw = vocoder.GraphedForward(model, copy_outputs=False)
    x, _, _ = raw(units, f0, vol, spk)
    y, _, _ = raw(*_inputs(1, 9, 32, 8))
    assert x.data_ptr() == y.data_ptr()


def test_graphed_combsubfast_draws_fresh_noise_on_every_replay():
    torch.manual_seed(4)
    model = vocoder.CombSubFast(44100, 512, n_unit=16).cuda().eval()
    fast = vocoder.GraphedForward(model)
    units, f0, vol, spk = _inputs(2, 18, 16, 5)
    a, _, _ = fast(units, f0, vol, spk)
    b, _, _ = fast(units, f0, vol, spk)
    c, _, _ = fast(units, f0, vol, spk)
    d_ab, d_bc = float((a - b).abs().max()), float((b - c).abs().max())
    assert 1e-4 < d_ab < 0.5 and 1e-4 < d_bc < 0.5           # only the (small) noise branch differs
    assert model._seed_device is None                          # eager calls on the same model stay untouched


@pytest.mark.parametrize('kind', ['sins', 'combsub'])
def test_graphed_filter_models(kind):
    torch.manual_seed(6)
    if kind == 'sins':
        model = vocoder.Sins(44100, 512, 128, 256, 256, n_unit=16).cuda().eval()
    else:
        model = vocoder.CombSub(44100, 512, 256, 512, 256, n_unit=16).cuda().eval()
    fast = vocoder.GraphedForward(model)
    units, f0, vol, spk = _inputs(1, 26, 16, 9)
    with torch.no_grad():
        _, ph_ref, (h_ref, _) = model(units, f0, vol, spk)
    s1, ph, (h1, n1) = fast(units, f0, vol, spk)
    s2, _, (h2, n2) = fast(units, f0, vol, spk)
    assert torch.equal(ph, ph_ref)
    assert float((h1 - h_ref).abs().max()) < 2e-5              # control rows to ~1e-6 (cuBLAS under capture)
    assert torch.equal(h1, h2)                                 # the harmonic branch does not see the noise
    assert not torch.equal(n1, n2)                             # refilled inside the graph
    assert float((s1 - (h1 + n1)).abs().max()) < 1e-6
