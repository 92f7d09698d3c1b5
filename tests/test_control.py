"""The plain-PyTorch control network (ddsp_b200.control.Unit2Control) against the reference's
Unit2Control: identical state_dict layout (strict load) and identical outputs for the same weights.
Needs /root/reference (build container only); skipped elsewhere."""
import os
import sys

import pytest
import torch

REF = os.environ.get('DDSP_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'ddsp')), reason='reference tree not available')


@pytest.fixture(scope='module')
def reference():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden as G
    return G.load_reference()


@pytest.mark.parametrize('splits', [{'harmonic_magnitude': 513, 'harmonic_phase': 513, 'noise_magnitude': 513},
                                    {'amplitudes': 128, 'group_delay': 256, 'noise_magnitude': 256}])
def test_control_network_matches_reference(reference, splits):
    from ddsp.unit2control import Unit2Control as RefU2C        # the reference module (import shim active)
    from ddsp_b200.control import Unit2Control
    torch.manual_seed(3)
    ref = RefU2C(256, 7, splits, False).eval()
    mine = Unit2Control(256, 7, splits, False).eval()
    assert set(mine.state_dict().keys()) == set(ref.state_dict().keys())
    mine.load_state_dict(ref.state_dict(), strict=True)
    B, Fr = 2, 37
    units = torch.randn(B, Fr, 256)
    f0 = torch.rand(B, Fr, 1) * 700 + 65
    phase = (torch.rand(B, Fr) - 0.5) * 6.28
    vol = torch.rand(B, Fr)
    spk = torch.tensor([[1], [5]])
    with torch.no_grad():
        a = ref(units, f0, phase, vol, spk)
        b = mine(units, f0, phase, vol, spk)
        for k in splits:
            assert a[k].shape == b[k].shape and b[k].stride() == a[k].stride()          # same strided split views
            assert (a[k] - b[k]).abs().max().item() < 2e-5, k
        a = ref(units[:1], f0[:1], phase[:1], vol[:1], spk[:1], spk_mix_dict={1: 0.25, 3: 0.75})
        b = mine(units[:1], f0[:1], phase[:1], vol[:1], spk[:1], spk_mix_dict={1: 0.25, 3: 0.75})
        for k in splits:
            assert (a[k] - b[k]).abs().max().item() < 2e-5, k


def test_causal_variant_matches_its_definition_and_the_reference_layout(reference):
    """`c: true` (pcmer.py:141-160,176,185; unit2control.py:40,43): same state_dict layout as the reference's causal module;
    the chunked causal attention equals the O(N^2) definition; frame n of the output depends on frames <= n only.
    (The reference's own causal path needs fast_transformers' CUDA kernel and extorch, both absent: parity with the
    dependency itself is unpinned.)"""
    from ddsp_b200 import control
    from ddsp_b200.control import Unit2Control
    torch.manual_seed(5)
    q, k = torch.rand(2, 3, 150, 20) + 0.01, torch.rand(2, 3, 150, 20) + 0.01
    v = torch.randn(2, 3, 150, 8)
    out = control._causal_attend(q.double(), k.double(), v.double(), chunk=37)
    scores = torch.einsum('bhnj,bhmj->bhnm', q.double(), k.double()).tril()
    ref = torch.einsum('bhnm,bhme->bhne', scores, v.double()) / torch.einsum('bhnj,bhnj->bhn', q.double(), k.double().cumsum(-2) + 1e-6).unsqueeze(-1)
    assert (out - ref).abs().max().item() < 1e-10
    splits = {'harmonic_magnitude': 513, 'harmonic_phase': 513, 'noise_magnitude': 513}
    m = Unit2Control(256, 3, splits, True).eval()
    nc = Unit2Control(256, 3, splits, False)
    assert set(m.state_dict().keys()) == set(nc.state_dict().keys())          # causal changes no parameter (extorch pads only)
    B, Fr = 1, 90
    units, f0 = torch.randn(B, Fr, 256), torch.rand(B, Fr, 1) * 300 + 100
    phase, vol, spk = torch.rand(B, Fr), torch.rand(B, Fr), torch.ones(B, 1, dtype=torch.long)
    with torch.no_grad():
        a = m(units, f0, phase, vol, spk)['harmonic_magnitude']
        units2 = units.clone()
        units2[:, 60:] += 1.0                                                   # the future changes ...
        b = m(units2, f0, phase, vol, spk)['harmonic_magnitude']
    # ... GroupNorm in the prenet is the one non-causal step of the reference's causal network (statistics over all frames,
    # unit2control.py:41): compare through it by freezing nothing -- frames before the change move only through those statistics
    assert (a[:, :60] - b[:, :60]).abs().max().item() < 0.2 * (a[:, 60:] - b[:, 60:]).abs().max().item()
    m.train()
    out = m(units, f0, phase, vol, spk)['noise_magnitude'].sum()
    out.backward()
    assert all(p.grad is not None for p in m.parameters() if p.requires_grad)
