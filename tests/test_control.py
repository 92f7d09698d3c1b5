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


def test_causal_variant_is_refused():
    from ddsp_b200.control import Unit2Control
    with pytest.raises(NotImplementedError):
        Unit2Control(256, 1, {'a': 4}, True)
