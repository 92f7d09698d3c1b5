"""Import shim for the unmodified reference modules (oracle/_ref/ or /root/reference).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference imports seven third-party packages that are not
installed in this image; none of them is touched by the synthesizer forward
path, so they are replaced by empty stub modules before `ddsp.vocoder` is imported:
pyworld, parselmouth, torchcrepe, resampy, sklearn(.cluster), extorch (Conv1dEx / Transpose as plain
torch modules), fast_transformers(.causal_product) and encoder.hubert.model (HubertSoft).
"""
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def install_stubs():
    import torch

    def stub(name):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m
        return m
    for n in ['pyworld', 'parselmouth', 'torchcrepe', 'resampy', 'sklearn', 'sklearn.cluster']:
        if n not in sys.modules:
            stub(n)
    if not hasattr(sys.modules['sklearn.cluster'], 'KMeans'):
        sys.modules['sklearn.cluster'].KMeans = object
    ext = stub('extorch')

    class Conv1dEx(torch.nn.Conv1d):                # unit2control.py:40,43 / pcmer.py:54 use padding="same"
        def __init__(self, *a, causal=False, **k):
            assert not causal
            super().__init__(*a, **k)

    class Transpose(torch.nn.Module):               # unit2control.py:39,44 / pcmer.py:51,57
        def __init__(self, a, b):
            super().__init__()
            self.a, self.b = a, b

        def forward(self, x):
            return x.transpose(self.a, self.b)
    ext.Conv1dEx, ext.Transpose = Conv1dEx, Transpose
    stub('fast_transformers')
    stub('fast_transformers.causal_product').CausalDotProduct = None    # only dereferenced when c=True


def available():
    return os.path.exists(os.path.join(REF_DIR, 'ddsp', 'vocoder.py'))


def load_reference(ref_root=None):
    """Returns the reference's (ddsp.core, ddsp.vocoder) modules.  ref_root defaults to oracle/_ref; with
    the full tree (/root/reference) the real `encoder` package is importable, with oracle/_ref it is stubbed."""
    install_stubs()
    root = ref_root or REF_DIR
    if not os.path.exists(os.path.join(root, 'encoder')):
        for n in ['encoder', 'encoder.hubert', 'encoder.hubert.model']:
            m = types.ModuleType(n)
            m.__spec__ = importlib.machinery.ModuleSpec(n, None)
            m.__path__ = []
            sys.modules[n] = m
        sys.modules['encoder.hubert.model'].HubertSoft = object
    for k in [k for k in sys.modules if k == 'ddsp' or k.startswith('ddsp.')]:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        import ddsp.core as core
        import ddsp.vocoder as vocoder
    finally:
        sys.path.remove(root)
    return core, vocoder


class FixedCtrl:
    """Factory for an nn.Module standing in for Unit2Control: returns fixed control tensors."""
    def __new__(cls, d):
        import torch

        class _Fixed(torch.nn.Module):
            def __init__(self, d):
                super().__init__()
                self.d = d

            def forward(self, *a, **k):
                return self.d
        return _Fixed(d)


MODEL_SPECS = {   # ctor args and control-tensor names/widths of the three synthesizers (configs/*.yaml)
    'combsubfast': (lambda v, n_unit=256, n_spk=1: v.CombSubFast(44100, 512, n_unit, n_spk),
                    [('harmonic_magnitude', 513), ('harmonic_phase', 513), ('noise_magnitude', 513)]),
    'combsub': (lambda v, n_unit=256, n_spk=1: v.CombSub(44100, 512, 256, 512, 256, n_unit, n_spk),
                [('group_delay', 256), ('harmonic_magnitude', 512), ('noise_magnitude', 256)]),
    'sins': (lambda v, n_unit=256, n_spk=1: v.Sins(44100, 512, 128, 256, 256, n_unit, n_spk),
             [('amplitudes', 128), ('group_delay', 256), ('noise_magnitude', 256)]),
}
