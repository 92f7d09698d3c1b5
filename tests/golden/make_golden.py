"""Generate the golden fixtures under tests/golden/ by running the REFERENCE code itself.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box).  The reference's synthesizer modules are imported unmodified through an import shim for
the third-party modules that are not installed here, `unit2ctrl` is
swapped for a module returning fixed control tensors, and `torch.rand_like` is patched to
return a fixed U so the noise excitation is identical on both sides.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

Each fixture holds the inputs, the fp32 reference outputs (torch CPU) and the outputs of the
same reference code fed with float64 inputs ("ref64", the arbiter).
"""
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('DDSP_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)


def install_stubs():
    def stub(name):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m
        return m
    for n in ['pyworld', 'parselmouth', 'torchcrepe', 'resampy', 'sklearn', 'sklearn.cluster']:
        stub(n)
    sys.modules['sklearn.cluster'].KMeans = object
    ext = stub('extorch')

    class Conv1dEx(torch.nn.Conv1d):
        def __init__(self, *a, causal=False, **k):
            assert not causal
            super().__init__(*a, **k)

    class Transpose(torch.nn.Module):
        def __init__(self, a, b):
            super().__init__()
            self.a, self.b = a, b

        def forward(self, x):
            return x.transpose(self.a, self.b)
    ext.Conv1dEx, ext.Transpose = Conv1dEx, Transpose
    stub('fast_transformers')
    cp = stub('fast_transformers.causal_product')
    cp.CausalDotProduct = None


class FixedCtrl(torch.nn.Module):
    def __init__(self, d):
        super().__init__()
        self.d = d

    def forward(self, *a, **k):
        return self.d


def load_reference():
    install_stubs()
    sys.path.insert(0, REF)
    import ddsp.core as core          # noqa
    import ddsp.vocoder as vocoder    # noqa
    return core, vocoder


def run_model(model, ctrl_names, ctrl, f0_frames, U, dtype):
    """ctrl: (B,F,sumK) numpy; returns reference outputs as numpy."""
    sizes = [k for _, k in ctrl_names]
    ct = torch.from_numpy(ctrl).to(dtype)
    views = torch.split(ct, sizes, dim=-1)
    model.unit2ctrl = FixedCtrl({n: v for (n, _), v in zip(ctrl_names, views)})
    B, Fr = f0_frames.shape
    f0 = torch.from_numpy(f0_frames).to(dtype).unsqueeze(-1)
    units = torch.zeros(B, Fr, 4, dtype=dtype)
    vol = torch.zeros(B, Fr, dtype=dtype)
    spk = torch.ones(B, 1, dtype=torch.long)
    Ut = torch.from_numpy(U)
    orig = torch.rand_like
    torch.rand_like = lambda x: Ut.to(x.dtype)
    try:
        with torch.no_grad():
            signal, phase, (harm, noise) = model(units, f0, vol, spk)
    finally:
        torch.rand_like = orig
    return (signal.double().numpy(), phase.double().numpy(), harm.double().numpy(), noise.double().numpy())


def main():
    core, vocoder = load_reference()
    from ddsp_b200.synthetic import make_inputs   # the same generator tests and bench use

    out = {}
    # ---- core-level fixtures ---------------------------------------------------------
    rng = np.random.default_rng(7)
    x = (rng.random((3, 9, 5)) * 800).astype(np.float32)
    up = core.upsample(torch.from_numpy(x), 512).numpy()
    out['upsample_x'] = x
    out['upsample_y_sub'] = up[:, ::37, :]                      # subsampled, bit-exact check
    out['upsample_y_sum'] = up.astype(np.float64).sum(axis=1)
    fo = core.upsample(torch.from_numpy(x[:, :, :1]), 512).squeeze(-1)
    rot = core.fo_to_rot(fo, 44100, None, True).numpy()
    out['rot_precise_sub'] = rot[:, ::37]
    ip = torch.tensor([0.3, -2.0, 5.0])
    rot_ip = core.fo_to_rot(fo, 44100, ip, True).numpy()
    out['rot_ip'] = ip.numpy()
    out['rot_precise_ip_sub'] = rot_ip[:, ::37]
    amp = rng.random((2, 6, 128)).astype(np.float32)
    pitch = np.array([[[172.265625], [800.0], [65.0], [0.0], [344.53125], [440.0]]] * 2, dtype=np.float32)
    masked = core.remove_above_fmax(torch.from_numpy(amp), torch.from_numpy(pitch), torch.tensor(44100) / 2).numpy()
    out['mask_amp'], out['mask_pitch'], out['mask_out'] = amp, pitch, masked
    # frequency_filter pieces
    mags = np.exp(0.57 * rng.standard_normal((2, 5, 256))).astype(np.float32)
    audio = (rng.random((2, 5 * 512)) * 2 - 1).astype(np.float32)
    for name, kw in [('none', dict(hann_window=False)), ('hann', dict(hann_window=True))]:
        ir = core._frequency_impulse_response(torch.complex(torch.from_numpy(mags).double(), torch.zeros(2, 5, 256).double()), **kw).numpy()
        y = core.frequency_filter(torch.from_numpy(audio).double(), torch.complex(torch.from_numpy(mags).double(), torch.zeros(2, 5, 256).double()), **kw).numpy()
        out[f'ff_{name}_ir'] = ir
        out[f'ff_{name}_y'] = y
    mags2 = np.exp(0.57 * rng.standard_normal((2, 5, 512))).astype(np.float32)
    f0f = np.array([[[220.0], [65.0], [800.0], [0.0], [330.0]]] * 2, dtype=np.float64)
    hw = 1.5 * 44100 / (torch.from_numpy(f0f) + 1e-3)
    m2 = torch.complex(torch.from_numpy(mags2).double(), torch.zeros(2, 5, 512).double())
    out['ff_dyn_ir'] = core._frequency_impulse_response(m2, True, hw).numpy()
    out['ff_dyn_y'] = core.frequency_filter(torch.from_numpy(audio).double(), m2, True, hw).numpy()
    ph = rng.standard_normal((2, 5, 256))
    m3 = torch.exp(1j * torch.cumsum(np.pi * torch.tanh(torch.from_numpy(ph)), -1))
    out['ff_ap_y'] = core.frequency_filter(torch.from_numpy(audio).double(), m3, hann_window=False).numpy()
    out['ff_mags'], out['ff_mags2'], out['ff_audio'], out['ff_f0f'], out['ff_ph'] = mags, mags2, audio, f0f, ph
    np.savez_compressed(os.path.join(HERE, 'core.npz'), **out)

    # ---- model-level fixtures --------------------------------------------------------
    specs = {
        'combsubfast': (lambda: vocoder.CombSubFast(44100, 512, 4, 1),
                        [('harmonic_magnitude', 513), ('harmonic_phase', 513), ('noise_magnitude', 513)]),
        'combsub': (lambda: vocoder.CombSub(44100, 512, 256, 512, 256, 4, 1),
                    [('group_delay', 256), ('harmonic_magnitude', 512), ('noise_magnitude', 256)]),
        'sins': (lambda: vocoder.Sins(44100, 512, 128, 256, 256, 4, 1),
                 [('amplitudes', 128), ('group_delay', 256), ('noise_magnitude', 256)]),
    }
    cases = [  # (tag, B, F, seed, zero_f0_fraction, store_inputs)
        ('small', 2, 7, 11, 0.5, True),
        ('odd', 1, 12, 12, 0.0, True),
        ('gui', 1, 130, 13, 0.1, False),
    ]
    for mname, (ctor, names) in specs.items():
        model = ctor().eval()
        sumk = sum(k for _, k in names)
        for tag, B, Fr, seed, zf, store in cases:
            inp = make_inputs(B, Fr, sumk, seed=seed, zero_f0_fraction=zf)
            o32 = run_model(model, names, inp['ctrl'], inp['f0_frames'], inp['U'], torch.float32)
            o64 = run_model(model, names, inp['ctrl'], inp['f0_frames'], inp['U'], torch.float64)
            d = dict(B=B, F=Fr, seed=seed, zero_f0_fraction=zf,
                     signal32=o32[0].astype(np.float32), phase32=o32[1].astype(np.float32),
                     harm32=o32[2].astype(np.float32), noise32=o32[3].astype(np.float32),
                     signal64=o64[0].astype(np.float32), harm64=o64[2].astype(np.float32),
                     noise64=o64[3].astype(np.float32))
            # f32 storage of the fp64 run is enough: residual 6e-8 relative, far below tolerances
            if store:
                d.update(ctrl=inp['ctrl'], f0_frames=inp['f0_frames'], U=inp['U'])
            else:   # larger case: inputs are regenerated from the seed, keep only what is compared
                for k in ('harm64', 'noise64', 'noise32'):
                    d.pop(k)
                if mname == 'combsubfast':
                    d.pop('harm32')
                if mname == 'sins':
                    d['phase32'] = d['phase32'][:, ::64]
            np.savez_compressed(os.path.join(HERE, f'{mname}_{tag}.npz'), **d)
            print(mname, tag, 'max|ref32-ref64| =', np.abs(o32[0] - o64[0]).max())


if __name__ == '__main__':
    main()
