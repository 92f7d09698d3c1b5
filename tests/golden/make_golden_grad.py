"""Golden GRADIENTS of the reference CombSubFast (ddsp/vocoder.py:437-492) w.r.t. its control tensors,
produced by autograd through the UNMODIFIED reference module (same stubs / FixedCtrl / patched
rand_like as make_golden.py).  Run here, in the build container (needs /root/reference):

    python tests/golden/make_golden_grad.py

Loss = sum(signal * R) with a fixed random R, i.e. dL/dsignal = R.  Stored per case: inputs, R, and
the gradients from an fp32 and an fp64 run (the latter cast to fp32) -> combsubfast_grad_<tag>.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.make_golden import FixedCtrl, load_reference   # noqa: E402

NAMES = [('harmonic_magnitude', 513), ('harmonic_phase', 513), ('noise_magnitude', 513)]


def grads_of(model, ctrl, f0_frames, U, R, dtype):
    ct = torch.from_numpy(ctrl).to(dtype).requires_grad_(True)
    views = torch.split(ct, [k for _, k in NAMES], dim=-1)
    model.unit2ctrl = FixedCtrl({n: v for (n, _), v in zip(NAMES, views)})
    B, Fr = f0_frames.shape
    f0 = torch.from_numpy(f0_frames).to(dtype).unsqueeze(-1)
    Ut = torch.from_numpy(U)
    orig = torch.rand_like
    torch.rand_like = lambda x: Ut.to(x.dtype)
    try:
        signal, _, _ = model(torch.zeros(B, Fr, 4, dtype=dtype), f0, torch.zeros(B, Fr, dtype=dtype),
                             torch.ones(B, 1, dtype=torch.long), infer=True)
    finally:
        torch.rand_like = orig
    (signal * torch.from_numpy(R).to(dtype)).sum().backward()
    return ct.grad.double().numpy()


def main():
    _, vocoder = load_reference()
    from ddsp_b200.synthetic import make_inputs
    model = vocoder.CombSubFast(44100, 512, 4, 1)
    for tag, B, Fr, seed, zf in [('small', 2, 7, 21, 0.3), ('even', 1, 12, 22, 0.0), ('one', 2, 1, 23, 0.0)]:
        inp = make_inputs(B, Fr, 1539, seed=seed, zero_f0_fraction=zf)
        R = np.random.default_rng(seed + 1000).standard_normal((B, Fr * 512)).astype(np.float32)
        g32 = grads_of(model, inp['ctrl'], inp['f0_frames'], inp['U'], R, torch.float32)
        g64 = grads_of(model, inp['ctrl'], inp['f0_frames'], inp['U'], R, torch.float64)
        np.savez_compressed(os.path.join(HERE, f'combsubfast_grad_{tag}.npz'), ctrl=inp['ctrl'],
                            f0_frames=inp['f0_frames'], U=inp['U'], R=R, grad32=g32.astype(np.float32),
                            grad64=g64.astype(np.float32))
        print(tag, 'max|g32-g64| =', np.abs(g32 - g64).max(), 'max|g64| =', np.abs(g64).max())


FILTER_MODELS = {
    'combsub': (lambda v: v.CombSub(44100, 512, 256, 512, 256, 4, 1),
                [('group_delay', 256), ('harmonic_magnitude', 512), ('noise_magnitude', 256)]),
    'sins': (lambda v: v.Sins(44100, 512, 128, 256, 256, 4, 1),
             [('amplitudes', 128), ('group_delay', 256), ('noise_magnitude', 256)]),
}


def filter_model_grads(model, names, ctrl, f0_frames, U, R, dtype):
    """Gradients of sum(signal * R) w.r.t. the packed control tensor, autograd through the reference module."""
    ct = torch.from_numpy(ctrl).to(dtype).requires_grad_(True)
    views = torch.split(ct, [k for _, k in names], dim=-1)
    model.unit2ctrl = FixedCtrl({n: v for (n, _), v in zip(names, views)})
    B, Fr = f0_frames.shape
    f0 = torch.from_numpy(f0_frames).to(dtype).unsqueeze(-1)
    Ut = torch.from_numpy(U)
    orig = torch.rand_like
    torch.rand_like = lambda x: Ut.to(x.dtype)
    try:
        signal, _, _ = model(torch.zeros(B, Fr, 4, dtype=dtype), f0, torch.zeros(B, Fr, dtype=dtype),
                             torch.ones(B, 1, dtype=torch.long), infer=True)
    finally:
        torch.rand_like = orig
    (signal * torch.from_numpy(R).to(dtype)).sum().backward()
    return ct.grad.double().numpy(), signal.detach().double().numpy()


def main_filter_models():
    """combsub_grad_small.npz / sins_grad_small.npz: gradients of the frequency_filter synthesizers (vocoder.py:381-423,
    :504-550) from the unmodified reference, fp64 run stored as fp32."""
    _, vocoder = load_reference()
    from ddsp_b200.synthetic import make_inputs
    for mname, (ctor, names) in FILTER_MODELS.items():
        model = ctor(vocoder)
        sumk = sum(k for _, k in names)
        B, Fr, seed = 2, 5, 31
        inp = make_inputs(B, Fr, sumk, seed=seed, zero_f0_fraction=0.2)
        R = np.random.default_rng(seed + 1000).standard_normal((B, Fr * 512)).astype(np.float32)
        g64, s64 = filter_model_grads(model, names, inp['ctrl'], inp['f0_frames'], inp['U'], R, torch.float64)
        g32, _ = filter_model_grads(model, names, inp['ctrl'], inp['f0_frames'], inp['U'], R, torch.float32)
        np.savez_compressed(os.path.join(HERE, f'{mname}_grad_small.npz'), ctrl=inp['ctrl'], f0_frames=inp['f0_frames'],
                            U=inp['U'], R=R, grad64=g64.astype(np.float32), signal64=s64.astype(np.float32))
        print(mname, 'max|g32-g64| =', np.abs(g32 - g64).max(), 'max|g64| =', np.abs(g64).max())


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'filter_models':
        main_filter_models()
    else:
        main()
