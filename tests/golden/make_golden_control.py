"""Golden fixtures of the WHOLE reference CombSubFast module (control network + synthesizer) and of the state_dict
layout of the three reference synthesizers.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_control.py

* state_dict_layout.json -- names, shapes and dtypes of `state_dict()` of the reference's Sins / CombSub /
  CombSubFast (ddsp/vocoder.py:372-550 with configs/*.yaml sizes): what a reference checkpoint contains.
* module_<tag>.npz -- inputs and outputs of `CombSubFast.forward` (vocoder.py:437-492) of the unmodified reference with
  the deterministic weights of `ddsp_b200.synthetic.synthetic_state_dict`, torch.rand_like patched to a fixed U;
  tags: xunit (n_unit 4, configs/combsub_xunit.yaml), yunit (n_unit 512, combsub_yunit.yaml), long (n_unit 256,
  more than 256 frames: the large-batch code path of the drop-in's control network).  The control rows are stored
  for every 9th frame.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.make_golden import load_reference   # noqa: E402


def main():
    _, vocoder = load_reference()
    from ddsp_b200.synthetic import make_f0, synthetic_state_dict
    layout = {}
    for name, ctor in [('Sins', lambda: vocoder.Sins(44100, 512, 128, 256, 256, 256, 1)),
                       ('CombSub', lambda: vocoder.CombSub(44100, 512, 256, 512, 256, 256, 1)),
                       ('CombSubFast', lambda: vocoder.CombSubFast(44100, 512, 256, 1))]:
        sd = ctor().state_dict()
        layout[name] = {k: [list(v.shape), str(v.dtype)] for k, v in sd.items()}
    with open(os.path.join(HERE, 'state_dict_layout.json'), 'w') as f:
        json.dump(layout, f, indent=0, sort_keys=True)

    for tag, n_unit, n_spk, B, Fr, seed in [('xunit', 4, 2, 2, 24, 41), ('yunit', 512, 1, 1, 17, 42), ('long', 256, 3, 1, 300, 43)]:
        model = vocoder.CombSubFast(44100, 512, n_unit, n_spk).eval()
        model.load_state_dict(synthetic_state_dict(model.state_dict(), seed=seed), strict=True)
        rng = np.random.default_rng(seed)
        units = rng.standard_normal((B, Fr, n_unit)).astype(np.float32)
        f0 = make_f0(B, Fr, rng, zero_f0_fraction=0.1)
        vol = rng.random((B, Fr)).astype(np.float32)
        spk = rng.integers(1, n_spk + 1, size=(B, 1)).astype(np.int64)
        U = rng.random((B, Fr * 512)).astype(np.float32)
        cap = {}
        model.unit2ctrl.register_forward_hook(lambda m, i, o: cap.update(o))
        orig = torch.rand_like
        torch.rand_like = lambda x: torch.from_numpy(U).to(x.dtype)
        try:
            with torch.no_grad():
                sig, ph, _ = model(torch.from_numpy(units), torch.from_numpy(f0).unsqueeze(-1), torch.from_numpy(vol),
                                   torch.from_numpy(spk))
        finally:
            torch.rand_like = orig
        ctrl = torch.cat([cap['harmonic_magnitude'], cap['harmonic_phase'], cap['noise_magnitude']], dim=-1).numpy()
        np.savez_compressed(os.path.join(HERE, f'module_{tag}.npz'), n_unit=n_unit, n_spk=n_spk, seed=seed, units=units,
                            f0_frames=f0, volume=vol, spk_id=spk, U=U, signal=sig.numpy(), phase_frames=ph.numpy()[..., 0],
                            ctrl_sub=ctrl[:, ::9])
        print(tag, 'signal peak', np.abs(sig.numpy()).max(), 'ctrl range', ctrl.min(), ctrl.max())


if __name__ == '__main__':
    main()
