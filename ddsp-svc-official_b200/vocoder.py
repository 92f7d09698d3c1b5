"""Drop-in synthesizer modules: same constructors, `forward` signature, return tuple and
state_dict keys as the reference's `ddsp/vocoder.py:343-550`, with the DSP (everything except
`Unit2Control`) running as hand-written sm_100a kernels.

`unit2ctrl` (the control network) is the state_dict-compatible `ddsp_b200.control.Unit2Control` (fused kernels and
tensor-core Linears under no_grad, plain torch ops under autograd), or any module passed as `unit2ctrl=` with the
same call signature returning the dict of control tensors.
"""
import os

import torch
import yaml

from . import core


class DotDict(dict):
    """Attribute access to nested config dicts (reference vocoder.py:335-341)."""

    def __getattr__(*args):
        val = dict.get(*args)
        return DotDict(val) if type(val) is dict else val

    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__


def _make_unit2ctrl(n_unit, n_spk, output_splits, c):
    """The control network: `ddsp_b200.control.Unit2Control`, state_dict-compatible with the reference's
    `ddsp/unit2control.py` + `ddsp/pcmer.py` (strict load; pinned to the reference in tests/test_module_parity.py) with its
    fused / tensor-core inference path.  DDSP_B200_REFERENCE_CONTROL=1 selects the reference's own class instead when the
    reference repo (and its extorch / fast-transformers dependencies) is importable."""
    if os.environ.get('DDSP_B200_REFERENCE_CONTROL') == '1':
        from ddsp.unit2control import Unit2Control      # the reference's control network
    else:
        from .control import Unit2Control
    return Unit2Control(n_unit, n_spk, output_splits, c)


class _SynthBase(torch.nn.Module):
    def __init__(self, sampling_rate, block_size):
        super().__init__()
        # same buffers as the reference (state_dict keys `sampling_rate`, `block_size`), plus
        # python ints so that forward never syncs on `.item()` (18 syncs in the reference)
        self.register_buffer('sampling_rate', torch.tensor(sampling_rate))
        self.register_buffer('block_size', torch.tensor(block_size))
        self._sr = int(sampling_rate)
        self._hop = int(block_size)
        self._noise_calls = 0
        self._seed_device = None        # set by GraphedForward: device counter added to the noise seed per replay

    _instances = 0          # every module draws from its own seed sequence (two models with equal call counts must differ)

    def _next_seed(self):
        if self._noise_calls == 0:
            _SynthBase._instances += 1
            self._instance = _SynthBase._instances
        self._noise_calls += 1
        return ((torch.initial_seed() * 0x9E3779B1 + self._instance) * 0x2545F4914F6CDD1D + self._noise_calls) & ((1 << 62) - 1)

    @staticmethod
    def _wants_grad(ctrls):
        return torch.is_grad_enabled() and any(t.requires_grad for t in ctrls.values())


class _FilterStageB(torch.autograd.Function):
    """Stage B of the `frequency_filter` synthesizers (Sins, CombSub-old) under autograd, so that the drop-in modules
    train (solver.py:111-113).  Forward: the hand-written kernels.  Backward: stage B is re-evaluated with the stock
    torch ops of `ddsp_b200.diffsynth` (the same function, pinned to the reference's own gradients in
    tests/test_diffsynth.py) under autograd and its gradients w.r.t. the three control tensors are returned; f0,
    phase and the (explicitly drawn) noise are data, as in the reference graph."""

    @staticmethod
    def forward(ctx, kind, c0, c1, c2, f0_frames, aux, phase_full, hop, sr, noise_u):
        if kind == 'sins':
            outs = core.sins_stage(c0, c1, c2, f0_frames, phase_full, hop, sr, noise_u=noise_u)
        else:
            outs = core.combsub_stage(c0, c1, c2, f0_frames, aux, hop, sr, noise_u=noise_u)
        ctx.save_for_backward(c0, c1, c2, f0_frames, phase_full, noise_u)
        ctx.cfg = (kind, hop, sr)
        return outs

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_signal, g_harmonic, g_noise):
        from . import diffsynth
        c0, c1, c2, f0_frames, phase_full, noise_u = ctx.saved_tensors
        kind, hop, sr = ctx.cfg
        with torch.enable_grad():
            p = [t.detach().requires_grad_(True) for t in (c0, c1, c2)]
            if kind == 'sins':
                outs = diffsynth.sins_stage(p[0], p[1], p[2], f0_frames, phase_full, noise_u, hop, sr)
            else:
                rot = phase_full / (2 * 3.141592653589793)
                outs = diffsynth.combsub_stage(p[0], p[1], p[2], f0_frames, rot, noise_u, hop, sr)
            grads = [g for g in (g_signal, g_harmonic, g_noise)]
            pairs = [(o, g) for o, g in zip(outs, grads) if g is not None]
            torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])
        return (None, p[0].grad, p[1].grad, p[2].grad, None, None, None, None, None, None)


class _CombSubFastStageB(torch.autograd.Function):
    """Stage B of CombSubFast with its hand-written gradient, so that the drop-in module trains
    (solver.py:111-113: `model(..., infer=False)` -> loss -> `backward()`).  Gradients flow to the three
    control tensors only; f0 / phase / noise are data, as in the reference graph."""

    @staticmethod
    def forward(ctx, hm, hp, nm, f0_frames, prefix, hop, sr, initial_phase, noise_u, seed, window):
        signal = core.combsubfast_stage(hm, hp, nm, f0_frames, prefix, hop, sr, initial_phase, noise_u=noise_u,
                                        seed=seed, window=window)
        ctx.save_for_backward(hm, hp, nm, f0_frames, prefix, noise_u, window)
        ctx.cfg = (hop, sr, seed)
        return signal

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_signal):
        hm, hp, nm, f0_frames, prefix, noise_u, window = ctx.saved_tensors
        hop, sr, seed = ctx.cfg
        ghm, ghp, gnm = core.combsubfast_backward_stage(grad_signal, hm, hp, nm, f0_frames, prefix, hop, sr,
                                                        noise_u=noise_u, seed=seed, window=window)
        return ghm, ghp, gnm, None, None, None, None, None, None, None, None


class CombSubFast(_SynthBase):
    """Reference: ddsp/vocoder.py:426-492."""

    def __init__(self, sampling_rate, block_size, n_unit=256, n_spk=1, c: bool = False, unit2ctrl=None):
        super().__init__(sampling_rate, block_size)
        print(' [DDSP Model] Combtooth Subtractive Synthesiser (ddsp_b200)')
        self.register_buffer('window', torch.sqrt(torch.hann_window(2 * block_size)))
        splits = {'harmonic_magnitude': block_size + 1, 'harmonic_phase': block_size + 1,
                  'noise_magnitude': block_size + 1}
        self.unit2ctrl = unit2ctrl if unit2ctrl is not None else _make_unit2ctrl(n_unit, n_spk, splits, c)

    def forward(self, units_frames, f0_frames, volume_frames, spk_id, spk_mix_dict=None, initial_phase=None,
                infer=True, noise_u=None, **kwargs):
        # stage A: vocoder.py:449-451
        f0_frames = core.as_f32(f0_frames)
        phase_frames, prefix, _ = core.phase_stage(f0_frames, self._hop, self._sr, initial_phase, infer)
        # control network (reference PyTorch): vocoder.py:454
        ctrls = self.unit2ctrl(units_frames, f0_frames, phase_frames, volume_frames, spk_id, spk_mix_dict=spk_mix_dict)
        # stage B: vocoder.py:455-490
        ctrls = {k: core.as_f32(v) for k, v in ctrls.items()}
        args = (ctrls['harmonic_magnitude'], ctrls['harmonic_phase'], ctrls['noise_magnitude'], f0_frames, prefix,
                self._hop, self._sr, initial_phase)
        if torch.is_grad_enabled() and any(t.requires_grad for t in ctrls.values()):
            signal = _CombSubFastStageB.apply(*args, noise_u, self._next_seed(), self.window)
        else:
            signal = core.combsubfast_stage(*args, noise_u=noise_u, seed=self._next_seed(), window=self.window,
                                            seed_device=self._seed_device)
        return signal, phase_frames.unsqueeze(-1), (signal, signal)      # vocoder.py:492


class CombSub(_SynthBase):
    """Reference: ddsp/vocoder.py:495-550 (the "old" combtooth subtractive synthesiser)."""

    def __init__(self, sampling_rate, block_size, n_mag_allpass, n_mag_harmonic, n_mag_noise, n_unit=256, n_spk=1,
                 c: bool = False, unit2ctrl=None):
        super().__init__(sampling_rate, block_size)
        print(' [DDSP Model] Combtooth Subtractive Synthesiser (Old Version) (ddsp_b200)')
        splits = {'group_delay': n_mag_allpass, 'harmonic_magnitude': n_mag_harmonic, 'noise_magnitude': n_mag_noise}
        self.unit2ctrl = unit2ctrl if unit2ctrl is not None else _make_unit2ctrl(n_unit, n_spk, splits, c)

    def forward(self, units_frames, f0_frames, volume_frames, spk_id, spk_mix_dict=None, initial_phase=None,
                infer=True, noise_u=None, **kwargs):
        f0_frames = core.as_f32(f0_frames)
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.unit2ctrl.parameters())
        phase_frames, prefix, phase_full = core.phase_stage(f0_frames, self._hop, self._sr, initial_phase, infer,
                                                            full_rate=grad)                                   # :515-517
        ctrls = self.unit2ctrl(units_frames, f0_frames, phase_frames, volume_frames, spk_id, spk_mix_dict=spk_mix_dict)
        ctrls = {k: core.as_f32(v) for k, v in ctrls.items()}
        if self._wants_grad(ctrls):
            if phase_full is None:
                phase_full = core.phase_stage(f0_frames, self._hop, self._sr, initial_phase, infer, full_rate=True)[2]
            if noise_u is None:     # torch.rand_like of vocoder.py:545, drawn here so that forward and backward share it
                noise_u = torch.rand(f0_frames.shape[0], f0_frames.shape[1] * self._hop, device=f0_frames.device)
            signal, harmonic, noise = _FilterStageB.apply('combsub', ctrls['group_delay'], ctrls['harmonic_magnitude'],
                                                          ctrls['noise_magnitude'], f0_frames, prefix, phase_full,
                                                          self._hop, self._sr, noise_u)
            return signal, phase_frames.unsqueeze(-1), (harmonic, noise)
        signal, harmonic, noise = core.combsub_stage(ctrls['group_delay'], ctrls['harmonic_magnitude'],
                                                     ctrls['noise_magnitude'], f0_frames, prefix, self._hop, self._sr,
                                                     noise_u=noise_u, seed=self._next_seed())                  # :521-548
        return signal, phase_frames.unsqueeze(-1), (harmonic, noise)                                          # :550


class Sins(_SynthBase):
    """Reference: ddsp/vocoder.py:372-423 (sinusoids additive synthesiser)."""

    def __init__(self, sampling_rate, block_size, n_harmonics, n_mag_allpass, n_mag_noise, n_unit=256, n_spk=1,
                 c: bool = False, unit2ctrl=None):
        super().__init__(sampling_rate, block_size)
        print(' [DDSP Model] Sinusoids Additive Synthesiser (ddsp_b200)')
        splits = {'amplitudes': n_harmonics, 'group_delay': n_mag_allpass, 'noise_magnitude': n_mag_noise}
        self.unit2ctrl = unit2ctrl if unit2ctrl is not None else _make_unit2ctrl(n_unit, n_spk, splits, c)

    def forward(self, units_frames, f0_frames, volume_frames, spk_id, spk_mix_dict=None, initial_phase=None,
                infer=True, max_upsample_dim=32, noise_u=None):
        # stage A with the full-rate phase (vocoder.py:391-393); `max_upsample_dim` only bounded the
        # reference's (B,T,32) temporaries and has no effect here
        f0_frames = core.as_f32(f0_frames)
        phase_frames, _, phase = core.phase_stage(f0_frames, self._hop, self._sr, initial_phase, infer, full_rate=True)
        ctrls = self.unit2ctrl(units_frames, f0_frames, phase_frames, volume_frames, spk_id, spk_mix_dict=spk_mix_dict)
        ctrls = {k: core.as_f32(v) for k, v in ctrls.items()}
        if self._wants_grad(ctrls):
            if noise_u is None:     # torch.rand_like of vocoder.py:418
                noise_u = torch.rand(phase.shape, device=phase.device)
            signal, harmonic, noise = _FilterStageB.apply('sins', ctrls['amplitudes'], ctrls['group_delay'],
                                                          ctrls['noise_magnitude'], f0_frames, None, phase, self._hop,
                                                          self._sr, noise_u)
            return signal, phase.unsqueeze(-1), (harmonic, noise)
        signal, harmonic, noise = core.sins_stage(ctrls['amplitudes'], ctrls['group_delay'], ctrls['noise_magnitude'],
                                                  f0_frames, phase, self._hop, self._sr, noise_u=noise_u,
                                                  seed=self._next_seed())                                     # :397-421
        return signal, phase.unsqueeze(-1), (harmonic, noise)                                                 # :423


class GraphedForward:
    """`fast = GraphedForward(model)`; `fast(units, f0, volume, spk_id, spk_mix_dict=None)` returns what
    `model(...)` returns under `torch.no_grad()`, but replays one CUDA graph per input shape instead of
    launching the ~60 kernels of control network + synthesizer one by one -- the per-block cost of the GUI
    callback (gui.py:125-127) drops from ~0.95 ms of host time to one graph launch (~0.3-0.4 ms on the
    device for 0.1-1.5 s blocks).

    Inputs are copied into static buffers, so any caller tensors work; outputs are fresh clones unless
    `copy_outputs=False` (then they are the graph's own buffers, valid until the next call with that
    shape).  The noise differs on every replay: CombSubFast adds a device-side counter to its in-kernel
    noise seed, Sins / CombSub (old) are fed a `noise_u` tensor refilled by `torch.rand` inside the graph.
    `spk_mix_dict` values are baked in at capture (part of the cache key)."""

    def __init__(self, model, copy_outputs=True, max_graphs=8):
        self.model = model
        self.copy_outputs = bool(copy_outputs)
        self.max_graphs = int(max_graphs)
        self._graphs = {}

    def _capture(self, units, f0, volume, spk_id, spk_mix_dict):
        model = self.model
        dev = f0.device
        static = {'units': units.clone(), 'f0': f0.clone(), 'volume': volume.clone(),
                  'spk_id': None if spk_id is None else spk_id.clone()}
        is_fast = isinstance(model, CombSubFast)
        counter = torch.zeros(1, dtype=torch.int64, device=dev) if is_fast else None
        noise = None if is_fast else torch.empty((f0.shape[0], f0.shape[1] * model._hop), dtype=torch.float32, device=dev)
        static['counter'], static['noise'] = counter, noise      # the graph writes them on every replay: keep them alive

        def run():
            kw = {}
            if is_fast:
                counter.add_(1)
            else:
                noise.uniform_()                  # torch.rand_like of vocoder.py:418,545, graph-safe generator
                kw['noise_u'] = noise
            return model(static['units'], static['f0'], static['volume'], static['spk_id'],
                         spk_mix_dict=spk_mix_dict, **kw)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        prev = model._seed_device
        model._seed_device = counter
        try:
            with torch.cuda.stream(side), torch.no_grad():
                run()                             # warm-up outside the capture (lazy tables, cuBLAS workspaces)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    out = run()
        finally:
            model._seed_device = prev
        torch.cuda.current_stream(dev).wait_stream(side)
        return graph, static, out

    def __call__(self, units_frames, f0_frames, volume_frames, spk_id=None, spk_mix_dict=None):
        key = (tuple(units_frames.shape), tuple(f0_frames.shape), tuple(volume_frames.shape), units_frames.dtype,
               None if spk_id is None else tuple(spk_id.shape),
               None if spk_mix_dict is None else tuple(sorted(spk_mix_dict.items())), f0_frames.device.index)
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            entry = self._graphs[key] = self._capture(units_frames, f0_frames, volume_frames, spk_id, spk_mix_dict)
        graph, static, out = entry
        static['units'].copy_(units_frames)
        static['f0'].copy_(f0_frames)
        static['volume'].copy_(volume_frames)
        if spk_id is not None:
            static['spk_id'].copy_(spk_id)
        graph.replay()
        if not self.copy_outputs:
            return out
        signal, phase, (harmonic, noise) = out
        s = signal.clone()
        if harmonic is signal:                     # CombSubFast returns (signal, signal) (vocoder.py:492)
            return s, phase.clone(), (s, s)
        return s, phase.clone(), (harmonic.clone(), noise.clone())


def load_model(model_path, device='cuda'):
    """Reference: ddsp/vocoder.py:343-369 (same config.yaml + checkpoint layout)."""
    config_file = os.path.join(os.path.split(model_path)[0], 'config.yaml')
    with open(config_file, 'r') as config:
        args = DotDict(yaml.safe_load(config))
    if args.model.type == 'Sins':
        model = Sins(sampling_rate=args.data.sampling_rate, block_size=args.data.block_size,
                     n_harmonics=args.model.n_harmonics, n_mag_allpass=args.model.n_mag_allpass,
                     n_mag_noise=args.model.n_mag_noise, n_unit=args.data.encoder_out_channels,
                     n_spk=args.model.n_spk, c=args.model.c)
    elif args.model.type == 'CombSub':
        model = CombSub(sampling_rate=args.data.sampling_rate, block_size=args.data.block_size,
                        n_mag_allpass=args.model.n_mag_allpass, n_mag_harmonic=args.model.n_mag_harmonic,
                        n_mag_noise=args.model.n_mag_noise, n_unit=args.data.encoder_out_channels,
                        n_spk=args.model.n_spk, c=args.model.c)
    elif args.model.type == 'CombSubFast':
        model = CombSubFast(sampling_rate=args.data.sampling_rate, block_size=args.data.block_size,
                            n_unit=args.data.encoder_out_channels, n_spk=args.model.n_spk, c=args.model.c)
    else:
        raise ValueError(f' [x] Unknown Model: {args.model.type}')
    print(' [Loading] ' + model_path)
    ckpt = torch.load(model_path, map_location=torch.device(device))
    model.to(device)
    model.load_state_dict(ckpt['model'])
    model.eval()
    return model, args
