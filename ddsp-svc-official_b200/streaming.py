"""Streaming CombSubFast: blocks of frames that continue one another.

The reference GUI (gui.py:373-388) re-synthesises its whole window -- the new block plus crossfade and
`extra_time` context -- from phase 0 on every block and hides the phase jump with a SOLA splice
(gui.py:408-426).  Here the synthesizer carries its state instead:

* the fp64 phase prefix of stage A (`ddsp_b200_phase_stream`), so the comb never jumps;
* the last `CONTEXT` = 3 frames of f0 and control rows, which are synthesised again at the head of the
  next block, because output hop h is final only once frames h .. h+2 are known (hop h overlap-adds
  STFT frames h and h+1, frame h+1 spans excitation hops h and h+1, and hop h+1 interpolates f0 towards
  frame h+2): every push of k frames returns k finished hops, `LATENCY` = 2 hops (23 ms) behind its input;
* the hop index of the in-kernel noise (`ddsp_b200_combsubfast_stream`), so re-synthesised hops get the
  same noise.

What comes out is the signal ONE call over all frames pushed so far would have produced -- same phase,
same excitation, same noise; equal to the last fp32 ulp or two (measured 7e-8) rather than bitwise, because
the inverse FFT handles STFT frames in pairs and which frames share a pair depends on the parity of a
block's first frame (tests/test_gpu_stream.py); `flush()` hands out the last two hops with the
reference's hold-last ending.
`FilterModelStream` (below) does the same for the `frequency_filter` models Sins and CombSub (old), whose (L-1)-tap
filters reach further: 1 + 2 (Sins) or 2 + 3 (CombSub) frames are re-synthesised around every block.
"""
import torch

from . import core

CONTEXT = 3        # frames re-synthesised at the head of every block
LATENCY = 2        # hops between the newest frame pushed and the newest hop returned


class CombSubFastStream:
    """Stage-level stream (control rows in, samples out) for `B` parallel clips.

        ph = stream.begin(f0_new)                      # (B,k) phase of the new frames, input of the control network
        audio = stream.finish(hm, hp, nm)              # (B, 512*n) finished samples, n = k (k-2 on the first push)
        audio = stream.push(hm, hp, nm, f0_new)        # both steps, when the rows do not depend on the phase
        tail = stream.flush()                          # the last 2 hops; the stream is reset afterwards
    """

    def __init__(self, block_size=512, sampling_rate=44100, window=None, seed=0, initial_phase=None):
        self.hop = int(block_size)
        self.sr = float(sampling_rate)
        self.window = window
        self.seed = int(seed)
        self._utterance = -1
        self.initial_phase = initial_phase
        self.reset()

    def reset(self):
        # a new utterance draws new noise: the stream seed moves on every reset() / flush() (re-synthesised hops inside one
        # utterance keep their noise through the hop index, not through the seed)
        self._utterance += 1
        self._seed_now = (self.seed + 0x9E3779B97F4A7C15 * self._utterance) & ((1 << 62) - 1)
        self.frames_pushed = 0          # frames handed to finish() so far
        self.hops_emitted = 0
        # f0 and control rows of the stream live in one linear device buffer each; the window of a block is
        # the slice [end - t, end + k): its t <= CONTEXT tail frames are already in place, so a push costs one
        # copy of the new f0 and one of the new rows (no concatenation).  When the buffer is full the tail
        # moves back to the front.
        self._f0_buf = None             # (B,cap)
        self._rows_buf = None           # (B,cap,3*(hop+1))
        self._end = 0                   # frames of the buffer in use
        self._t = 0                     # tail frames (<= CONTEXT) ending at _end
        self._noise_tail = None         # (B,t*hop) injected noise of the tail, parity mode only
        self._carry = None              # (B,) fp64 view: prefix at the first tail frame
        self._pending = None            # state of a begin() waiting for its finish()
        self._last_tail = None          # (B,2*hop) hold-last ending of the newest block, for flush()

    def _room(self, B, k, device):
        """Make sure k more frames fit behind the tail."""
        t, K3 = self._t, 3 * (self.hop + 1)
        if self._f0_buf is not None and self._f0_buf.shape[0] != B:
            raise ValueError('the number of clips must not change inside a stream (reset() first)')
        if self._f0_buf is not None and self._end + k <= self._f0_buf.shape[1]:
            return
        cap = max(64, 4 * (k + CONTEXT))
        if self._f0_buf is not None and t + k <= self._f0_buf.shape[1] and self._end - t >= t:
            f0_buf, rows_buf = self._f0_buf, self._rows_buf            # same buffers, tail back to the front
        else:
            f0_buf = torch.empty((B, cap), dtype=torch.float32, device=device)
            rows_buf = torch.empty((B, cap, K3), dtype=torch.float32, device=device)
        if t:
            f0_buf[:, :t] = self._f0_buf[:, self._end - t:self._end]
            rows_buf[:, :t] = self._rows_buf[:, self._end - t:self._end]
        self._f0_buf, self._rows_buf, self._end = f0_buf, rows_buf, t

    # ---------------------------------------------------------------------------------------------
    def begin(self, f0_new):
        """Stage A for the k new frames `f0_new` (B,k) or (B,k,1).  Returns their frame-rate phase (B,k) --
        the `phase_frames` input of Unit2Control (vocoder.py:451,454)."""
        f0_new = core._f0_2d(core.as_f32(f0_new))
        B, k = f0_new.shape
        if k < 1:
            raise ValueError('begin() needs at least one new frame')
        self._room(B, k, f0_new.device)
        t, end = self._t, self._end
        self._f0_buf[:, end:end + k] = f0_new
        f0_win = self._f0_buf[:, end - t:end + k]
        phase_win, prefix = core.phase_stage_stream(f0_win, self.hop, self.sr, carry=self._carry,
                                                    initial_phase=self.initial_phase if self._carry is None else None)
        self._pending = (f0_win, prefix, t)
        return phase_win[:, t:]

    def finish(self, harmonic_magnitude, harmonic_phase, noise_magnitude, noise_u=None):
        """Stage B for the frames given to `begin`: control rows (B,k,513) each (any strides) and, in parity
        mode, their noise (B,k*512).  Returns the finished samples (B, 512*n)."""
        if self._pending is None:
            raise RuntimeError('finish() without begin()')
        f0_win, prefix, t = self._pending
        B, W = f0_win.shape
        k = W - t
        K = self.hop + 1
        rows = (harmonic_magnitude, harmonic_phase, noise_magnitude)
        for r in rows:
            if tuple(r.shape) != (B, k, K):
                raise ValueError(f'control rows must be (B, {k}, {K}); got {tuple(r.shape)}')
        end = self._end
        dst = self._rows_buf[:, end:end + k]
        hm_, hp_, nm_ = rows
        if (all(r.dtype == torch.float32 and r.is_cuda for r in rows) and hm_.stride() == hp_.stride() == nm_.stride() and hm_.stride(2) == 1
                and hp_.data_ptr() - hm_.data_ptr() == 4 * K and nm_.data_ptr() - hp_.data_ptr() == 4 * K):
            # the usual case: the three tensors are the split views of one (B,k,1539) tensor -> one copy
            dst.copy_(hm_.as_strided((B, k, 3 * K), hm_.stride()))
        else:
            for i, r in enumerate(rows):
                dst[:, :, i * K:(i + 1) * K] = core.as_f32(r)
        rows_win = self._rows_buf[:, end - t:end + k]
        if (noise_u is None) != (self._noise_tail is None) and t:
            raise ValueError('either every block of a stream injects noise_u or none does')
        noise_win = None
        if noise_u is not None:
            noise_u = core._need_cuda_f32(noise_u, 'noise_u')
            if tuple(noise_u.shape) != (B, k * self.hop):
                raise ValueError('noise_u must be (B, k*block_size)')
            noise_win = noise_u.contiguous() if t == 0 else torch.cat((self._noise_tail, noise_u), dim=1)
        hm, hp, nm = torch.split(rows_win, K, dim=-1)
        first_hop = self.frames_pushed - t                 # stream index of the window's first hop
        signal = core.combsubfast_stage(hm, hp, nm, f0_win, prefix, self.hop, self.sr, noise_u=noise_win,
                                        seed=self._seed_now, window=self.window, hop_offset=first_hop)
        # hop j of the window is final for j <= W-3 (and for j >= 1 unless the window starts the stream: lo is
        # 1 in the steady state, 0 while the window still begins at frame 0)
        lo = self.hops_emitted - first_hop
        hi = max(lo, W - LATENCY)
        out = signal[:, lo * self.hop:hi * self.hop]
        self._last_tail = signal[:, hi * self.hop:]
        # state for the next block: its window starts at frame W - CONTEXT of this one
        keep = min(CONTEXT, W)
        self._end = end + k
        self._t = keep
        self._noise_tail = None if noise_win is None else noise_win[:, (W - keep) * self.hop:]
        self._carry = prefix[:, W - keep]
        self.frames_pushed += k
        self.hops_emitted += hi - lo
        self._pending = None
        return out

    def push(self, harmonic_magnitude, harmonic_phase, noise_magnitude, f0_new, noise_u=None):
        ops = core._torchext.ops()
        fast = ops is not None and noise_u is None and self._noise_tail is None and self._pending is None and all(
            isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32
            for x in (f0_new, harmonic_magnitude, harmonic_phase, noise_magnitude))
        if not fast:
            self.begin(f0_new)
            return self.finish(harmonic_magnitude, harmonic_phase, noise_magnitude, noise_u=noise_u)
        # the whole block -- both buffer copies, stage A from the carried prefix, stage B -- in ONE operator of the
        # extension host (csrc/torch_ext.cpp csf_stream_push); what remains here is the stream's bookkeeping
        B, k = f0_new.shape[0], f0_new.shape[1]
        if k < 1:
            raise ValueError('push() needs at least one new frame')
        self._room(B, k, f0_new.device)
        t, end = self._t, self._end
        first_hop = self.frames_pushed - t
        ip = None
        if self._carry is None and self.initial_phase is not None:
            ip = torch.as_tensor(self.initial_phase, dtype=torch.float32, device=f0_new.device)
        signal, _, prefix = core._op(ops.csf_stream_push, self._f0_buf, self._rows_buf, f0_new, harmonic_magnitude,
                                     harmonic_phase, noise_magnitude, end, t, self.hop, self.sr, ip, self._carry,
                                     self._seed_now, self.window, first_hop)
        W = t + k
        lo = self.hops_emitted - first_hop
        hi = max(lo, W - LATENCY)
        hop = self.hop
        out = signal[:, lo * hop:hi * hop]
        self._last_tail = signal[:, hi * hop:]
        keep = min(CONTEXT, W)
        self._end = end + k
        self._t = keep
        self._carry = prefix[:, W - keep]
        self.frames_pushed += k
        self.hops_emitted += hi - lo
        return out

    def flush(self):
        """The hops still held back (at most LATENCY), ending the way one call over all pushed frames ends
        (f0 and control rows held after the last frame, core.py:17 / vocoder.py:470-474).  Resets the stream."""
        if self._last_tail is None:
            raise RuntimeError('flush() on an empty stream')
        out = self._last_tail
        self.reset()
        return out


class StreamingCombSubFast:
    """Module-level stream around a `ddsp_b200.vocoder.CombSubFast` (or any module with its `unit2ctrl`,
    `window`, `_hop`, `_sr`):

        audio, phase = s.push(units, f0, volume, spk_id, n_context=c)

    `units` (B,c+k,n_unit), `f0`, `volume` (B,c+k,1): the k new frames preceded by c frames that were
    pushed before and are repeated only as left context for the (non-causal) control network -- what
    `extra_time` is for in gui.py:373-388.  The network runs over all c+k frames, the synthesizer only
    continues with the rows of the k new ones.  c may be 0 and must not exceed `history` (frames of
    phase kept) or the frames pushed so far."""

    def __init__(self, model, history=256, seed=None, initial_phase=None):
        self.model = model
        self.history = int(history)
        self.stream = CombSubFastStream(model._hop, model._sr, window=model.window,
                                        seed=model._next_seed() if seed is None else seed,
                                        initial_phase=initial_phase)
        self._phase_hist = None

    def reset(self):
        self.stream.reset()
        self._phase_hist = None

    @torch.no_grad()
    def push(self, units_frames, f0_frames, volume_frames, spk_id=None, spk_mix_dict=None, n_context=0, noise_u=None):
        c = int(n_context)
        f0_frames = core.as_f32(f0_frames)
        have = 0 if self._phase_hist is None else self._phase_hist.shape[1]
        if c < 0 or c > have or c >= f0_frames.shape[1]:
            raise ValueError(f'n_context={c}: only {have} frames of history are kept and at least one frame must be new')
        f0_2d = core._f0_2d(f0_frames)
        phase_new = self.stream.begin(f0_2d[:, c:])
        phase_win = phase_new if c == 0 else torch.cat((self._phase_hist[:, have - c:], phase_new), dim=1)
        ctrls = self.model.unit2ctrl(units_frames, f0_frames, phase_win, volume_frames, spk_id, spk_mix_dict=spk_mix_dict)
        audio = self.stream.finish(ctrls['harmonic_magnitude'][:, c:], ctrls['harmonic_phase'][:, c:],
                                   ctrls['noise_magnitude'][:, c:], noise_u=noise_u)
        hist = phase_new if self._phase_hist is None else torch.cat((self._phase_hist, phase_new), dim=1)
        self._phase_hist = hist[:, -self.history:]
        return audio, phase_new.unsqueeze(-1)

    def flush(self):
        out = self.stream.flush()
        self._phase_hist = None
        return out


class FilterModelStream:
    """Stage-level stream for the `frequency_filter` models (`Sins`, vocoder.py:381-423; `CombSub` old, :504-550):

        audio = stream.push(c0, c1, c2, f0_new)        # (B, 512*n) finished samples of (signal, harmonic, noise)
        tail  = stream.flush()                          # the hops still held back; resets the stream

    `model` = 'sins': c0, c1, c2 = amplitudes, group_delay, noise_magnitude; 'combsub': group_delay,
    harmonic_magnitude, noise_magnitude -- rows of the k new frames, (B,k,K) each.

    The LTV-FIR output y[t] = sum_s x[s] h_s[t + L/2 - s] (core.py:185-336) needs input samples up to L/2 on either side,
    and a sample of hop h needs frame h+1 (interpolated f0 and impulse response).  With the block's window starting `LEFT`
    frames before the first hop still owed and the newest `RIGHT` hops held back, every hop handed out equals what ONE call
    over all frames pushed so far produces (to the last ulp or two: the overlap-add order follows the run partition of the
    call):  CombSub chains a 510-tap and a 1022-tap filter -> LEFT = 2, RIGHT = 3;  Sins has 510-tap filters only ->
    LEFT = 1, RIGHT = 2.  The phase is carried as the fp64 prefix (`ddsp_b200_phase_stream`), the in-kernel noise by hop
    index (`hop_offset`), so re-synthesised hops repeat exactly."""

    GEOMETRY = {'sins': (1, 2), 'combsub': (2, 3)}

    def __init__(self, model, block_size=512, sampling_rate=44100, seed=0, initial_phase=None):
        if model not in self.GEOMETRY:
            raise ValueError("model must be 'sins' or 'combsub'")
        self.model = model
        self.left, self.right = self.GEOMETRY[model]
        self.hop = int(block_size)
        self.sr = float(sampling_rate)
        self.seed = int(seed)
        self.initial_phase = initial_phase
        self._utterance = -1
        self.reset()

    def reset(self):
        self._utterance += 1
        self._seed_now = (self.seed + 0x9E3779B97F4A7C15 * self._utterance) & ((1 << 62) - 1)
        self.frames_pushed = 0
        self.hops_emitted = 0
        self._f0 = None                 # (B,w) frames kept: the next window's left part
        self._rows = None               # list of three (B,w,K)
        self._noise = None              # (B,w*hop) injected noise of the kept frames (parity mode)
        self._first = 0                 # stream index of the first kept frame
        self._carry = None              # (B,) fp64: prefix at the first kept frame
        self._held = None               # the hops held back, ended the hold-last way (for flush())

    def _synth(self, rows, f0_win, noise_win, first):
        B, W = f0_win.shape
        if self.model == 'sins':
            _, prefix, phase = core.phase_stage_stream(f0_win, self.hop, self.sr, carry=self._carry,
                                                       initial_phase=self.initial_phase if self._carry is None else None,
                                                       full_rate=True)
            out = core.sins_stage(rows[0], rows[1], rows[2], f0_win, phase, self.hop, self.sr, noise_u=noise_win,
                                  seed=self._seed_now, hop_offset=first)
        else:
            _, prefix = core.phase_stage_stream(f0_win, self.hop, self.sr, carry=self._carry,
                                                initial_phase=self.initial_phase if self._carry is None else None)
            out = core.combsub_stage(rows[0], rows[1], rows[2], f0_win, prefix, self.hop, self.sr, noise_u=noise_win,
                                     seed=self._seed_now, hop_offset=first)
        return out, prefix

    def push(self, c0, c1, c2, f0_new, noise_u=None):
        """k new frames in, the newly finished hops out as (signal, harmonic, noise), each (B, 512*n) -- n = k in the
        steady state, k - RIGHT on the first push (possibly 0 samples)."""
        f0_new = core._f0_2d(core.as_f32(f0_new))
        B, k = f0_new.shape
        if k < 1:
            raise ValueError('push() needs at least one new frame')
        new_rows = [core.as_f32(r) for r in (c0, c1, c2)]
        for r in new_rows:
            if r.shape[0] != B or r.shape[1] != k:
                raise ValueError('control rows must be (B, k, K) for the k new frames')
        if self._f0 is not None and self._f0.shape[0] != B:
            raise ValueError('the number of clips must not change inside a stream (reset() first)')
        if (noise_u is None) != (self._noise is None) and self._f0 is not None:
            raise ValueError('either every block of a stream injects noise_u or none does')
        if noise_u is not None:
            noise_u = core._need_cuda_f32(noise_u, 'noise_u')
            if tuple(noise_u.shape) != (B, k * self.hop):
                raise ValueError('noise_u must be (B, k*block_size)')
        if self._f0 is None:
            f0_win, rows, noise_win = f0_new.contiguous(), [r.contiguous() for r in new_rows], noise_u
        else:
            f0_win = torch.cat((self._f0, f0_new), dim=1)
            rows = [torch.cat((a, b), dim=1) for a, b in zip(self._rows, new_rows)]
            noise_win = None if noise_u is None else torch.cat((self._noise, noise_u), dim=1)
        first = self._first
        W = f0_win.shape[1]
        (signal, harmonic, noise), prefix = self._synth(rows, f0_win, noise_win, first)
        n_total = first + W                                  # frames pushed so far
        lo = self.hops_emitted - first                       # window hop of the first hop still owed
        hi = max(lo, W - self.right)
        hop = self.hop
        out = tuple(t[:, lo * hop:hi * hop] for t in (signal, harmonic, noise))
        self._held = tuple(t[:, hi * hop:] for t in (signal, harmonic, noise))
        self.hops_emitted += hi - lo
        self.frames_pushed = n_total
        # keep the frames from LEFT before the first hop still owed
        keep_from = max(0, self.hops_emitted - self.left - first)
        self._f0 = f0_win[:, keep_from:].contiguous()
        self._rows = [r[:, keep_from:].contiguous() for r in rows]
        self._noise = None if noise_win is None else noise_win[:, keep_from * hop:].contiguous()
        self._carry = prefix[:, keep_from].clone()
        self._first = first + keep_from
        return out

    def flush(self):
        """The hops still held back, ending the way one call over all pushed frames ends.  Resets the stream."""
        if self._held is None:
            raise RuntimeError('flush() on an empty stream')
        out = self._held
        self.reset()
        return out
