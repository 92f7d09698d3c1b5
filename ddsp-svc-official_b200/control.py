"""Control network (first step): a plain-PyTorch `Unit2Control` that is
state_dict-compatible with the reference's `ddsp/unit2control.py` + `ddsp/pcmer.py`, so that the
drop-in synthesizer modules are usable (and reference checkpoints load with `strict=True`) without
the reference repo and its un-vendored dependencies (`extorch`, `pytorch-fast-transformers`).

This is NOT part of the hand-written synthesizer path: the GEMMs and convolutions are stock PyTorch
ops (cuBLAS / cuDNN), exactly like the reference's own control network.  Under `torch.no_grad()` on
CUDA the memory-bound chains between the GEMMs run as two fused kernels of `csrc/control.cuh`
(FAVOR+ feature map; GLU -> depthwise conv -> SiLU in channels-last layout); with autograd enabled or
on CPU the plain ops below run.  The causal configuration (`c: true`; no shipped config sets it) runs on the plain
ops only: causal convolutions and chunked causal linear attention.

Structure (the module / parameter names are dictated by the checkpoint layout):
    unit_prenet : T - Conv1d(k3) - GroupNorm(4) - LeakyReLU - Conv1d(k3) - T        unit2control.py:38-45
    f0/phase/volume_embed : Linear(1, 256);  spk_embed : Embedding                  :50-53
    dec_post    : PCmer(3 layers, 8 heads) - LayerNorm - weight_norm(Linear)         :56-62
    PCmer layer : x + Attn(LN(x));  x + ConvModule(x)                                pcmer.py:20-37
    Attn        : Performer (FAVOR+) softmax-kernel linear attention, 266 features   pcmer.py:69-78,124-160,191-251
    ConvModule  : LN - T - Conv1x1(256->1024) - GLU - depthwise Conv(k31) - SiLU - Conv1x1(512->256) - T   pcmer.py:41-63
"""
import math
import os

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.utils import weight_norm

_DIM = 256
_HEADS = 8
_DIM_HEAD = 64
_FUSED_PROJECTION_MIN_ROWS = 1 << 16      # (batch x frames x heads) above which the fused projection kernel wins
_ATTENTION_KERNEL_MAX_FRAMES = 256        # (batch x frames) up to which the fused attention kernels win (measured: 130 yes, 431 no)


class _Swap(nn.Module):
    """(B, T, C) <-> (B, C, T); parameter-free, occupies a slot in the Sequential like the reference's Transpose."""

    def __init__(self, a, b):
        super().__init__()
        self.a, self.b = a, b

    def forward(self, x):
        return x.transpose(self.a, self.b)


def _orthogonal_gaussian_features(n_rows, n_cols):
    """Random-feature matrix for FAVOR+: stacked Q factors of Gaussian blocks, rows rescaled to the
    norms of Gaussian vectors (pcmer.py:80-120 with scaling=0).  Only used for fresh initialisation;
    checkpoints carry their own `projection_matrix` buffer."""
    blocks, left = [], n_rows
    while left > 0:
        q, _ = torch.linalg.qr(torch.randn(n_cols, n_cols), mode='reduced')
        blocks.append(q.t()[:min(left, n_cols)])
        left -= n_cols
    mat = torch.cat(blocks)
    return torch.randn(n_rows, n_cols).norm(dim=1).unsqueeze(1) * mat


def _softmax_features(x, proj, is_query, eps=1e-4):
    """Positive random features approximating the softmax kernel (pcmer.py:124-160).  x: (B,H,N,D)."""
    d = x.shape[-1]
    scale = d ** -0.25
    ratio = proj.shape[0] ** -0.5
    dash = torch.einsum('bhnd,jd->bhnj', scale * x, proj.to(x.dtype))
    diag = (x * x).sum(dim=-1, keepdim=True) * (0.5 * scale * scale)
    if is_query:
        return ratio * (torch.exp(dash - diag - dash.amax(dim=-1, keepdim=True)) + eps)
    return ratio * torch.exp(dash - diag + eps)


def _split_cached(owner, slot, tensors, build=None):
    """(hi, lo) TF32 split of a weight (or of `build(*tensors)`), cached on `owner` until one of `tensors` changes
    (in-place update, load_state_dict, .to())."""
    from . import core
    key = tuple((t._version, t.data_ptr()) for t in tensors)
    cache = owner.__dict__.setdefault('_tc_cache', {})
    hit = cache.get(slot)
    if hit is None or hit[0] != key:
        w = build(*[t.detach() for t in tensors]) if build is not None else tensors[0].detach()
        w = w.reshape(w.shape[0], -1).contiguous()
        # DDSP_B200_NO_PRESPLIT=1 (experiments): leave the split to the kernel's splitter warps (half the weight bytes per tile)
        hit = (key, w, None) if os.environ.get('DDSP_B200_NO_PRESPLIT') == '1' else (key,) + tuple(core.split_tf32(w))
        cache[slot] = hit
    return hit[1], hit[2]


def _tc_path(x):
    """Calls larger than a GUI block run every GEMM-shaped step of the network on the tensor cores (csrc/gemm_attn.cuh)."""
    return _fused_ok(x) and x.shape[0] * x.shape[1] > _ATTENTION_KERNEL_MAX_FRAMES and x.shape[-1] == _DIM


def _fused_ok(x):
    """The fused CUDA stages (csrc/control.cuh) are inference-only and fp32-only; everything else --
    CPU tensors, autograd, other dtypes -- takes the plain PyTorch ops below."""
    return x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled()


class _CausalConv1d(nn.Conv1d):
    """`extorch.Conv1dEx(..., padding="same", causal=True)` (unit2control.py:40,43, pcmer.py:54): the output frame n sees the
    input frames n-k+1 .. n only (left padding by kernel_size - 1).  extorch is an un-vendored, unpinned dependency that
    is not installed here: this follows its documented meaning (parity unpinned for c=True)."""

    def __init__(self, c_in, c_out, kernel_size, groups=1):
        super().__init__(c_in, c_out, kernel_size, padding=0, groups=groups)

    def forward(self, x):
        return super().forward(F.pad(x, (self.kernel_size[0] - 1, 0)))


def _causal_attend(q, k, v, eps=1e-6, chunk=128):
    """Causal linear attention (pcmer.py:141-160, `fast_transformers.causal_product.CausalDotProduct`):
    out_n = sum_{m<=n} (q_n . k_m) v_m / (q_n . (sum_{m<=n} k_m + eps)), evaluated chunk-wise: inside a chunk the masked
    score matrix, across chunks the running (features x dim) state -- O(N) memory, differentiable plain ops."""
    b, h, n, j = q.shape
    k_cum = k.cumsum(dim=-2) + eps
    d_inv = 1.0 / torch.einsum('bhnj,bhnj->bhn', q, k_cum)
    state = q.new_zeros(b, h, j, v.shape[-1])
    outs = []
    for s0 in range(0, n, chunk):
        qc, kc, vc = q[:, :, s0:s0 + chunk], k[:, :, s0:s0 + chunk], v[:, :, s0:s0 + chunk]
        scores = torch.einsum('bhnj,bhmj->bhnm', qc, kc).tril_()
        outs.append(torch.einsum('bhnm,bhme->bhne', scores, vc) + torch.einsum('bhnj,bhje->bhne', qc, state))
        state = state + torch.einsum('bhmj,bhme->bhje', kc, vc)
    return torch.cat(outs, dim=2) * d_inv.unsqueeze(-1)


class _FastAttention(nn.Module):
    def __init__(self, dim_head, causal=False):
        super().__init__()
        n_features = int(dim_head * math.log(dim_head))
        self.causal = causal
        self.register_buffer('projection_matrix', _orthogonal_gaussian_features(n_features, dim_head))

    def forward(self, q, k, v):
        q = _softmax_features(q, self.projection_matrix, True)
        k = _softmax_features(k, self.projection_matrix, False)
        return _causal_attend(q, k, v) if self.causal else self.attend(q, k, v)

    @staticmethod
    def attend(q, k, v):
        # non-causal linear attention (pcmer.py:69-78)
        k_sum = k.sum(dim=-2)
        d_inv = 1.0 / (torch.einsum('bhnj,bhj->bhn', q, k_sum) + 1e-8)
        context = torch.einsum('bhnj,bhne->bhje', k, v)
        return torch.einsum('bhje,bhnj,bhn->bhne', context, q, d_inv)


class _SelfAttention(nn.Module):
    def __init__(self, dim, heads, causal=False):
        super().__init__()
        inner = _DIM_HEAD * heads
        self.heads = heads
        self.causal = causal
        self.fast_attention = _FastAttention(_DIM_HEAD, causal)
        self.to_q = nn.Linear(dim, inner)
        self.to_k = nn.Linear(dim, inner)
        self.to_v = nn.Linear(dim, inner)
        self.to_out = nn.Linear(inner, dim)

    def _merged_qkv_weight(self):
        """[W_q; W_k; W_v] for the one-GEMM projection of the streaming path, rebuilt when a weight changes
        (in-place update, load_state_dict, .to())."""
        ws = (self.to_q.weight, self.to_k.weight, self.to_v.weight)
        key = tuple((w._version, w.data_ptr()) for w in ws)
        if getattr(self, '_qkv_key', None) != key:
            self._qkv_cache = torch.cat([w.detach() for w in ws])
            self._qkv_key = key
        return self._qkv_cache

    def forward_tc(self, xn, residual, ln):
        """Tensor-core path: xn = LayerNorm(x) -> attention -> `to_out` with bias + residual in the epilogue.  Returns
        (residual + attention, LayerNorm_ln(residual + attention)); writes the first in place over `residual`."""
        from . import core
        fa = self.fast_attention
        w_hi, w_lo = _split_cached(self, 'qkv', (self.to_q.weight, self.to_k.weight, self.to_v.weight), lambda a, b, c: torch.cat([a, b, c]))
        cache = self.__dict__['_tc_cache']
        bkey = tuple((t._version, t.data_ptr()) for t in (self.to_q.bias, self.to_k.bias, self.to_v.bias, fa.projection_matrix))
        if cache.get('aux', (None,))[0] != bkey:
            cache['aux'] = (bkey, torch.cat([self.to_q.bias.detach(), self.to_k.bias.detach(), self.to_v.bias.detach()]).contiguous(),
                            (_DIM_HEAD ** -0.25 * fa.projection_matrix.detach()).contiguous())
        _, b_qkv, proj_scaled = cache['aux']
        att = core.favor_attention(xn, w_hi, w_lo, b_qkv, proj_scaled, self.heads)
        o_hi, o_lo = _split_cached(self, 'out', (self.to_out.weight,))
        return core.linear_ex(att, o_hi, self.to_out.bias, residual=residual, out=residual, weight_lo=o_lo,
                              ln=(ln.weight, ln.bias, ln.eps))

    def forward(self, x, residual=None):
        """`residual` (fused path only): returns residual + attention(x), with the output bias and the
        residual folded into the output GEMM (addmm, beta = 1) instead of two more elementwise passes."""
        b, n, _ = x.shape
        split = lambda t: t.view(b, n, self.heads, _DIM_HEAD).transpose(1, 2)      # noqa: E731
        if _fused_ok(x) and not self.causal:
            from . import core
            proj = self.fast_attention.projection_matrix
            if b * n * self.heads >= _FUSED_PROJECTION_MIN_ROWS:
                # large batches: the three projections run on the tensor cores (3xTF32 tcgen05 GEMM, csrc/gemm_tc.cuh);
                # the random-feature projection is fused into the feature kernel, the (B,N,H,266) products stay on
                # chip; the q/k Linear biases are added there too
                q = core.performer_project_features(core.linear(x, self.to_q.weight), proj, self.heads, True,
                                                    x_bias=self.to_q.bias)
                k = core.performer_project_features(core.linear(x, self.to_k.weight), proj, self.heads, False,
                                                    x_bias=self.to_k.bias)
                out = self.fast_attention.attend(q, k, split(core.linear(x, self.to_v.weight, self.to_v.bias)))
                out = out.transpose(1, 2).reshape(b * n, self.heads * _DIM_HEAD)
            elif b * n > _ATTENTION_KERNEL_MAX_FRAMES:
                # mid-sized calls: tensor-core GEMMs + the one-pass feature kernel
                q = core.linear(x, self.to_q.weight, self.to_q.bias)
                k = core.linear(x, self.to_k.weight, self.to_k.bias)
                scale = _DIM_HEAD ** -0.25
                q, k = [core.performer_features(torch.matmul((scale * t).view(-1, _DIM_HEAD), proj.t()), t, self.heads, is_q)
                        for t, is_q in ((q, True), (k, False))]
                out = self.fast_attention.attend(q, k, split(core.linear(x, self.to_v.weight, self.to_v.bias)))
                out = out.transpose(1, 2).reshape(b * n, self.heads * _DIM_HEAD)
            else:
                # streaming blocks: the whole attention after the three (bias-free) projection GEMMs is one
                # kernel (<= 16 frames) or three tiled ones -- launch latencies, not bytes, bound a GUI block
                q, k, v = F.linear(x, self._merged_qkv_weight()).chunk(3, dim=-1)      # one GEMM, three strided slices
                out = core.performer_attention(q, k, v, proj, self.heads, self.to_q.bias, self.to_k.bias,
                                               self.to_v.bias).view(b * n, -1)
            if b * n > _ATTENTION_KERNEL_MAX_FRAMES:
                # bias and residual are added in the epilogue of the tensor-core GEMM
                return core.linear(out, self.to_out.weight, self.to_out.bias, residual=residual).view(b, n, -1)
            if residual is None:
                return self.to_out(out).view(b, n, -1)
            # residual + bias lands in a fresh buffer that the GEMM then accumulates into in place (no copy of C)
            return (residual + self.to_out.bias).reshape(b * n, -1).addmm_(out, self.to_out.weight.t()).view(b, n, -1)
        out = self.fast_attention(split(self.to_q(x)), split(self.to_k(x)), split(self.to_v(x)))
        return self.to_out(out.transpose(1, 2).reshape(b, n, self.heads * _DIM_HEAD))


class _ConvModule(nn.Module):
    def __init__(self, dim, expansion=2, kernel_size=31, causal=False):
        super().__init__()
        inner = dim * expansion
        self.causal = causal
        self.net = nn.Sequential(
            nn.LayerNorm(dim),
            _Swap(1, 2),
            nn.Conv1d(dim, inner * 2, 1),
            nn.GLU(dim=1),
            _CausalConv1d(inner, inner, kernel_size, groups=inner) if causal else
            nn.Conv1d(inner, inner, kernel_size, padding='same', groups=inner),
            nn.SiLU(),
            nn.Conv1d(inner, dim, 1),
            _Swap(1, 2),
            nn.Dropout(0.0),
        )

    def forward_tc(self, xn, residual, next_ln):
        """Tensor-core path: xn = this module's LayerNorm(x) (already produced by the previous GEMM's epilogue).
        Returns (residual + module, next_ln(residual + module)) -- or only the first when `next_ln` is None."""
        from . import core
        _, _, pw1, _, dw, _, pw2, _, _ = self.net
        # pointwise conv + GLU in one GEMM (weight rows interleaved per column tile), then depthwise conv + SiLU
        w1_hi, w1_lo = _split_cached(self, 'pw1', (pw1.weight,), lambda w: core.glu_interleave(w))
        cache = self.__dict__['_tc_cache']
        bkey = (pw1.bias._version, pw1.bias.data_ptr())
        if cache.get('pw1_bias', (None,))[0] != bkey:
            cache['pw1_bias'] = (bkey, core.glu_interleave(pw1.weight.detach(), pw1.bias.detach())[1])
        s = core.dwconv_silu(core.linear_glu(xn, w1_hi, cache['pw1_bias'][1], weight_lo=w1_lo), dw.weight, dw.bias)
        w2_hi, w2_lo = _split_cached(self, 'pw2', (pw2.weight,))
        ln = None if next_ln is None else (next_ln.weight, next_ln.bias, next_ln.eps)
        return core.linear_ex(s, w2_hi, pw2.bias, residual=residual, out=residual, weight_lo=w2_lo, ln=ln)

    def forward(self, x, residual=None):
        """`residual` (fused path only): returns residual + module(x) with bias and residual folded into the
        last GEMM."""
        if _fused_ok(x) and not self.causal:
            from . import core
            ln, _, pw1, _, dw, _, pw2, _, _ = self.net
            big = x.shape[0] * x.shape[1] > _ATTENTION_KERNEL_MAX_FRAMES
            lin = core.linear if big else F.linear                                 # tensor cores for anything but streaming blocks
            u = lin(ln(x), pw1.weight.squeeze(-1))                                 # channels last: no transposes
            s = core.glu_dwconv_silu(u, dw.weight, dw.bias, u_bias=pw1.bias)
            if big:
                return core.linear(s, pw2.weight.squeeze(-1), pw2.bias, residual=residual)
            if residual is None:
                return F.linear(s, pw2.weight.squeeze(-1), pw2.bias)
            b, n, _ = x.shape
            return (residual + pw2.bias).reshape(b * n, -1).addmm_(s.view(b * n, -1), pw2.weight.squeeze(-1).t()).view(b, n, -1)
        return self.net(x)


class _EncoderLayer(nn.Module):
    def __init__(self, heads, dim, causal=False):
        super().__init__()
        self.causal = causal
        self.norm = nn.LayerNorm(dim)
        self.attn = _SelfAttention(dim, heads, causal)
        self.local_mixer = _ConvModule(dim, causal=causal)

    def forward(self, x):
        if _fused_ok(x) and not self.causal:
            x = self.attn(self.norm(x), residual=x)
            return self.local_mixer(x, residual=x)
        x = x + self.attn(self.norm(x))
        return x + self.local_mixer(x)


class PCmer(nn.Module):
    def __init__(self, num_layers, num_heads, dim_model, causal=False):
        super().__init__()
        self.causal = causal
        self.net = nn.Sequential(*[_EncoderLayer(num_heads, dim_model, causal) for _ in range(num_layers)])

    def forward(self, x):
        return self.net(x)

    def forward_tc(self, x, post_ln, xn=None):
        """Tensor-core path over all layers; every LayerNorm but the first is produced by the epilogue of the GEMM
        that finishes its input (pcmer.py:25-37); the first one arrives as `xn` when the embedding kernel made it.
        Returns post_ln(PCmer(x)); x is overwritten (residual stream)."""
        layers = list(self.net)
        if xn is None:
            xn = F.layer_norm(x, (x.shape[-1],), layers[0].norm.weight, layers[0].norm.bias, layers[0].norm.eps)
        for i, layer in enumerate(layers):
            x, xn = layer.attn.forward_tc(xn, x, layer.local_mixer.net[0])
            nxt = layers[i + 1].norm if i + 1 < len(layers) else post_ln
            x, xn = layer.local_mixer.forward_tc(xn, x, nxt)
        return xn


class Unit2Control(nn.Module):
    """`Unit2Control(n_unit, n_spk, output_splits, c=False)` -- same call signature and return value
    (dict of strided `torch.split` views of one (B, Frame, sum K) tensor) as unit2control.py:23-101."""

    def __init__(self, ndim_feat_i, n_spk, output_splits, c=False):
        super().__init__()
        self.causal = bool(c)
        conv = (lambda i, o: _CausalConv1d(i, o, 3)) if c else (lambda i, o: nn.Conv1d(i, o, 3, padding='same'))
        self.unit_prenet = nn.Sequential(
            _Swap(1, 2),
            conv(ndim_feat_i, _DIM),
            nn.GroupNorm(4, _DIM),
            nn.LeakyReLU(),
            conv(_DIM, _DIM),
            _Swap(1, 2),
        )
        self.f0_embed = nn.Linear(1, _DIM)
        self.phase_embed = nn.Linear(1, _DIM)
        self.volume_embed = nn.Linear(1, _DIM)
        self.spk_embed = nn.Embedding(n_spk, _DIM)
        n_out = sum(output_splits.values())
        self.dec_post = nn.Sequential(PCmer(3, _HEADS, _DIM, self.causal), nn.LayerNorm(_DIM), weight_norm(nn.Linear(_DIM, n_out)))
        self.output_splits = dict(output_splits)

    # Opt-in: run the fp32 GEMMs of the network on the TF32 tensor cores (the reference's cuDNN
    # convolutions already do by PyTorch default; its Linear layers do not).  Off by default so that the
    # control rows match the reference's fp32 Linear arithmetic.
    matmul_tf32 = False

    def forward(self, units, f0, phase, volume, spk_id, spk_mix_dict=None):
        if self.matmul_tf32 and units.is_cuda:
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = True
            try:
                return self._forward(units, f0, phase, volume, spk_id, spk_mix_dict)
            finally:
                torch.backends.cuda.matmul.allow_tf32 = prev
        return self._forward(units, f0, phase, volume, spk_id, spk_mix_dict)

    def _prenet_tc(self, units):
        """unit_prenet on the tensor cores, channels last (no transposes): each Conv1d(k=3) is one 3xTF32 GEMM over the
        zero-padded frames read as overlapping rows (core.conv3_frames); GroupNorm + LeakyReLU in one in-place pass.
        Returns a (B, N, 256) view (row stride 256, clip stride (N+2)*256)."""
        from . import core
        _, conv1, gn, act, conv2, _ = self.unit_prenet
        B, N, _ = units.shape
        w1_hi, w1_lo = _split_cached(self, 'pre1', (conv1.weight,), core.conv3_weight)
        w2_hi, w2_lo = _split_cached(self, 'pre2', (conv2.weight,), core.conv3_weight)
        xp = core.pad_frames(units)
        hp = torch.empty((B, N + 2, _DIM), dtype=torch.float32, device=units.device)
        rows = hp.view(-1, _DIM)
        core.conv3_frames(xp, w1_hi, conv1.bias, rows[1:rows.shape[0] - 1], weight_lo=w1_lo)       # frame n -> padded frame n + 1
        core.groupnorm_leaky_(hp, gn.weight, gn.bias, gn.eps, gn.num_groups, act.negative_slope)
        yp = torch.empty((B, N + 2, _DIM), dtype=torch.float32, device=units.device)
        core.conv3_frames(hp, w2_hi, conv2.bias, yp.view(-1, _DIM)[:rows.shape[0] - 2], weight_lo=w2_lo)
        return yp[:, :N]

    def _forward(self, units, f0, phase, volume, spk_id, spk_mix_dict=None):
        tc_pre = (_fused_ok(units) and not self.causal and units.dim() == 3 and units.shape[0] * units.shape[1] > _ATTENTION_KERNEL_MAX_FRAMES
                  and units.shape[-1] % 32 == 0 and f0.dtype == torch.float32 and phase.dtype == torch.float32
                  and volume.dtype == torch.float32 and os.environ.get('DDSP_B200_STOCK_PRENET') != '1')
        x = self._prenet_tc(units) if tc_pre else self.unit_prenet(units)
        if _fused_ok(x) and f0.dtype == torch.float32 and phase.dtype == torch.float32 and volume.dtype == torch.float32:
            from . import core
            if spk_mix_dict is not None:                 # weighted mix of speaker embeddings (unit2control.py:89-93)
                spk = sum(v * self.spk_embed.weight[int(k) - 1] for k, v in spk_mix_dict.items()).reshape(1, -1)
            else:
                spk = self.spk_embed(spk_id - 1)
            xn = None
            if tc_pre:                                   # embedding sum + the first LayerNorm of PCmer in one pass
                x, xn = core.embed_sum_ln(x, f0, phase, volume, self.f0_embed, self.phase_embed, self.volume_embed, spk,
                                          self.dec_post[0].net[0].norm)
            else:
                x = core.embed_sum(x, f0, phase, volume, self.f0_embed, self.phase_embed, self.volume_embed, spk)
            names, sizes = list(self.output_splits), list(self.output_splits.values())
            if _tc_path(x) and not self.causal:
                # output projection on the tensor cores into a buffer whose row stride is padded to a multiple of 4
                # floats (128-bit stores); the synthesizer consumes the strided views as they are
                pcmer, norm, proj = self.dec_post
                n_out = sum(sizes)
                buf = torch.empty(x.shape[:-1] + ((n_out + 3) // 4 * 4,), dtype=torch.float32, device=x.device)
                # weight_norm keeps (weight_g, weight_v) and rebuilds `weight` only inside Module.__call__
                if hasattr(proj, 'weight_g'):
                    w_hi, w_lo = _split_cached(self, 'proj', (proj.weight_v, proj.weight_g), lambda v, g: torch._weight_norm(v, g, 0))
                else:
                    w_hi, w_lo = _split_cached(self, 'proj', (proj.weight,))
                if not x.is_contiguous():
                    x = x.contiguous()
                e = core.linear_ex(pcmer.forward_tc(x, norm, xn), w_hi, proj.bias, out=buf[..., :n_out], weight_lo=w_lo)
            else:
                e = self.dec_post(x)
            return dict(zip(names, torch.split(e, sizes, dim=-1)))
        x = x + self.f0_embed((1 + f0 / 700).log()) + self.phase_embed(phase.unsqueeze(-1) / math.pi) \
            + self.volume_embed(volume.unsqueeze(-1))
        if spk_mix_dict is not None:                     # weighted mix of speaker embeddings (unit2control.py:89-93)
            for k, v in spk_mix_dict.items():
                idx = torch.tensor([[int(k) - 1]], dtype=torch.long, device=units.device)
                x = x + v * self.spk_embed(idx)
        else:
            x = x + self.spk_embed(spk_id - 1)
        e = self.dec_post(x)
        names, sizes = list(self.output_splits), list(self.output_splits.values())
        return dict(zip(names, torch.split(e, sizes, dim=-1)))
