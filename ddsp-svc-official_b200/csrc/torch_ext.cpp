// torch_ext.cpp -- the PyTorch-extension host of the synthesizer path: TORCH_LIBRARY(ddsp_b200, ...) operators that
// validate dtype / device / shape, allocate the outputs, pick up torch's current CUDA stream and call the thin C ABI
// (include/ddsp_b200.h).  No kernel lives here; lib/ddsp_b200_torch.so links lib/libddsp_b200.so.
//
// They stand where the reference's modules call into ddsp/core.py + torch ops:
//   ddsp_b200::phase         vocoder.py:391-393 / 449-451 / 515-517   (upsample + fo_to_rot, core.py:7-51)
//   ddsp_b200::combsubfast   vocoder.py:455-490
//   ddsp_b200::combsub       vocoder.py:521-548
//   ddsp_b200::sins          vocoder.py:397-421
// Shape / batch mismatches raise ValueError (TORCH_CHECK_VALUE), as core.py:212-213 does; a CPU tensor raises
// (there is no CPU path).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include <tuple>

#include "../../include/ddsp_b200.h"

namespace {

using torch::Tensor;

void check_status(int rc) {
    if (rc == DDSP_B200_OK) return;
    std::string msg = ddsp_b200_strerror(rc);
    if (rc == DDSP_B200_ERR_CUDA) msg += " [cudaError " + std::to_string(ddsp_b200_last_cuda_error()) + "]";
    TORCH_CHECK_VALUE(rc != DDSP_B200_ERR_BATCH_MISMATCH, msg);
    TORCH_CHECK(false, "ddsp_b200: ", msg);
}

const Tensor& need_f32(const Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (ddsp_b200 has no CPU path)");
    TORCH_CHECK_TYPE(t.scalar_type() == torch::kFloat32, name, " must be float32");
    return t;
}

// (B, Frame, 1) or (B, Frame) -> strided (B, Frame) view
Tensor f0_2d(const Tensor& f0) {
    need_f32(f0, "f0_frames");
    if (f0.dim() == 3) {
        TORCH_CHECK_VALUE(f0.size(2) == 1, "f0_frames must be (B, Frame, 1)");
        return f0.select(2, 0);
    }
    TORCH_CHECK_VALUE(f0.dim() == 2, "f0_frames must be (B, Frame, 1) or (B, Frame)");
    return f0;
}

const float* opt_ptr(const c10::optional<Tensor>& t) { return t.has_value() ? t->data_ptr<float>() : nullptr; }

// The three control tensors arrive as torch.split views of one (B,F,sumK) tensor (unit2control.py:10-20): passed
// through untouched when they share (batch, row) strides and have unit inner stride, gathered otherwise.
void common_views(Tensor& a, Tensor& b, Tensor& c) {
    bool ok = true;
    for (const Tensor* t : {&a, &b, &c})
        ok = ok && t->dim() == 3 && t->stride(2) == 1 && t->stride(0) == a.stride(0) && t->stride(1) == a.stride(1);
    if (ok) return;
    const int64_t ka = a.size(2), kb = b.size(2), kc = c.size(2);
    Tensor packed = torch::cat({a.contiguous(), b.contiguous(), c.contiguous()}, -1);
    a = packed.narrow(2, 0, ka);
    b = packed.narrow(2, ka, kb);
    c = packed.narrow(2, ka + kb, kc);
}

Tensor checked_noise(const c10::optional<Tensor>& noise_u, int64_t B, int64_t T) {
    if (!noise_u.has_value()) return Tensor();
    Tensor u = need_f32(*noise_u, "noise_u").contiguous();
    TORCH_CHECK_VALUE(u.dim() == 2 && u.size(0) == B && u.size(1) == T, "noise_u must be (B, T)");
    return u;
}

std::tuple<Tensor, Tensor, Tensor> phase(const Tensor& f0_frames, int64_t hop, double sr,
                                         const c10::optional<Tensor>& initial_phase, bool infer, bool full_rate,
                                         const c10::optional<Tensor>& carry) {
    Tensor f0 = f0_2d(f0_frames);
    const int64_t B = f0.size(0), F = f0.size(1);
    c10::cuda::CUDAGuard guard(f0.device());
    auto opts = f0.options();
    Tensor phase_frames = torch::empty({B, F}, opts);
    Tensor prefix = torch::empty({B, F}, opts.dtype(torch::kFloat64));
    Tensor phase_full = full_rate ? torch::empty({B, F * hop}, opts) : Tensor();
    Tensor ip;
    if (initial_phase.has_value()) {
        ip = initial_phase->to(f0.device(), torch::kFloat32).reshape({-1}).contiguous();
        TORCH_CHECK_VALUE(ip.numel() == B, "initial_phase must have one entry per clip");
    }
    void* st = at::cuda::getCurrentCUDAStream().stream();
    if (carry.has_value()) {
        const Tensor& c = *carry;
        TORCH_CHECK_VALUE(c.is_cuda() && c.scalar_type() == torch::kFloat64 && c.dim() == 1 && c.numel() == B,
                          "carry must be a CUDA float64 tensor with one entry per clip");
        TORCH_CHECK_VALUE(!full_rate, "a streamed block has no full-rate phase output");
        check_status(ddsp_b200_phase_stream(f0.data_ptr<float>(), f0.stride(0), f0.stride(1), (int)B, (int)F, (int)hop, sr,
                                            ip.defined() ? ip.data_ptr<float>() : nullptr, c.data_ptr<double>(), c.stride(0),
                                            phase_frames.data_ptr<float>(), prefix.data_ptr<double>(), st));
    } else {
        check_status(ddsp_b200_phase(f0.data_ptr<float>(), f0.stride(0), f0.stride(1), (int)B, (int)F, (int)hop, sr,
                                     ip.defined() ? ip.data_ptr<float>() : nullptr, infer ? 1 : 0,
                                     phase_frames.data_ptr<float>(), prefix.data_ptr<double>(),
                                     phase_full.defined() ? phase_full.data_ptr<float>() : nullptr, st));
    }
    return {phase_frames, prefix, phase_full};
}

Tensor combsubfast(const Tensor& harmonic_magnitude, const Tensor& harmonic_phase, const Tensor& noise_magnitude,
                   const Tensor& f0_frames, const Tensor& prefix, int64_t hop, double sr,
                   const c10::optional<Tensor>& noise_u, int64_t seed, const c10::optional<Tensor>& window,
                   const c10::optional<Tensor>& seed_device, int64_t hop_offset, const c10::optional<Tensor>& out) {
    Tensor hm = need_f32(harmonic_magnitude, "harmonic_magnitude"), hp = need_f32(harmonic_phase, "harmonic_phase"),
           nm = need_f32(noise_magnitude, "noise_magnitude");
    common_views(hm, hp, nm);
    Tensor f0 = f0_2d(f0_frames);
    const int64_t B = f0.size(0), F = f0.size(1), T = F * hop;
    for (const Tensor* t : {&hm, &hp, &nm})
        TORCH_CHECK_VALUE(t->size(0) == B && t->size(1) == F && t->size(2) == hop + 1,
                          "control tensors must be (B, Frame, block_size + 1)");
    TORCH_CHECK_VALUE(prefix.is_cuda() && prefix.scalar_type() == torch::kFloat64 && prefix.is_contiguous() &&
                          prefix.numel() == B * F,
                      "prefix must be the (B, Frame) float64 tensor returned by ddsp_b200::phase");
    c10::cuda::CUDAGuard guard(f0.device());
    Tensor u = checked_noise(noise_u, B, T);
    Tensor w;
    if (window.has_value()) {
        w = need_f32(*window, "window").contiguous();
        TORCH_CHECK_VALUE(w.numel() == 2 * hop, "window must have 2*block_size entries");
    }
    const uint64_t* sd = nullptr;
    if (seed_device.has_value()) {
        TORCH_CHECK_VALUE(seed_device->is_cuda() && seed_device->scalar_type() == torch::kInt64 && seed_device->numel() == 1,
                          "seed_device must be a one-element int64 CUDA tensor");
        sd = reinterpret_cast<const uint64_t*>(seed_device->data_ptr<int64_t>());
    }
    Tensor signal = out.has_value() ? *out : torch::empty({B, T}, f0.options());
    if (out.has_value())
        TORCH_CHECK_VALUE(signal.is_cuda() && signal.scalar_type() == torch::kFloat32 && signal.is_contiguous() &&
                              signal.numel() == B * T, "out must be a contiguous (B, T) float32 CUDA tensor");
    void* st = at::cuda::getCurrentCUDAStream().stream();
    check_status(ddsp_b200_combsubfast_stream(
        hm.data_ptr<float>(), hp.data_ptr<float>(), nm.data_ptr<float>(), hm.stride(0), hm.stride(1), f0.data_ptr<float>(),
        f0.stride(0), f0.stride(1), prefix.data_ptr<double>(), u.defined() ? u.data_ptr<float>() : nullptr,
        (uint64_t)seed & ((1ull << 62) - 1), sd, hop_offset, w.defined() ? w.data_ptr<float>() : nullptr, (int)B, (int)F,
        (int)hop, sr, signal.data_ptr<float>(), st));
    return signal;
}

// Stage A + stage B of CombSubFast in ONE operator, for callers whose control rows do not depend on the phase (rows that
// come from outside, a streaming plugin): one dispatcher crossing and no Python between the two stages.
// Returns (signal, phase_frames, prefix).
std::tuple<Tensor, Tensor, Tensor> combsubfast_ab(const Tensor& harmonic_magnitude, const Tensor& harmonic_phase,
                                                  const Tensor& noise_magnitude, const Tensor& f0_frames, int64_t hop,
                                                  double sr, const c10::optional<Tensor>& initial_phase,
                                                  const c10::optional<Tensor>& carry, const c10::optional<Tensor>& noise_u,
                                                  int64_t seed, const c10::optional<Tensor>& window,
                                                  const c10::optional<Tensor>& seed_device, int64_t hop_offset) {
    auto [phase_frames, prefix, unused] = phase(f0_frames, hop, sr, initial_phase, true, false, carry);
    (void)unused;
    Tensor signal = combsubfast(harmonic_magnitude, harmonic_phase, noise_magnitude, f0_frames, prefix, hop, sr, noise_u, seed,
                                window, seed_device, hop_offset, c10::nullopt);
    return {signal, phase_frames, prefix};
}

// One block of a carried-state CombSubFast stream (streaming.CombSubFastStream.push) in one operator:
//   f0_buf (B,cap), rows_buf (B,cap,3K): the stream's linear buffers; frames [end-t, end) are the tail kept from the
//   previous block.  The k new frames are copied behind the tail, stage A continues from `carry` over the window
//   [end-t, end+k), stage B synthesises the window with the noise hop index `hop_offset`.
// Returns (signal of the window (B, (t+k)*hop), phase_frames (B,t+k), prefix (B,t+k)).
std::tuple<Tensor, Tensor, Tensor> csf_stream_push(Tensor f0_buf, Tensor rows_buf, const Tensor& f0_new,
                                                   const Tensor& harmonic_magnitude, const Tensor& harmonic_phase,
                                                   const Tensor& noise_magnitude, int64_t end, int64_t t, int64_t hop,
                                                   double sr, const c10::optional<Tensor>& initial_phase,
                                                   const c10::optional<Tensor>& carry, int64_t seed,
                                                   const c10::optional<Tensor>& window, int64_t hop_offset) {
    Tensor f0n = f0_2d(f0_new);
    const int64_t B = f0n.size(0), k = f0n.size(1), K = hop + 1;
    TORCH_CHECK_VALUE(k >= 1, "a push needs at least one new frame");
    TORCH_CHECK_VALUE(f0_buf.dim() == 2 && rows_buf.dim() == 3 && f0_buf.size(0) == B && rows_buf.size(0) == B &&
                          rows_buf.size(2) == 3 * K && f0_buf.size(1) == rows_buf.size(1) && end + k <= f0_buf.size(1) &&
                          t >= 0 && t <= end,
                      "stream buffers do not fit the block (B, cap) / (B, cap, 3*(block_size+1))");
    const Tensor* rows[3] = {&harmonic_magnitude, &harmonic_phase, &noise_magnitude};
    for (const Tensor* r : rows) {
        need_f32(*r, "control rows");
        TORCH_CHECK_VALUE(r->dim() == 3 && r->size(0) == B && r->size(1) == k && r->size(2) == K,
                          "control rows must be (B, k, block_size + 1)");
    }
    c10::cuda::CUDAGuard guard(f0n.device());
    f0_buf.narrow(1, end, k).copy_(f0n);
    Tensor dst = rows_buf.narrow(1, end, k);
    const Tensor &hm = harmonic_magnitude, &hp = harmonic_phase, &nm = noise_magnitude;
    if (hm.stride(2) == 1 && hm.strides() == hp.strides() && hm.strides() == nm.strides() &&
        hp.data_ptr<float>() == hm.data_ptr<float>() + K && nm.data_ptr<float>() == hp.data_ptr<float>() + K) {
        // the split views of one (B,k,3K) tensor: one copy
        dst.copy_(hm.as_strided({B, k, 3 * K}, hm.strides()));
    } else {
        for (int i = 0; i < 3; ++i) dst.narrow(2, i * K, K).copy_(*rows[i]);
    }
    Tensor f0_win = f0_buf.narrow(1, end - t, t + k), rows_win = rows_buf.narrow(1, end - t, t + k);
    auto [phase_frames, prefix, unused] = phase(f0_win, hop, sr, carry.has_value() ? c10::nullopt : initial_phase, true, false, carry);
    (void)unused;
    Tensor signal = combsubfast(rows_win.narrow(2, 0, K), rows_win.narrow(2, K, K), rows_win.narrow(2, 2 * K, K), f0_win, prefix,
                                hop, sr, c10::nullopt, seed, window, c10::nullopt, hop_offset, c10::nullopt);
    return {signal, phase_frames, prefix};
}

std::tuple<Tensor, Tensor, Tensor> filter_model(bool is_sins, const Tensor& c0_, const Tensor& c1_, const Tensor& c2_,
                                                const Tensor& f0_frames, const Tensor& aux, int64_t hop, double sr,
                                                const c10::optional<Tensor>& noise_u, int64_t seed) {
    Tensor c0 = need_f32(c0_, "control tensor 0"), c1 = need_f32(c1_, "control tensor 1"), c2 = need_f32(c2_, "control tensor 2");
    common_views(c0, c1, c2);
    Tensor f0 = f0_2d(f0_frames);
    const int64_t B = f0.size(0), F = f0.size(1), T = F * hop;
    for (const Tensor* t : {&c0, &c1, &c2})
        TORCH_CHECK_VALUE(t->size(0) == B && t->size(1) == F, "control tensors must be (B, Frame, K)");
    c10::cuda::CUDAGuard guard(f0.device());
    Tensor u = checked_noise(noise_u, B, T);
    auto opts = f0.options();
    Tensor signal = torch::empty({B, T}, opts), harmonic = torch::empty({B, T}, opts), noise = torch::empty({B, T}, opts);
    const int k0 = (int)c0.size(2), k1 = (int)c1.size(2), k2 = (int)c2.size(2);
    void* st = at::cuda::getCurrentCUDAStream().stream();
    if (is_sins) {
        Tensor ph = need_f32(aux, "phase").contiguous();
        TORCH_CHECK_VALUE(ph.numel() == B * T, "phase must be the (B, T) tensor returned by ddsp_b200::phase(full_rate=True)");
        const size_t wsb = ddsp_b200_sins_workspace_bytes((int)B, (int)F, k0, k1, k2);
        Tensor ws = torch::empty({(int64_t)wsb}, opts.dtype(torch::kUInt8));
        check_status(ddsp_b200_sins(c0.data_ptr<float>(), k0, c1.data_ptr<float>(), k1, c2.data_ptr<float>(), k2, c0.stride(0),
                                    c0.stride(1), f0.data_ptr<float>(), f0.stride(0), f0.stride(1), ph.data_ptr<float>(),
                                    u.defined() ? u.data_ptr<float>() : nullptr, (uint64_t)seed & ((1ull << 62) - 1), (int)B,
                                    (int)F, (int)hop, sr, signal.data_ptr<float>(), harmonic.data_ptr<float>(),
                                    noise.data_ptr<float>(), ws.data_ptr(), wsb, st));
    } else {
        TORCH_CHECK_VALUE(aux.is_cuda() && aux.scalar_type() == torch::kFloat64 && aux.is_contiguous() && aux.numel() == B * F,
                          "prefix must be the (B, Frame) float64 tensor returned by ddsp_b200::phase");
        const size_t wsb = ddsp_b200_combsub_workspace_bytes((int)B, (int)F, k0, k1, k2);
        Tensor ws = torch::empty({(int64_t)wsb}, opts.dtype(torch::kUInt8));
        check_status(ddsp_b200_combsub(c0.data_ptr<float>(), k0, c1.data_ptr<float>(), k1, c2.data_ptr<float>(), k2, c0.stride(0),
                                       c0.stride(1), f0.data_ptr<float>(), f0.stride(0), f0.stride(1), aux.data_ptr<double>(),
                                       nullptr, u.defined() ? u.data_ptr<float>() : nullptr,
                                       (uint64_t)seed & ((1ull << 62) - 1), (int)B, (int)F, (int)hop, sr,
                                       signal.data_ptr<float>(), harmonic.data_ptr<float>(), noise.data_ptr<float>(),
                                       ws.data_ptr(), wsb, st));
    }
    return {signal, harmonic, noise};
}

std::tuple<Tensor, Tensor, Tensor> combsub(const Tensor& group_delay, const Tensor& harmonic_magnitude,
                                           const Tensor& noise_magnitude, const Tensor& f0_frames, const Tensor& prefix,
                                           int64_t hop, double sr, const c10::optional<Tensor>& noise_u, int64_t seed) {
    return filter_model(false, group_delay, harmonic_magnitude, noise_magnitude, f0_frames, prefix, hop, sr, noise_u, seed);
}

std::tuple<Tensor, Tensor, Tensor> sins(const Tensor& amplitudes, const Tensor& group_delay, const Tensor& noise_magnitude,
                                        const Tensor& f0_frames, const Tensor& phase_full, int64_t hop, double sr,
                                        const c10::optional<Tensor>& noise_u, int64_t seed) {
    return filter_model(true, amplitudes, group_delay, noise_magnitude, f0_frames, phase_full, hop, sr, noise_u, seed);
}

int64_t last_launch_count() { return ddsp_b200_last_launch_count(); }

}  // namespace

TORCH_LIBRARY(ddsp_b200, m) {
    m.def("phase(Tensor f0_frames, int block_size, float sampling_rate, Tensor? initial_phase=None, bool infer=True, "
          "bool full_rate=False, Tensor? carry=None) -> (Tensor, Tensor, Tensor)");
    m.def("combsubfast(Tensor harmonic_magnitude, Tensor harmonic_phase, Tensor noise_magnitude, Tensor f0_frames, "
          "Tensor prefix, int block_size, float sampling_rate, Tensor? noise_u=None, int seed=0, Tensor? window=None, "
          "Tensor? seed_device=None, int hop_offset=0, Tensor? out=None) -> Tensor");
    m.def("combsubfast_ab(Tensor harmonic_magnitude, Tensor harmonic_phase, Tensor noise_magnitude, Tensor f0_frames, "
          "int block_size, float sampling_rate, Tensor? initial_phase=None, Tensor? carry=None, Tensor? noise_u=None, "
          "int seed=0, Tensor? window=None, Tensor? seed_device=None, int hop_offset=0) -> (Tensor, Tensor, Tensor)");
    m.def("csf_stream_push(Tensor(a!) f0_buf, Tensor(b!) rows_buf, Tensor f0_new, Tensor harmonic_magnitude, "
          "Tensor harmonic_phase, Tensor noise_magnitude, int end, int tail, int block_size, float sampling_rate, "
          "Tensor? initial_phase=None, Tensor? carry=None, int seed=0, Tensor? window=None, int hop_offset=0) "
          "-> (Tensor, Tensor, Tensor)");
    m.def("combsub(Tensor group_delay, Tensor harmonic_magnitude, Tensor noise_magnitude, Tensor f0_frames, Tensor prefix, "
          "int block_size, float sampling_rate, Tensor? noise_u=None, int seed=0) -> (Tensor, Tensor, Tensor)");
    m.def("sins(Tensor amplitudes, Tensor group_delay, Tensor noise_magnitude, Tensor f0_frames, Tensor phase_full, "
          "int block_size, float sampling_rate, Tensor? noise_u=None, int seed=0) -> (Tensor, Tensor, Tensor)");
    m.def("last_launch_count() -> int");
}

TORCH_LIBRARY_IMPL(ddsp_b200, CUDA, m) {
    m.impl("phase", &phase);
    m.impl("combsubfast", &combsubfast);
    m.impl("combsubfast_ab", &combsubfast_ab);
    m.impl("csf_stream_push", &csf_stream_push);
    m.impl("combsub", &combsub);
    m.impl("sins", &sins);
}

TORCH_LIBRARY_IMPL(ddsp_b200, CompositeExplicitAutograd, m) { m.impl("last_launch_count", &last_launch_count); }
