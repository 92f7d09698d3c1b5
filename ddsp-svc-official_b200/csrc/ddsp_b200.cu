// ddsp_b200.cu -- the C ABI (include/ddsp_b200.h): argument checks + kernel launches.
// No torch types, no allocation, no synchronisation; everything is enqueued on the caller's stream.
#include "../../include/ddsp_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <initializer_list>
#include <mutex>

#include "combsubfast.cuh"
#include "combsubfast_bwd.cuh"
#include "control.cuh"
#include "excite.cuh"
#include "frontend.cuh"
#include "gemm_attn.cuh"
#include "gemm_tc.cuh"
#include "ltvfir.cuh"
#include "phase.cuh"

namespace {

thread_local int g_last_cuda_error = 0;
thread_local int g_launches = 0;
thread_local int64_t g_noise_hop_offset = 0;     // set by the *_stream entry points of the frequency_filter models

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return DDSP_B200_ERR_CUDA;
}
#define CUDA_TRY(expr)                                   \
    do {                                                 \
        cudaError_t _e = (expr);                         \
        if (_e != cudaSuccess) return cuda_fail(_e);     \
    } while (0)
#define LAUNCH_CHECK()                                   \
    do {                                                 \
        ++g_launches;                                    \
        cudaError_t _e = cudaGetLastError();             \
        if (_e != cudaSuccess) return cuda_fail(_e);     \
    } while (0)

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

using LtvKernel = void (*)(const ddsp::LtvParams);
// Specialised instantiations for the combinations the synthesizers use + generic fallbacks.
LtvKernel ltv_ir_select(int enc, int win, int n_mag = 0) {
    if (enc == DDSP_B200_MAG_EXP && win == DDSP_B200_WINDOW_DYNAMIC && n_mag == 512)      // CombSub's harmonic filter (vocoder.py:541)
        return ddsp::ltv_ir_kernel<DDSP_B200_MAG_EXP, DDSP_B200_WINDOW_DYNAMIC, 512>;
    if (enc == DDSP_B200_MAG_ALLPASS_TANH && win == DDSP_B200_WINDOW_NONE)
        return ddsp::ltv_ir_kernel<DDSP_B200_MAG_ALLPASS_TANH, DDSP_B200_WINDOW_NONE>;
    if (enc == DDSP_B200_MAG_EXP && win == DDSP_B200_WINDOW_DYNAMIC)
        return ddsp::ltv_ir_kernel<DDSP_B200_MAG_EXP, DDSP_B200_WINDOW_DYNAMIC>;
    if (enc == DDSP_B200_MAG_EXP && win == DDSP_B200_WINDOW_HANN)
        return ddsp::ltv_ir_kernel<DDSP_B200_MAG_EXP, DDSP_B200_WINDOW_HANN>;
    return ddsp::ltv_ir_kernel<-1, -1>;
}
LtvKernel ltv_conv_select(int amode, int n_mag = 512) {
    if (n_mag == 256)      // L = 510: packed half-frame convolution (ltv_conv510_kernel)
        return amode == 0 ? ddsp::ltv_conv510_kernel<0> : amode == 1 ? ddsp::ltv_conv510_kernel<1> : ddsp::ltv_conv510_kernel<2>;
    return amode == 0 ? ddsp::ltv_conv_kernel<0> : amode == 1 ? ddsp::ltv_conv_kernel<1> : ddsp::ltv_conv_kernel<2>;
}

// Immutable per-device tables (FFT twiddles + exact sqrt-Hann window; Bluestein chirps for the
// L=510 and L=1022 impulse responses), filled on first use.
__device__ __align__(16) float g_tables[ddsp::kTableBytes / 4];
__device__ __align__(16) float g_chirp[3][ddsp::kChirpFloats];     // L=510, L=1022, L=510 dual (K=510)
const float* g_chirp_ptr[64][3] = {{nullptr, nullptr, nullptr}};

std::mutex g_init_mutex;
// 0 = nothing set up, 1 = attributes + FFT / window tables, 2 = + Bluestein chirp tables (frequency_filter models)
std::atomic<int> g_device_level[64];
const float* g_tables_ptr[64] = {nullptr};

// One-time per-device setup: opt-in shared memory sizes and the constant tables.  This is the only place the
// library synchronises (once per device and level, on the first call that needs it); afterwards an entry point
// reads one atomic and takes no lock.  The chirp tables (an O(N^2) fp64 DFT, ~1 ms) are built only for callers
// of the frequency_filter path.
int ensure_device_ready(cudaStream_t st, const float** tables, int level = 2) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DDSP_B200_ERR_UNSUPPORTED;
    if (g_device_level[dev].load(std::memory_order_acquire) >= level) {
        *tables = g_tables_ptr[dev];
        return 0;
    }
    std::lock_guard<std::mutex> lock(g_init_mutex);
    int have = g_device_level[dev].load(std::memory_order_acquire);
    if (have < level) {
        // the one-time setup synchronises the stream, which a capturing stream cannot do: ask for one warm-up call
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        CUDA_TRY(cudaStreamIsCapturing(st, &cap));
        if (cap != cudaStreamCaptureStatusNone) return DDSP_B200_ERR_CAPTURE;
    }
    if (have < 1) {
        CUDA_TRY(cudaFuncSetAttribute(ddsp::combsubfast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ddsp::kCsfSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(ddsp::combsubfast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ddsp::kCsfSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(ddsp::combsubfast_backward_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, ddsp::kCsbSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(ddsp::combsubfast_backward_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, ddsp::kCsbSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(ddsp::performer_project_features_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, ddsp::kPpfSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(ddsp::performer_project_features_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, ddsp::kPpfSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(ddsp::performer_attention_small_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, ddsp::kPasSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(ddsp::performer_context_partial_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, ddsp::kPasSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(ddsp::performer_output_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, ddsp::kPasSmemBytes));
        float* ptr = nullptr;
        CUDA_TRY(cudaGetSymbolAddress((void**)&ptr, g_tables));
        ddsp::fft_tables_kernel<<<4, 256, 0, st>>>(reinterpret_cast<float4*>(ptr), ptr + 2048);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(st));
        g_tables_ptr[dev] = ptr;
        have = 1;
        g_device_level[dev].store(1, std::memory_order_release);
    }
    if (have < level) {
        for (LtvKernel fn : {ltv_ir_select(DDSP_B200_MAG_EXP, DDSP_B200_WINDOW_DYNAMIC, 512),
                             ltv_ir_select(DDSP_B200_MAG_ALLPASS_TANH, DDSP_B200_WINDOW_NONE),
                             ltv_ir_select(DDSP_B200_MAG_EXP, DDSP_B200_WINDOW_DYNAMIC),
                             ltv_ir_select(DDSP_B200_MAG_EXP, DDSP_B200_WINDOW_HANN), ltv_ir_select(-1, -1)})
            CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ddsp::kLtvIrSmemBytes));
        for (int am = 0; am < 3; ++am)
            for (int nm : {256, 512})
                CUDA_TRY(cudaFuncSetAttribute(ltv_conv_select(am, nm), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              ddsp::kLtvConvSmemBytes));
        float* cptr = nullptr;
        CUDA_TRY(cudaGetSymbolAddress((void**)&cptr, g_chirp));
        CUDA_TRY(cudaFuncSetAttribute(ddsp::ltv_ir_dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ddsp::kLtvDualSmemBytes));
        for (int v = 0; v < 3; ++v) {
            const int L = v == 1 ? 1022 : 510, K = v == 2 ? 510 : L / 2 + 1, n_out = v == 1 ? 512 : 510;
            float* base = cptr + (size_t)v * ddsp::kChirpFloats;
            ddsp::chirp_tables_kernel<<<4, 256, 0, st>>>(reinterpret_cast<float2*>(base),
                                                         reinterpret_cast<float2*>(base) + 512, L, K, n_out);
            CUDA_TRY(cudaGetLastError());
            g_chirp_ptr[dev][v] = base;
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        g_device_level[dev].store(2, std::memory_order_release);
    }
    *tables = g_tables_ptr[dev];
    return 0;
}

// Launch with programmatic stream serialisation: the grid may be set up while its predecessor in the
// stream drains; the kernel itself waits in cudaGridDependencySynchronize() before touching memory.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Split the frame pairs of every clip into runs (one warp each) so that all runs of the launch fit
// one resident wave of `slots` warps and differ by at most one pair: as many runs per clip as the
// slots allow, the first `run_rem` of them one pair longer.
void csf_partition(ddsp::CsfParams& P, int B, int64_t slots) {
    int64_t R = slots / B;
    if (R < 1) R = 1;
    if (R > P.pairs_per_clip) R = P.pairs_per_clip;
    P.runs_per_clip = (int)R;
    P.run_len = P.pairs_per_clip / (int)R;
    P.run_rem = P.pairs_per_clip % (int)R;
}

inline int64_t grid_for(int64_t total, int per_block, int64_t cap) {
    int64_t g = (total + per_block - 1) / per_block;
    if (g < 1) g = 1;
    return g > cap ? cap : g;
}

}  // namespace

extern "C" {

int ddsp_b200_version(void) { return DDSP_B200_ABI_VERSION; }

const char* ddsp_b200_strerror(int status) {
    switch (status) {
        case DDSP_B200_OK: return "ok";
        case DDSP_B200_ERR_INVALID_ARGUMENT: return "invalid argument (null pointer, bad size or stride)";
        case DDSP_B200_ERR_UNSUPPORTED: return "unsupported configuration (block_size must be 512; n_mag in {256,512})";
        case DDSP_B200_ERR_WORKSPACE: return "workspace too small";
        case DDSP_B200_ERR_CUDA: return "CUDA runtime error (see ddsp_b200_last_cuda_error)";
        case DDSP_B200_ERR_BATCH_MISMATCH: return "batch size of audio and impulse response must be the same";
        case DDSP_B200_ERR_CAPTURE:
            return "the first call on a device sets up constant tables and cannot run under CUDA stream capture: issue one warm-up call first";
        default: return "unknown status";
    }
}

int ddsp_b200_last_cuda_error(void) { return g_last_cuda_error; }
int ddsp_b200_last_launch_count(void) { return g_launches; }

int ddsp_b200_upsample(const float* x, int64_t sB, int64_t sF, int64_t sC, int B, int F, int C, int factor,
                       float* y, void* stream) {
    g_launches = 0;
    if (!x || !y || B <= 0 || F <= 0 || C <= 0 || factor <= 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    const int64_t total = (int64_t)B * F * factor * C;
    // torch: rwidth = (float)(in-1) / (out-1) with in = F+1, out = F*factor+1 (align_corners=True)
    const float rwidth = (float)F / (float)((int64_t)F * factor);
    ddsp::upsample_kernel<<<(unsigned)grid_for(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(
        x, sB, sF, sC, B, F, C, factor, rwidth, y);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

size_t ddsp_b200_fo_to_rot_workspace_bytes(int B, int64_t T) {
    if (B <= 0 || T <= 0) return 0;
    const int64_t nchunks = (T + ddsp::kRotChunk - 1) / ddsp::kRotChunk;
    return (size_t)B * (size_t)nchunks * sizeof(double);
}

int ddsp_b200_fo_to_rot(const float* fo, int B, int64_t T, double sr, const float* initial_phase, int precise,
                        float* rot, void* workspace, size_t workspace_bytes, void* stream) {
    g_launches = 0;
    if (!fo || !rot || !workspace || B <= 0 || T <= 0 || !(sr > 0)) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < ddsp_b200_fo_to_rot_workspace_bytes(B, T)) return DDSP_B200_ERR_WORKSPACE;
    const int nchunks = (int)((T + ddsp::kRotChunk - 1) / ddsp::kRotChunk);
    if (B > 65535) return DDSP_B200_ERR_UNSUPPORTED;
    const dim3 grid(nchunks, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (precise) {
        double* sums = (double*)workspace;
        ddsp::rot_chunk_sums_kernel<double><<<grid, 256, 0, st>>>(fo, T, nchunks, sr, sums);
        LAUNCH_CHECK();
        ddsp::rot_scan_sums_kernel<double><<<B, 1024, 0, st>>>(nchunks, sums);
        LAUNCH_CHECK();
        ddsp::rot_apply_kernel<double><<<grid, 256, 0, st>>>(fo, T, nchunks, sr, initial_phase, sums, rot);
        LAUNCH_CHECK();
    } else {
        float* sums = (float*)workspace;
        ddsp::rot_chunk_sums_kernel<float><<<grid, 256, 0, st>>>(fo, T, nchunks, (float)sr, sums);
        LAUNCH_CHECK();
        ddsp::rot_scan_sums_kernel<float><<<B, 1024, 0, st>>>(nchunks, sums);
        LAUNCH_CHECK();
        ddsp::rot_apply_kernel<float><<<grid, 256, 0, st>>>(fo, T, nchunks, (float)sr, initial_phase, sums, rot);
        LAUNCH_CHECK();
    }
    return DDSP_B200_OK;
}

int ddsp_b200_remove_above_fmax(const float* amplitudes, int64_t aB, int64_t aF, const float* pitch, int64_t pB,
                                int64_t pF, float fmax, int level_start, int B, int F, int K, float* out,
                                void* stream) {
    g_launches = 0;
    if (!amplitudes || !pitch || !out || B <= 0 || F <= 0 || K <= 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    const int64_t total = (int64_t)B * F * K;
    ddsp::remove_above_fmax_kernel<<<(unsigned)grid_for(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(
        amplitudes, aB, aF, pitch, pB, pF, fmax, level_start, B, F, K, out);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

static int phase_impl(const float* f0_frames, int64_t fB, int64_t fF, int B, int F, int hop, double sr,
                      const float* initial_phase, const double* carry, int64_t carry_stride, float* phase_frames,
                      double* prefix, float* phase_full, void* stream) {
    g_launches = 0;
    if (!f0_frames || !phase_frames || !prefix || B <= 0 || F <= 0 || !(sr > 0)) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (hop != ddsp::kHop) return DDSP_B200_ERR_UNSUPPORTED;
    if ((int64_t)F * hop >= (1ll << 24)) return DDSP_B200_ERR_UNSUPPORTED;   // fp32-exact sample index (torch's own limit)
    cudaStream_t st = (cudaStream_t)stream;
    const double inv_sr = 1.0 / sr;
    const int64_t hops = (int64_t)B * F;
    // Small clips: one launch does totals + scan per clip.  Otherwise spread the totals over the chip.
    if (F <= 64 || (int64_t)B * 32 >= (int64_t)sm_count() * 64) {
        ddsp::phase_fused_kernel<<<B, 1024, 0, st>>>(f0_frames, fB, fF, F, inv_sr, initial_phase, carry, carry_stride, prefix,
                                                     phase_frames);
        LAUNCH_CHECK();
    } else {
        ddsp::hop_totals_kernel<<<(unsigned)((hops + 8 * ddsp::kHopsPerWarp - 1) / (8 * ddsp::kHopsPerWarp)), 256, 0, st>>>(
            f0_frames, fB, fF, B, F, prefix);
        LAUNCH_CHECK();
        CUDA_TRY(launch_pdl(ddsp::phase_scan_kernel, dim3(B), dim3(1024), 0, st, f0_frames, fB, fF, F, inv_sr, initial_phase,
                            carry, carry_stride, prefix, phase_frames));
        LAUNCH_CHECK();
    }
    if (phase_full) {
        ddsp::phase_full_kernel<<<(unsigned)((hops + 7) / 8), 256, 0, st>>>(f0_frames, fB, fF, B, F, inv_sr,
                                                                             initial_phase, prefix, phase_full);
        LAUNCH_CHECK();
    }
    return DDSP_B200_OK;
}

int ddsp_b200_phase(const float* f0_frames, int64_t fB, int64_t fF, int B, int F, int hop, double sr,
                    const float* initial_phase, int precise, float* phase_frames, double* prefix,
                    float* phase_full, void* stream) {
    (void)precise;   // the fused path always accumulates in fp64 (inference path, core.py:40)
    return phase_impl(f0_frames, fB, fF, B, F, hop, sr, initial_phase, nullptr, 0, phase_frames, prefix, phase_full,
                      stream);
}

int ddsp_b200_phase_stream(const float* f0_frames, int64_t fB, int64_t fF, int B, int F, int hop, double sr,
                           const float* initial_phase, const double* carry, int64_t carry_stride,
                           float* phase_frames, double* prefix, void* stream) {
    if (carry && (carry == prefix || carry_stride < 0)) return DDSP_B200_ERR_INVALID_ARGUMENT;
    return phase_impl(f0_frames, fB, fF, B, F, hop, sr, initial_phase, carry, carry_stride, phase_frames, prefix,
                      nullptr, stream);
}

int ddsp_b200_phase_stream_full(const float* f0_frames, int64_t fB, int64_t fF, int B, int F, int hop, double sr,
                                const float* initial_phase, const double* carry, int64_t carry_stride,
                                float* phase_frames, double* prefix, float* phase_full, void* stream) {
    if (carry && (carry == prefix || carry_stride < 0)) return DDSP_B200_ERR_INVALID_ARGUMENT;
    return phase_impl(f0_frames, fB, fF, B, F, hop, sr, initial_phase, carry, carry_stride, phase_frames, prefix,
                      phase_full, stream);
}

static int combsubfast_impl(const float* harmonic_magnitude, const float* harmonic_phase, const float* noise_magnitude,
                            int64_t cB, int64_t cF, const float* f0_frames, int64_t fB, int64_t fF,
                            const double* prefix, const float* noise_u, uint64_t seed,
                            const uint64_t* seed_device, int64_t hop_offset, const float* window, int B, int F,
                            int hop, double sr, float* signal, void* stream) {
    g_launches = 0;
    if (!harmonic_magnitude || !harmonic_phase || !noise_magnitude || !f0_frames || !prefix || !signal || B <= 0 ||
        F <= 0 || !(sr > 0))
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (hop != ddsp::kHop) return DDSP_B200_ERR_UNSUPPORTED;
    if ((int64_t)F * hop >= (1ll << 24)) return DDSP_B200_ERR_UNSUPPORTED;
    ddsp::CsfParams P;
    if (int rc = ensure_device_ready((cudaStream_t)stream, &P.tables, 1)) return rc;
    P.hm = harmonic_magnitude; P.hp = harmonic_phase; P.nm = noise_magnitude;
    P.cB = cB; P.cF = cF;
    P.f0_frames = f0_frames; P.fB = fB; P.fF = fF;
    P.prefix = prefix; P.noise_u = noise_u; P.window = window;
    P.signal = signal; P.seed = seed; P.seed_device = seed_device; P.B = B; P.F = F;
    // hop h of this call is hop hop_offset + h of the stream: the in-kernel noise is keyed by
    // (hop*32 + lane)*c + key (common.cuh noise_seed), so the offset folds into the per-clip key
    P.key_offset = (uint32_t)((uint64_t)hop_offset * 32ull) * 0x9E3779B1u;
    P.pairs_per_clip = (F + 2) / 2;                 // frames 0..F in pairs
    csf_partition(P, B, (int64_t)sm_count() * ddsp::kCsfWarps);
    P.inv_sr = 1.0 / sr; P.sr = (float)sr;
    const int64_t runs = (int64_t)B * P.runs_per_clip;
    const unsigned grid = (unsigned)((runs + ddsp::kCsfWarps - 1) / ddsp::kCsfWarps);
    if (P.runs_per_clip > 1) {
        const int n_seams = B * (P.runs_per_clip - 1);
        CUDA_TRY(launch_pdl(ddsp::csf_zero_seams_kernel, dim3(n_seams), dim3(128), 0, (cudaStream_t)stream, signal, F,
                            P.run_len, P.run_rem, P.runs_per_clip, n_seams));
        LAUNCH_CHECK();
    }
    // Programmatic dependent launch: the CTAs may be scheduled and stage their tables while the
    // preceding kernel (seam zeroing / stage A) drains; they wait in cudaGridDependencySynchronize()
    // before touching anything a previous kernel wrote.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ddsp::kCsfThreads);
    cfg.dynamicSmemBytes = ddsp::kCsfSmemBytes; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (noise_u) CUDA_TRY(cudaLaunchKernelEx(&cfg, ddsp::combsubfast_kernel<true>, P));
    else CUDA_TRY(cudaLaunchKernelEx(&cfg, ddsp::combsubfast_kernel<false>, P));
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_combsubfast(const float* harmonic_magnitude, const float* harmonic_phase, const float* noise_magnitude,
                          int64_t cB, int64_t cF, const float* f0_frames, int64_t fB, int64_t fF,
                          const double* prefix, const float* initial_phase, const float* noise_u, uint64_t seed,
                          const float* window, int B, int F, int hop, double sr, float* signal, void* stream) {
    (void)initial_phase;   // carried by `prefix`
    return combsubfast_impl(harmonic_magnitude, harmonic_phase, noise_magnitude, cB, cF, f0_frames, fB, fF, prefix,
                            noise_u, seed, nullptr, 0, window, B, F, hop, sr, signal, stream);
}

int ddsp_b200_combsubfast_stream(const float* harmonic_magnitude, const float* harmonic_phase,
                                 const float* noise_magnitude, int64_t cB, int64_t cF, const float* f0_frames,
                                 int64_t fB, int64_t fF, const double* prefix, const float* noise_u, uint64_t seed,
                                 const uint64_t* seed_device, int64_t hop_offset, const float* window, int B, int F,
                                 int hop, double sr, float* signal, void* stream) {
    if (hop_offset < 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    return combsubfast_impl(harmonic_magnitude, harmonic_phase, noise_magnitude, cB, cF, f0_frames, fB, fF, prefix,
                            noise_u, seed, seed_device, hop_offset, window, B, F, hop, sr, signal, stream);
}

int ddsp_b200_combsubfast_backward(const float* harmonic_magnitude, const float* harmonic_phase,
                                   const float* noise_magnitude, int64_t cB, int64_t cF, const float* f0_frames,
                                   int64_t fB, int64_t fF, const double* prefix, const float* noise_u, uint64_t seed,
                                   const float* window, const float* grad_signal, int B, int F, int hop, double sr,
                                   float* grad_harmonic_magnitude, float* grad_harmonic_phase,
                                   float* grad_noise_magnitude, int64_t gB, int64_t gF, void* stream) {
    g_launches = 0;
    if (!harmonic_magnitude || !harmonic_phase || !noise_magnitude || !f0_frames || !prefix || !grad_signal ||
        !grad_harmonic_magnitude || !grad_harmonic_phase || !grad_noise_magnitude || B <= 0 || F <= 0 || !(sr > 0))
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (hop != ddsp::kHop) return DDSP_B200_ERR_UNSUPPORTED;
    if ((int64_t)F * hop >= (1ll << 24)) return DDSP_B200_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    ddsp::CsbParams PB;
    ddsp::CsfParams& P = PB.fwd;
    if (int rc = ensure_device_ready(st, &P.tables, 1)) return rc;
    P.hm = harmonic_magnitude; P.hp = harmonic_phase; P.nm = noise_magnitude;
    P.cB = cB; P.cF = cF;
    P.f0_frames = f0_frames; P.fB = fB; P.fF = fF;
    P.prefix = prefix; P.noise_u = noise_u; P.window = window;
    P.signal = nullptr; P.seed = seed; P.seed_device = nullptr; P.key_offset = 0; P.B = B; P.F = F;
    P.pairs_per_clip = (F + 2) / 2;
    csf_partition(P, B, (int64_t)sm_count() * ddsp::kCsbWarps);
    P.inv_sr = 1.0 / sr; P.sr = (float)sr;
    PB.grad_signal = grad_signal;
    PB.ghm = grad_harmonic_magnitude; PB.ghp = grad_harmonic_phase; PB.gnm = grad_noise_magnitude;
    PB.gB = gB; PB.gF = gF;
    // Filter row F-1 serves frames F-1 and F (vocoder.py:473,476): its gradient is the sum of two
    // atomic adds onto zeros; every other row is written exactly once.
    for (float* g : {PB.ghm, PB.ghp, PB.gnm})
        CUDA_TRY(cudaMemset2DAsync(g + (int64_t)(F - 1) * gF, (size_t)gB * sizeof(float), 0,
                                   (size_t)(hop + 1) * sizeof(float), (size_t)B, st));
    const int64_t runs = (int64_t)B * P.runs_per_clip;
    const unsigned grid = (unsigned)((runs + ddsp::kCsbWarps - 1) / ddsp::kCsbWarps);
    if (noise_u) ddsp::combsubfast_backward_kernel<true><<<grid, ddsp::kCsbThreads, ddsp::kCsbSmemBytes, st>>>(PB);
    else ddsp::combsubfast_backward_kernel<false><<<grid, ddsp::kCsbThreads, ddsp::kCsbSmemBytes, st>>>(PB);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// frequency_filter and the two synthesizers built on it
// ------------------------------------------------------------------------------------------------
namespace {

size_t ltv_spec_bytes(int B, int F) { return (size_t)B * F * ddsp::kLtvSpecFloat2 * sizeof(float2); }

// Fill the parameter block shared by the IR and the convolution kernels of one filter.
int ltv_params(ddsp::LtvParams& P, const float* audio, int audio_mode, uint64_t seed, const float* mags, int64_t mB,
               int64_t mF, int n_mag, int encoding, float mag_scale, int window_mode, const float* f0_frames,
               int64_t fB, int64_t fF, double sr, int B, int F, float* out, void* spec_ws, cudaStream_t st) {
    if (n_mag != 256 && n_mag != 512) return DDSP_B200_ERR_UNSUPPORTED;
    if (n_mag == 512 && (encoding == DDSP_B200_MAG_ALLPASS_TANH || encoding == DDSP_B200_MAG_COMPLEX))
        return DDSP_B200_ERR_UNSUPPORTED;     // the L=1022 path assumes real magnitudes (symmetric IR)
    if (window_mode == DDSP_B200_WINDOW_DYNAMIC && !f0_frames) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (audio_mode != 2 && (!audio || ((uintptr_t)audio & 7))) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (!spec_ws || ((uintptr_t)spec_ws & 15)) return DDSP_B200_ERR_WORKSPACE;
    if (int rc = ensure_device_ready(st, &P.tw_tables)) return rc;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    P.chirp = g_chirp_ptr[dev][n_mag == 256 ? 0 : 1];
    P.audio = audio; P.audio_mode = audio_mode; P.seed = seed;
    P.key_offset = (uint32_t)((uint64_t)g_noise_hop_offset * 32ull) * 0x9E3779B1u;     // same folding as combsubfast_impl
    P.mags = mags; P.mB = mB; P.mF = mF; P.n_mag = n_mag; P.encoding = encoding; P.mag_scale = mag_scale;
    P.window_mode = window_mode; P.f0_frames = f0_frames; P.fB = fB; P.fF = fF; P.sr15 = (float)(1.5 * sr);
    P.spec = (float2*)spec_ws; P.out = out; P.add_in = nullptr; P.sum_out = nullptr; P.B = B; P.F = F;
    const int frames = F + 1;
    const int64_t slots = (int64_t)sm_count() * ddsp::kLtvWarps;
    // one resident wave: as many runs per clip as the chip's warp slots allow, each at least 4 frames long (keeps the
    // 3-hop seams a minority of the work), the first run_rem of them one frame longer than the rest
    int64_t R = slots / B;
    if (R > frames / 4) R = frames / 4;
    if (R < 1) R = 1;
    P.runs_per_clip = (int)R;
    P.run_len = frames / (int)R;
    P.run_rem = frames % (int)R;
    return DDSP_B200_OK;
}

unsigned ltv_ir_grid(int B, int F) {
    int64_t grid = ((int64_t)B * F + ddsp::kLtvWarps - 1) / ddsp::kLtvWarps;
    if (grid > sm_count()) grid = sm_count();               // persistent: one CTA per SM, warps stride over frames
    return (unsigned)grid;
}

// 1) impulse responses -> tap spectra (frames independent)
int launch_ltv_ir(const ddsp::LtvParams& P, cudaStream_t st) {
    ltv_ir_select(P.encoding, P.window_mode, P.n_mag)<<<ltv_ir_grid(P.B, P.F), ddsp::kLtvThreads, ddsp::kLtvIrSmemBytes, st>>>(P);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

// 2) framing + convolution + overlap-add
int launch_ltv_conv(const ddsp::LtvParams& P, cudaStream_t st) {
    if (P.runs_per_clip > 1) {
        // only the hops at run seams are accumulated with atomics (2 hops for L = 510, 3 for L = 1022)
        const int n_seams = P.B * (P.runs_per_clip - 1), hops = P.n_mag == 256 ? 2 : 3;
        ddsp::ltv_zero_seams_kernel<<<(unsigned)(n_seams * hops), 128, 0, st>>>(P.out, P.F, P.run_len, P.run_rem, P.runs_per_clip,
                                                                               P.n_mag - 1, hops, n_seams);
        LAUNCH_CHECK();
    }
    const int64_t runs = (int64_t)P.B * P.runs_per_clip;
    const unsigned grid = (unsigned)((runs + ddsp::kLtvWarps - 1) / ddsp::kLtvWarps);
    ltv_conv_select(P.audio_mode, P.n_mag)<<<grid, ddsp::kLtvThreads, ddsp::kLtvConvSmemBytes, st>>>(P);
    LAUNCH_CHECK();
    if (P.sum_out && P.runs_per_clip > 1) {
        const int n_seams = P.B * (P.runs_per_clip - 1);
        ddsp::ltv_sum_seams_kernel<<<(unsigned)(n_seams * 2), 128, 0, st>>>(P.out, P.add_in, P.sum_out, P.F, P.run_len, P.run_rem,
                                                                           P.runs_per_clip, P.n_mag - 1, 2, n_seams);
        LAUNCH_CHECK();
    }
    return DDSP_B200_OK;
}

int launch_ltv(const float* audio, int audio_mode, uint64_t seed, const float* mags, int64_t mB, int64_t mF, int n_mag,
               int encoding, float mag_scale, int window_mode, const float* f0_frames, int64_t fB, int64_t fF,
               double sr, int B, int F, float* out, void* spec_ws, cudaStream_t st) {
    ddsp::LtvParams P;
    if (int rc = ltv_params(P, audio, audio_mode, seed, mags, mB, mF, n_mag, encoding, mag_scale, window_mode, f0_frames,
                            fB, fF, sr, B, F, out, spec_ws, st)) return rc;
    if (int rc = launch_ltv_ir(P, st)) return rc;
    return launch_ltv_conv(P, st);
}

// All-pass (group delay) + noise impulse responses of a frame from one shared Bluestein pass (both L=510),
// followed by the two convolutions.  `between` (optional) runs after the all-pass convolution: CombSub's
// harmonic filter consumes its output.
int launch_allpass_and_noise(const float* allpass_in, float* allpass_out, const float* group_delay,
                             const float* noise_u, uint64_t seed, const float* noise_magnitude, int64_t cB, int64_t cF,
                             double sr, int B, int F, float* noise_out, void* spec_a, void* spec_n, cudaStream_t st,
                             int (*between)(void*), void* between_arg, const float* sum_in, float* sum_out) {
    ddsp::LtvParams Pa, Pn;
    if (int rc = ltv_params(Pa, allpass_in, 0, 0, group_delay, cB, cF, 256, DDSP_B200_MAG_ALLPASS_TANH, 1.0f,
                            DDSP_B200_WINDOW_NONE, nullptr, 0, 0, sr, B, F, allpass_out, spec_a, st)) return rc;
    if (int rc = ltv_params(Pn, noise_u, noise_u ? 1 : 2, seed, noise_magnitude, cB, cF, 256, DDSP_B200_MAG_EXP,
                            1.0f / 128.0f, DDSP_B200_WINDOW_HANN, nullptr, 0, 0, sr, B, F, noise_out, spec_n, st)) return rc;
    // the noise convolution runs last and writes signal = sum_in + noise beside the noise itself (vocoder.py:421,548)
    Pn.add_in = sum_in; Pn.sum_out = sum_out;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    ddsp::LtvDualParams D;
    D.gd = group_delay; D.nm = noise_magnitude; D.mB = cB; D.mF = cF; D.noise_scale = 1.0f / 128.0f;
    D.tw_tables = Pa.tw_tables; D.chirp_c = g_chirp_ptr[dev][0]; D.chirp_d_dual = g_chirp_ptr[dev][2] + 1024;
    D.spec_a = (float2*)spec_a; D.spec_n = (float2*)spec_n; D.B = B; D.F = F;
    ddsp::ltv_ir_dual_kernel<<<ltv_ir_grid(B, F), ddsp::kLtvThreads, ddsp::kLtvDualSmemBytes, st>>>(D);
    LAUNCH_CHECK();
    int launches = g_launches; g_launches = 0;
    if (int rc = launch_ltv_conv(Pa, st)) return rc;
    launches += g_launches; g_launches = 0;
    if (between) {
        if (int rc = between(between_arg)) return rc;
        launches += g_launches; g_launches = 0;
    }
    if (int rc = launch_ltv_conv(Pn, st)) return rc;
    g_launches += launches;
    return DDSP_B200_OK;
}

int launch_add(const float* a, const float* b, float* out, int64_t n, cudaStream_t st) {
    if (n % 4) return DDSP_B200_ERR_INVALID_ARGUMENT;
    ddsp::add_kernel<<<(unsigned)grid_for(n / 4, 256, 148 * 16), 256, 0, st>>>(
        reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), reinterpret_cast<float4*>(out), n / 4);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

}  // namespace

extern "C" {

int ddsp_b200_apply_frame_mask(float* signal, const float* mask_frames, int64_t mB, int64_t mF, int B, int F, int hop,
                               void* stream) {
    g_launches = 0;
    if (!signal || !mask_frames || B <= 0 || F <= 0 || ((uintptr_t)signal & 15)) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (hop != ddsp::kHop) return DDSP_B200_ERR_UNSUPPORTED;
    const int64_t n4 = (int64_t)B * F * (hop / 4);
    ddsp::apply_frame_mask_kernel<<<(unsigned)grid_for(n4, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(signal), mask_frames, mB, mF, B, F);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_apply_volume_mask(float* signal, const float* volume_frames, int64_t vB, int64_t vF, double threshold, int B,
                                int F, int hop, void* stream) {
    g_launches = 0;
    if (!signal || !volume_frames || B <= 0 || F <= 0 || ((uintptr_t)signal & 15)) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (hop != ddsp::kHop) return DDSP_B200_ERR_UNSUPPORTED;
    const int64_t n4 = (int64_t)B * F * (hop / 4);
    ddsp::apply_volume_mask_kernel<<<(unsigned)grid_for(n4, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(signal), volume_frames, vB, vF, threshold, B, F);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_performer_features(const float* dash, const float* x, int B, int N, int H, int M, int is_query, float eps,
                                 float* out, void* stream) {
    g_launches = 0;
    if (!dash || !x || !out || B <= 0 || N <= 0 || H <= 0 || M <= 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (M > 384 || ((uintptr_t)x & 7)) return DDSP_B200_ERR_UNSUPPORTED;
    const int64_t rows = (int64_t)B * N * H;
    const unsigned grid = (unsigned)grid_for(rows, 8, (int64_t)sm_count() * 32);
    const float ratio = 1.0f / sqrtf((float)M);
    if (is_query)
        ddsp::performer_features_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(dash, x, out, B, N, H, M, 0.0625f,
                                                                                     ratio, eps);
    else
        ddsp::performer_features_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(dash, x, out, B, N, H, M, 0.0625f,
                                                                                      ratio, eps);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_performer_project_features(const float* x, const float* x_bias, const float* projection, int B, int N,
                                         int H, int M, int is_query, float eps, float* out, void* stream) {
    g_launches = 0;
    if (!x || !projection || !out || B <= 0 || N <= 0 || H <= 0 || M <= 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (M < 256 || M > ddsp::kPpfCols || ((uintptr_t)x & 15) || ((uintptr_t)x_bias & 15)) return DDSP_B200_ERR_UNSUPPORTED;
    const float* tables = nullptr;
    if (int rc = ensure_device_ready((cudaStream_t)stream, &tables, 1)) return rc;     // shared-memory opt-in
    const int64_t groups = ((int64_t)B * N * H + ddsp::kPpfRows - 1) / ddsp::kPpfRows;
    const unsigned grid = (unsigned)grid_for(groups, ddsp::kPpfWarps, (int64_t)sm_count() * 2);
    const float normalizer = 0.35355339059327373f;             // 64^-0.25 (pcmer.py:137)
    const float ratio = 1.0f / sqrtf((float)M);
    if (is_query)
        ddsp::performer_project_features_kernel<true><<<grid, ddsp::kPpfWarps * 32, ddsp::kPpfSmemBytes, (cudaStream_t)stream>>>(
            x, x_bias, projection, out, B, N, H, M, normalizer, 0.0625f, ratio, eps);
    else
        ddsp::performer_project_features_kernel<false><<<grid, ddsp::kPpfWarps * 32, ddsp::kPpfSmemBytes, (cudaStream_t)stream>>>(
            x, x_bias, projection, out, B, N, H, M, normalizer, 0.0625f, ratio, eps);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

size_t ddsp_b200_performer_attention_workspace_bytes(int B, int N, int H) {
    if (B <= 0 || N <= 0 || H <= 0) return 0;
    const size_t tiles = (size_t)(N + ddsp::kPasRows - 1) / ddsp::kPasRows;
    return (size_t)B * H * (tiles + 1) * ddsp::kPctxFloats * sizeof(float);
}

int ddsp_b200_performer_attention(const float* q, const float* k, const float* v, int64_t row_stride,
                                  const float* q_bias, const float* k_bias, const float* v_bias, const float* projection,
                                  int B, int N, int H, int M, float eps, float* out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    g_launches = 0;
    if (!q || !k || !v || !projection || !out || B <= 0 || N <= 0 || H <= 0 || M <= 0 || row_stride < (int64_t)H * 64 ||
        (row_stride & 3) || ((uintptr_t)q & 15) || ((uintptr_t)k & 15) || ((uintptr_t)v & 15))
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    const int64_t rs = row_stride;
    if (M > ddsp::kPpfCols || B > 65535) return DDSP_B200_ERR_UNSUPPORTED;
    const float* tables = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = ensure_device_ready(st, &tables, 1)) return rc;     // shared-memory opt-in
    const float ratio = 1.0f / sqrtf((float)M);
    if (N <= 2 * ddsp::kPasRows) {
        ddsp::performer_attention_small_kernel<<<dim3(H, B), ddsp::kPasThreads, ddsp::kPasSmemBytes, st>>>(
            q, k, v, q_bias, k_bias, v_bias, projection, out, N, H, M, rs, ratio, eps);
        LAUNCH_CHECK();
        return DDSP_B200_OK;
    }
    const int tiles = (N + ddsp::kPasRows - 1) / ddsp::kPasRows;
    if (tiles > 65535) return DDSP_B200_ERR_UNSUPPORTED;
    if (!workspace || ((uintptr_t)workspace & 15) || workspace_bytes < ddsp_b200_performer_attention_workspace_bytes(B, N, H))
        return DDSP_B200_ERR_WORKSPACE;
    float* partial = (float*)workspace;
    float* context = partial + (size_t)B * H * tiles * ddsp::kPctxFloats;
    ddsp::performer_context_partial_kernel<<<dim3(H, B, tiles), ddsp::kPasThreads, ddsp::kPasSmemBytes, st>>>(
        k, v, k_bias, v_bias, projection, partial, N, H, M, rs, ratio, eps);
    LAUNCH_CHECK();
    ddsp::performer_context_reduce_kernel<<<dim3((ddsp::kPctxFloats + 255) / 256, B * H), 256, 0, st>>>(partial, context, tiles);
    LAUNCH_CHECK();
    ddsp::performer_output_kernel<<<dim3(H, B, tiles), ddsp::kPasThreads, ddsp::kPasSmemBytes, st>>>(
        q, q_bias, projection, context, out, N, H, M, rs, ratio, eps);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_embed_sum(const float* x, int64_t xB, int64_t xN, int64_t xC, const float* f0, int64_t fB, int64_t fN,
                        const float* phase, int64_t pB, int64_t pN, const float* volume, int64_t vB, int64_t vN,
                        const float* w_f0, const float* b_f0, const float* w_phase, const float* b_phase,
                        const float* w_volume, const float* b_volume, const float* spk, int64_t sB, int B, int N, int C,
                        float* out, void* stream) {
    g_launches = 0;
    if (!x || !f0 || !phase || !volume || !w_f0 || !b_f0 || !w_phase || !b_phase || !w_volume || !b_volume || !spk ||
        !out || B <= 0 || N <= 0 || C <= 0)
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    const int64_t total = (int64_t)B * N * C;
    ddsp::embed_sum_kernel<<<(unsigned)grid_for(total, 256, (int64_t)sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(
        x, xB, xN, xC, f0, fB, fN, phase, pB, pN, volume, vB, vN, w_f0, b_f0, w_phase, b_phase, w_volume, b_volume, spk,
        sB, B, N, C, out);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_glu_dwconv_silu(const float* u, const float* u_bias, const float* weight, const float* bias, int B, int T,
                              int C, float* out, void* stream) {
    g_launches = 0;
    if (!u || !weight || !bias || !out || B <= 0 || T <= 0 || C <= 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (B > 65535) return DDSP_B200_ERR_UNSUPPORTED;
    const dim3 grid((C + ddsp::kDwTileC - 1) / ddsp::kDwTileC, (T + ddsp::kDwTileT - 1) / ddsp::kDwTileT, B);
    ddsp::glu_dwconv_silu_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(u, u_bias, weight, bias, out, B, T, C);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_dwconv_silu(const float* g, const float* weight, const float* bias, int B, int T, int C, float* out,
                          void* stream) {
    g_launches = 0;
    if (!g || !weight || !bias || !out || B <= 0 || T <= 0 || C <= 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (B > 65535) return DDSP_B200_ERR_UNSUPPORTED;
    const dim3 grid((C + 255) / 256, (T + ddsp::kDw2Run - 1) / ddsp::kDw2Run, B);
    if (C == 512) ddsp::dwconv_silu_kernel<512><<<grid, 256, 0, (cudaStream_t)stream>>>(g, weight, bias, out, T, C);
    else ddsp::dwconv_silu_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(g, weight, bias, out, T, C);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_pad_frames(const float* x, int64_t xB, int64_t xN, int B, int N, int C, float* out, void* stream) {
    g_launches = 0;
    if (!x || !out || B <= 0 || N <= 0 || C <= 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if ((C & 3) || (xB & 3) || (xN & 3) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
        return DDSP_B200_ERR_UNSUPPORTED;
    const int64_t total = (int64_t)B * (N + 2) * (C / 4);
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    ddsp::pad_frames_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, xB, xN, B, N, C, reinterpret_cast<float4*>(out));
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_groupnorm_leaky(float* hp, const float* gamma, const float* beta, float eps, float slope, int groups, int B,
                              int N, int C, double* sums, void* stream) {
    g_launches = 0;
    if (!hp || !gamma || !beta || !sums || B <= 0 || N <= 0 || C <= 0 || groups <= 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (groups > 32 || C % groups || ((C / groups) & 3) || C / 4 > ddsp::kGnThreads || ddsp::kGnThreads % (C / 4) || B > 65535 ||
        (reinterpret_cast<uintptr_t>(hp) & 15) || (reinterpret_cast<uintptr_t>(gamma) & 15) || (reinterpret_cast<uintptr_t>(beta) & 15))
        return DDSP_B200_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * groups * B, st));
    const int chunks = std::max(1, std::min(32, (sm_count() * 4 + B - 1) / B));
    ddsp::groupnorm_stats_kernel<<<dim3(chunks, B), ddsp::kGnThreads, 0, st>>>(hp, N, C, groups, sums);
    LAUNCH_CHECK();
    const int64_t total = (int64_t)B * (N + 2) * (C / 4);
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    ddsp::groupnorm_leaky_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<float4*>(hp), sums, gamma, beta, eps, slope, B, N, C, groups);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_embed_sum_ln(const float* x, int64_t xB, int64_t xN, const float* f0, int64_t fB, int64_t fN, const float* phase,
                           int64_t pB, int64_t pN, const float* volume, int64_t vB, int64_t vN, const float* w_f0,
                           const float* b_f0, const float* w_phase, const float* b_phase, const float* w_volume,
                           const float* b_volume, const float* spk, int64_t sB, const float* ln_gamma, const float* ln_beta,
                           float ln_eps, int B, int N, int C, float* out, float* out_ln, void* stream) {
    g_launches = 0;
    if (!x || !f0 || !phase || !volume || !w_f0 || !b_f0 || !w_phase || !b_phase || !w_volume || !b_volume || !spk || !ln_gamma ||
        !ln_beta || !out || !out_ln || B <= 0 || N <= 0)
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (C != 256 || (xB & 3) || (xN & 3) || (sB & 3)) return DDSP_B200_ERR_UNSUPPORTED;
    for (const void* p : {(const void*)x, (const void*)w_f0, (const void*)b_f0, (const void*)w_phase, (const void*)b_phase,
                          (const void*)w_volume, (const void*)b_volume, (const void*)spk, (const void*)ln_gamma,
                          (const void*)ln_beta, (const void*)out, (const void*)out_ln})
        if (reinterpret_cast<uintptr_t>(p) & 15) return DDSP_B200_ERR_UNSUPPORTED;
    const int64_t rows = (int64_t)B * N;
    const int blocks = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sm_count() * 8);
    ddsp::embed_sum_ln256_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, xB, xN, f0, fB, fN, phase, pB, pN, volume, vB, vN,
                                                                          w_f0, b_f0, w_phase, b_phase, w_volume, b_volume, spk,
                                                                          sB, ln_gamma, ln_beta, ln_eps, B, N, out, out_ln);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

size_t ddsp_b200_frequency_filter_workspace_bytes(int B, int F, int n_mag) {
    (void)n_mag;
    if (B <= 0 || F <= 0) return 0;
    // tap spectra (B,F,1024) complex  +  a (B,T) scratch that is only touched when accumulate != 0
    return ltv_spec_bytes(B, F) + (size_t)B * F * ddsp::kHop * sizeof(float);
}

int ddsp_b200_frequency_filter(const float* audio, const float* mags, int64_t mB, int64_t mF, int n_mag,
                               int mag_encoding, float mag_scale, int window_mode, const float* f0_frames, int64_t fB,
                               int64_t fF, double sr, int B, int F, int hop, float* out, int accumulate,
                               void* workspace, size_t workspace_bytes, void* stream) {
    g_launches = 0;
    if (!audio || !mags || !out || B <= 0 || F <= 0 || !(sr > 0) || audio == out) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (hop != ddsp::kHop) return DDSP_B200_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < ddsp_b200_frequency_filter_workspace_bytes(B, F, n_mag))
        return DDSP_B200_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* tmp = (float*)((char*)workspace + ltv_spec_bytes(B, F));
    if (int rc = launch_ltv(audio, 0, 0, mags, mB, mF, n_mag, mag_encoding, mag_scale, window_mode, f0_frames, fB, fF, sr,
                            B, F, accumulate ? tmp : out, workspace, st))
        return rc;
    if (!accumulate) return DDSP_B200_OK;
    const int l = g_launches;
    if (int rc = launch_add(out, tmp, out, (int64_t)B * F * hop, st)) return rc;
    g_launches += l;
    return DDSP_B200_OK;
}

size_t ddsp_b200_combsub_workspace_bytes(int B, int F, int n_mag_allpass, int n_mag_harmonic, int n_mag_noise) {
    (void)n_mag_allpass; (void)n_mag_harmonic; (void)n_mag_noise;
    // combtooth + all-passed harmonic + tap spectra of the filter in flight
    return (B > 0 && F > 0) ? (size_t)2 * B * F * ddsp::kHop * sizeof(float) + 2 * ltv_spec_bytes(B, F) : 0;
}

int ddsp_b200_combsub(const float* group_delay, int n_mag_allpass, const float* harmonic_magnitude, int n_mag_harmonic,
                      const float* noise_magnitude, int n_mag_noise, int64_t cB, int64_t cF, const float* f0_frames,
                      int64_t fB, int64_t fF, const double* prefix, const float* initial_phase, const float* noise_u,
                      uint64_t seed, int B, int F, int hop, double sr, float* signal, float* harmonic, float* noise,
                      void* workspace, size_t workspace_bytes, void* stream) {
    g_launches = 0;
    (void)initial_phase;   // carried by `prefix`
    if (!group_delay || !harmonic_magnitude || !noise_magnitude || !f0_frames || !prefix || !signal || !harmonic ||
        !noise || B <= 0 || F <= 0 || !(sr > 0))
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (hop != ddsp::kHop) return DDSP_B200_ERR_UNSUPPORTED;
    if ((int64_t)F * hop >= (1ll << 24)) return DDSP_B200_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < ddsp_b200_combsub_workspace_bytes(B, F, n_mag_allpass, n_mag_harmonic, n_mag_noise))
        return DDSP_B200_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)B * F * hop;
    float* comb = (float*)workspace;
    float* h1 = comb + n;
    void* spec = (void*)(h1 + n);
    int launches = 0;
    // vocoder.py:539  combtooth (no unvoiced zeroing in the old CombSub)
    ddsp::combtooth_kernel<<<(unsigned)(((int64_t)B * F + 7) / 8), 256, 0, st>>>(f0_frames, fB, fF, B, F, 1.0 / sr,
                                                                                   (float)sr, prefix, 0, comb);
    LAUNCH_CHECK();
    launches += g_launches; g_launches = 0;
    void* spec_n = (void*)((char*)spec + ltv_spec_bytes(B, F));
    if (n_mag_allpass == 256 && n_mag_noise == 256) {
        // :540 all-pass and :545-546 noise impulse responses share one Bluestein pass; the harmonic
        // filter (:541-542, f0-dependent window) runs between the two convolutions on the all-pass output
        struct Mid { const float* h1; const float* hm; int n_mag; int64_t cB, cF; const float* f0; int64_t fB, fF;
                     double sr; int B, F; float* harmonic; void* spec; cudaStream_t st; } mid = {
            h1, harmonic_magnitude, n_mag_harmonic, cB, cF, f0_frames, fB, fF, sr, B, F, harmonic, spec, st};
        auto between = [](void* a) -> int {
            Mid* m = (Mid*)a;
            return launch_ltv(m->h1, 0, 0, m->hm, m->cB, m->cF, m->n_mag, DDSP_B200_MAG_EXP, 1.0f, DDSP_B200_WINDOW_DYNAMIC,
                              m->f0, m->fB, m->fF, m->sr, m->B, m->F, m->harmonic, m->spec, m->st);
        };
        if (int rc = launch_allpass_and_noise(comb, h1, group_delay, noise_u, seed, noise_magnitude, cB, cF, sr, B, F,
                                              noise, spec, spec_n, st, between, &mid, harmonic, signal)) return rc;
        launches += g_launches; g_launches = 0;
        g_launches += launches;
        return DDSP_B200_OK;
    } else {
        // :540  all-pass (group delay), no window
        if (int rc = launch_ltv(comb, 0, 0, group_delay, cB, cF, n_mag_allpass, DDSP_B200_MAG_ALLPASS_TANH, 1.0f,
                                DDSP_B200_WINDOW_NONE, nullptr, 0, 0, sr, B, F, h1, spec, st)) return rc;
        launches += g_launches; g_launches = 0;
        // :541-542  harmonic magnitude filter with the f0-dependent window
        if (int rc = launch_ltv(h1, 0, 0, harmonic_magnitude, cB, cF, n_mag_harmonic, DDSP_B200_MAG_EXP, 1.0f,
                                DDSP_B200_WINDOW_DYNAMIC, f0_frames, fB, fF, sr, B, F, harmonic, spec, st)) return rc;
        launches += g_launches; g_launches = 0;
        // :545-546  filtered noise
        if (int rc = launch_ltv(noise_u, noise_u ? 1 : 2, seed, noise_magnitude, cB, cF, n_mag_noise, DDSP_B200_MAG_EXP,
                                1.0f / 128.0f, DDSP_B200_WINDOW_HANN, nullptr, 0, 0, sr, B, F, noise, spec, st)) return rc;
        launches += g_launches; g_launches = 0;
    }
    // :548
    if (int rc = launch_add(harmonic, noise, signal, n, st)) return rc;
    g_launches += launches;
    return DDSP_B200_OK;
}

size_t ddsp_b200_sins_workspace_bytes(int B, int F, int n_harmonics, int n_mag_allpass, int n_mag_noise) {
    (void)n_harmonics; (void)n_mag_allpass; (void)n_mag_noise;
    return (B > 0 && F > 0) ? (size_t)B * F * ddsp::kHop * sizeof(float) + 2 * ltv_spec_bytes(B, F) : 0;   // sinusoid mix + 2 tap spectra
}

int ddsp_b200_sins(const float* amplitudes, int n_harmonics, const float* group_delay, int n_mag_allpass,
                   const float* noise_magnitude, int n_mag_noise, int64_t cB, int64_t cF, const float* f0_frames,
                   int64_t fB, int64_t fF, const float* phase_full, const float* noise_u, uint64_t seed, int B, int F,
                   int hop, double sr, float* signal, float* harmonic, float* noise, void* workspace,
                   size_t workspace_bytes, void* stream) {
    g_launches = 0;
    if (!amplitudes || !group_delay || !noise_magnitude || !f0_frames || !phase_full || !signal || !harmonic || !noise ||
        B <= 0 || F <= 0 || !(sr > 0) || n_harmonics <= 0)
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (hop != ddsp::kHop || n_harmonics > ddsp::kSinsMaxHarm || (n_harmonics & 1)) return DDSP_B200_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < ddsp_b200_sins_workspace_bytes(B, F, n_harmonics, n_mag_allpass, n_mag_noise))
        return DDSP_B200_ERR_WORKSPACE;
    if ((int64_t)B * F > 0x7fffffffLL) return DDSP_B200_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)B * F * hop;
    float* sinus = (float*)workspace;
    void* spec = (void*)(sinus + n);
    int launches = 0;
    // vocoder.py:397,402-412  oscillator bank with the Nyquist mask (fmax = sr/2)
    ddsp::sins_osc_kernel<<<(unsigned)((int64_t)B * F), ddsp::kOscThreads, 0, st>>>(amplitudes, cB, cF, n_harmonics, f0_frames, fB, fF,
                                                                       F, (float)(sr / 2.0), phase_full, sinus);
    LAUNCH_CHECK();
    launches += g_launches; g_launches = 0;
    void* spec_n = (void*)((char*)spec + ltv_spec_bytes(B, F));
    if (n_mag_allpass == 256 && n_mag_noise == 256) {
        // :415 all-pass and :418-419 noise: impulse responses from one shared Bluestein pass
        if (int rc = launch_allpass_and_noise(sinus, harmonic, group_delay, noise_u, seed, noise_magnitude, cB, cF, sr, B,
                                              F, noise, spec, spec_n, st, nullptr, nullptr, harmonic, signal)) return rc;
        launches += g_launches; g_launches = 0;
        g_launches += launches;
        return DDSP_B200_OK;
    } else {
        // :415  all-pass
        if (int rc = launch_ltv(sinus, 0, 0, group_delay, cB, cF, n_mag_allpass, DDSP_B200_MAG_ALLPASS_TANH, 1.0f,
                                DDSP_B200_WINDOW_NONE, nullptr, 0, 0, sr, B, F, harmonic, spec, st)) return rc;
        launches += g_launches; g_launches = 0;
        // :418-419  filtered noise
        if (int rc = launch_ltv(noise_u, noise_u ? 1 : 2, seed, noise_magnitude, cB, cF, n_mag_noise, DDSP_B200_MAG_EXP,
                                1.0f / 128.0f, DDSP_B200_WINDOW_HANN, nullptr, 0, 0, sr, B, F, noise, spec, st)) return rc;
        launches += g_launches; g_launches = 0;
    }
    // :421
    if (int rc = launch_add(harmonic, noise, signal, n, st)) return rc;
    g_launches += launches;
    return DDSP_B200_OK;
}

int ddsp_b200_combsub_stream(const float* group_delay, int n_mag_allpass, const float* harmonic_magnitude, int n_mag_harmonic,
                             const float* noise_magnitude, int n_mag_noise, int64_t cB, int64_t cF, const float* f0_frames,
                             int64_t fB, int64_t fF, const double* prefix, const float* noise_u, uint64_t seed,
                             int64_t hop_offset, int B, int F, int hop, double sr, float* signal, float* harmonic,
                             float* noise, void* workspace, size_t workspace_bytes, void* stream) {
    if (hop_offset < 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    g_noise_hop_offset = hop_offset;
    const int rc = ddsp_b200_combsub(group_delay, n_mag_allpass, harmonic_magnitude, n_mag_harmonic, noise_magnitude, n_mag_noise,
                                     cB, cF, f0_frames, fB, fF, prefix, nullptr, noise_u, seed, B, F, hop, sr, signal, harmonic,
                                     noise, workspace, workspace_bytes, stream);
    g_noise_hop_offset = 0;
    return rc;
}

int ddsp_b200_sins_stream(const float* amplitudes, int n_harmonics, const float* group_delay, int n_mag_allpass,
                          const float* noise_magnitude, int n_mag_noise, int64_t cB, int64_t cF, const float* f0_frames,
                          int64_t fB, int64_t fF, const float* phase_full, const float* noise_u, uint64_t seed,
                          int64_t hop_offset, int B, int F, int hop, double sr, float* signal, float* harmonic, float* noise,
                          void* workspace, size_t workspace_bytes, void* stream) {
    if (hop_offset < 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    g_noise_hop_offset = hop_offset;
    const int rc = ddsp_b200_sins(amplitudes, n_harmonics, group_delay, n_mag_allpass, noise_magnitude, n_mag_noise, cB, cF,
                                  f0_frames, fB, fF, phase_full, noise_u, seed, B, F, hop, sr, signal, harmonic, noise,
                                  workspace, workspace_bytes, stream);
    g_noise_hop_offset = 0;
    return rc;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Linear layers of the control network on the tensor cores (csrc/gemm_tc.cuh)
// ------------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point query (no link-time dependency on libcuda)
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// (rows, cols) fp32 matrix with row stride ld (elements), cols contiguous; box = 32 columns (128 bytes) x box_rows,
// 128-byte swizzle, out-of-bounds elements read as zero
int make_map_2d(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return DDSP_B200_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)ddsp::tc::kBK, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_last_cuda_error = (int)r;
        return DDSP_B200_ERR_CUDA;
    }
    return 0;
}

template <int BN>
int launch_linear(const float* A, int64_t lda, const float* W, int64_t ldw, ddsp::tc::LinearParams P, cudaStream_t st) {
    using C = ddsp::tc::Cfg<BN>;
    static bool attr_set[64] = {false};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DDSP_B200_ERR_UNSUPPORTED;
    if (!attr_set[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(ddsp::tc::linear_tf32x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      C::kSmemBytes));
        attr_set[dev] = true;
    }
    CUtensorMap map_a, map_w;
    if (int rc = make_map_2d(&map_a, A, P.M, P.K, lda, ddsp::tc::kBM)) return rc;
    if (int rc = make_map_2d(&map_w, W, P.N, P.K, ldw, BN)) return rc;
    P.tiles_m = (P.M + ddsp::tc::kBM - 1) / ddsp::tc::kBM;
    P.tiles_n = (P.N + BN - 1) / BN;
    const int64_t tiles = P.virtual_tiles > 0 ? P.virtual_tiles : (int64_t)P.tiles_m * P.tiles_n;
    const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
    ddsp::tc::linear_tf32x3_kernel<BN><<<grid, ddsp::tc::kThreads, C::kSmemBytes, st>>>(map_a, map_w, P);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int linear_dispatch(const float* A, int64_t lda, const float* W, int64_t ldw, ddsp::tc::LinearParams P, int block_n,
                    cudaStream_t st) {
    if (block_n == 0) {
        // least padded work; ties go to the wider tile (fewer re-reads of A)
        int best = 256;
        int64_t best_cost = ((P.N + 255) / 256) * 256;
        for (int bn : {224, 128, 64}) {
            const int64_t cost = (int64_t)((P.N + bn - 1) / bn) * bn;
            if (cost < best_cost) { best_cost = cost; best = bn; }
        }
        block_n = best;
    }
    switch (block_n) {
        case 256: return launch_linear<256>(A, lda, W, ldw, P, st);
        case 224: return launch_linear<224>(A, lda, W, ldw, P, st);
        case 128: return launch_linear<128>(A, lda, W, ldw, P, st);
        case 64: return launch_linear<64>(A, lda, W, ldw, P, st);
        default: return DDSP_B200_ERR_UNSUPPORTED;
    }
}

}  // namespace

extern "C" {

int ddsp_b200_linear_tf32x3(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                            const float* residual, int64_t ldr, float* C, int64_t ldc, int M, int N, int K, void* stream) {
    g_launches = 0;
    if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0 || lda < K || ldw < K || ldc < N || (residual && ldr < N))
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    // TMA: 16-byte aligned base addresses and row strides
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) || (lda & 3) || (ldw & 3))
        return DDSP_B200_ERR_UNSUPPORTED;
    ddsp::tc::LinearParams P = {};
    P.bias = bias; P.residual = residual; P.C = C; P.ldr = ldr; P.ldc = ldc;
    P.M = M; P.N = N; P.K = K; P.store_output = 1; P.virtual_tiles = 0;
    return linear_dispatch(A, lda, W, ldw, P, 0, (cudaStream_t)stream);
}

int ddsp_b200_tc_microbench(const float* A, const float* W, float* C, int N, int K, int block_n, int virtual_tiles,
                            void* stream) {
    g_launches = 0;
    if (!A || !W || !C || N <= 0 || K <= 0 || virtual_tiles <= 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    ddsp::tc::LinearParams P = {};
    P.C = C; P.ldc = N; P.M = ddsp::tc::kBM; P.N = N; P.K = K; P.store_output = 0; P.virtual_tiles = virtual_tiles;
    return linear_dispatch(A, K, W, K, P, block_n, (cudaStream_t)stream);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Second-generation tensor-core path of the control network (csrc/gemm_attn.cuh): batched 3xTF32 GEMM with fused
// epilogues (bias / residual / LayerNorm, head-split q|k|v, attention normalisation) and the FAVOR+ feature GEMM
// ------------------------------------------------------------------------------------------------
namespace {

// fp32 tensor of rank 3 or 4 (innermost dimension contiguous), box = 32 floats (one 128-byte swizzle row) x box_rows x 1 (x 1)
int make_map_nd(CUtensorMap* m, const float* base, int rank, const int64_t* dims, const int64_t* strides_elems, int box_rows,
                int box_cols = ddsp::tc::kBK) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return DDSP_B200_ERR_UNSUPPORTED;
    cuuint64_t d[4];
    cuuint64_t sb[3];
    cuuint32_t box[4] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1, 1};      // 32 columns: 128-byte swizzle rows, 16: 64-byte
    cuuint32_t es[4] = {1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) d[i] = (cuuint64_t)dims[i];
    for (int i = 0; i + 1 < rank; ++i) {
        if (strides_elems[i] & 3) return DDSP_B200_ERR_UNSUPPORTED;
        sb[i] = (cuuint64_t)strides_elems[i] * sizeof(float);
    }
    if (reinterpret_cast<uintptr_t>(base) & 15) return DDSP_B200_ERR_UNSUPPORTED;
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), d, sb, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_last_cuda_error = (int)r;
        return DDSP_B200_ERR_CUDA;
    }
    return 0;
}
// (rows x cols) matrices with row stride ld, Z of them batch_stride apart
int make_map_3(CUtensorMap* m, const float* base, int64_t cols, int64_t rows, int64_t Z, int64_t ld, int64_t batch_stride,
               int box_rows, int box_cols = ddsp::tc::kBK) {
    const int64_t dims[3] = {cols, rows, Z};
    const int64_t strides[2] = {ld, Z > 1 ? batch_stride : ld * rows};
    return make_map_nd(m, base, 3, dims, strides, box_rows, box_cols);
}

template <typename K>
int set_smem_once(K kernel, int bytes, bool* flags) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DDSP_B200_ERR_UNSUPPORTED;
    if (!flags[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        flags[dev] = true;
    }
    return 0;
}

template <int BN, int EPI, int AM = ddsp::tc::kBM, int SLICES = 2, int BK = ddsp::tc::kBK, bool NSTACK = false>
int launch_gemm3x(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& wlo, const CUtensorMap& c, const CUtensorMap& c2,
                  ddsp::tc::GemmParams P, cudaStream_t st, const CUtensorMap* a_lo = nullptr) {
    using C = ddsp::tc::GCfg<BN, AM, SLICES, BK, NSTACK>;
    static bool attr_set[64] = {false};
    if (int rc = set_smem_once(ddsp::tc::gemm3x_kernel<BN, EPI, AM, SLICES, BK, NSTACK>, C::kSmemBytes, attr_set)) return rc;
    P.tiles_m = (P.M + ddsp::tc::kBM - 1) / ddsp::tc::kBM;
    P.tiles_n = (P.N + BN - 1) / BN;
    const int64_t tiles = (int64_t)P.Z * P.tiles_m * P.tiles_n;
    if (tiles > 0x7fffffffLL) return DDSP_B200_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
    static const int forced_pf = [] { const char* e = getenv("DDSP_B200_GEMM_PREFETCH"); return e ? atoi(e) : -1; }();   // experiments
    // the operands streamed once from HBM pass through a 2..3-stage ring: an L2 prefetch cursor 3 k-blocks ahead measured
    // 219 -> 207 us (context GEMM), 148.5 -> 142 us (output GEMM), 90.4 -> 88.6 us (N = 256 layers), 191 -> 185 us (final
    // projection); 6 / 12 k-blocks ahead are no better (profiles/r02_gemm_l2_prefetch.txt)
    P.l2_prefetch = 3;
    if (forced_pf >= 0) P.l2_prefetch = forced_pf;
    ddsp::tc::gemm3x_kernel<BN, EPI, AM, SLICES, BK, NSTACK><<<grid, ddsp::tc::kThreads, C::kSmemBytes, st>>>(a, a_lo ? *a_lo : a, w, wlo, c, c2, P);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

}  // namespace

extern "C" {

int ddsp_b200_linear_tf32x3_ex(const float* A, int64_t lda, const float* W, const float* W_lo, int64_t ldw, const float* bias,
                               const float* residual, int64_t ldr, float* C, int64_t ldc, const float* ln_gamma,
                               const float* ln_beta, float ln_eps, float* C_ln, int64_t ldc_ln, int M, int N, int K,
                               void* stream) {
    g_launches = 0;
    // lda < K with K a multiple of lda: row m is the window of K/lda consecutive lda-float rows (overlapping rows -- the
    // im2col view of a Conv1d over channels-last frames, unit2control.py:40-44)
    if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0 || lda <= 0 || (lda < K && K % lda) || ldw < K || ldc < N || (residual && ldr < N))
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    if ((ln_gamma != nullptr) != (C_ln != nullptr) || (ln_gamma && !ln_beta) || (ln_gamma && ldc_ln < N))
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (ln_gamma && (N > 256 || (N & 31))) return DDSP_B200_ERR_UNSUPPORTED;
    if (residual && ((N & 31) || (ldr & 3) || (reinterpret_cast<uintptr_t>(residual) & 15))) return DDSP_B200_ERR_UNSUPPORTED;
    if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return DDSP_B200_ERR_UNSUPPORTED;
    ddsp::tc::GemmParams P = {};
    P.Z = 1; P.M = M; P.N = N; P.K = K;
    P.w_presplit = W_lo ? 1 : 0;
    P.bias = bias; P.residual = residual; P.ldr = ldr;
    P.ln_gamma = ln_gamma; P.ln_beta = ln_beta; P.ln_eps = ln_eps;
    int bn = 256;
    static const int forced_bn = [] { const char* e = getenv("DDSP_B200_GEMM_BN"); return e ? atoi(e) : 0; }();   // experiments
    if (!ln_gamma && (forced_bn == 128 || forced_bn == 224 || forced_bn == 256)) bn = forced_bn;
    else if (!ln_gamma) {
        int64_t best_cost = ((N + 255) / 256) * 256;
        for (int cand : {224, 128}) {
            const int64_t cost = (int64_t)((N + cand - 1) / cand) * cand;
            if (cost < best_cost) { best_cost = cost; bn = cand; }
        }
    }
    static const int forced_bk = [] { const char* e = getenv("DDSP_B200_GEMM_BK"); return e ? atoi(e) : 0; }();   // experiments
    const int bk = (bn == 256 && forced_bk == 16) ? 16 : 32;
    CUtensorMap ma, mw, mwl, mc, mc2;
    if (int rc = make_map_3(&ma, A, K, M, 1, lda, 0, ddsp::tc::kBM, bk)) return rc;
    if (int rc = make_map_3(&mw, W, K, N, 1, ldw, 0, bn, bk)) return rc;
    if (int rc = make_map_3(&mwl, W_lo ? W_lo : W, K, N, 1, ldw, 0, bn, bk)) return rc;
    if (int rc = make_map_3(&mc, C, N, M, 1, ldc, 0, 32)) return rc;
    if (int rc = make_map_3(&mc2, C_ln ? C_ln : C, N, M, 1, C_ln ? ldc_ln : ldc, 0, 32)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (bn) {
        case 256:
            if (bk == 16) return launch_gemm3x<256, ddsp::tc::EPI_PLAIN, ddsp::tc::kBM, 2, 16>(ma, mw, mwl, mc, mc2, P, st);
            return launch_gemm3x<256, ddsp::tc::EPI_PLAIN>(ma, mw, mwl, mc, mc2, P, st);
        case 224: return launch_gemm3x<224, ddsp::tc::EPI_PLAIN>(ma, mw, mwl, mc, mc2, P, st);
        default: return launch_gemm3x<128, ddsp::tc::EPI_PLAIN>(ma, mw, mwl, mc, mc2, P, st);
    }
}

int ddsp_b200_linear_glu(const float* A, int64_t lda, const float* W, const float* W_lo, int64_t ldw, const float* bias,
                         float* C, int64_t ldc, int M, int N, int K, void* stream) {
    g_launches = 0;
    if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0 || lda < K || ldw < K || ldc < N / 2) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if ((N & 255) || (bias && (reinterpret_cast<uintptr_t>(bias) & 15))) return DDSP_B200_ERR_UNSUPPORTED;
    ddsp::tc::GemmParams P = {};
    P.Z = 1; P.M = M; P.N = N; P.K = K;
    P.w_presplit = W_lo ? 1 : 0;
    P.bias = bias;
    CUtensorMap ma, mw, mwl, mc;
    if (int rc = make_map_3(&ma, A, K, M, 1, lda, 0, ddsp::tc::kBM)) return rc;
    if (int rc = make_map_3(&mw, W, K, N, 1, ldw, 0, 256)) return rc;
    if (int rc = make_map_3(&mwl, W_lo ? W_lo : W, K, N, 1, ldw, 0, 256)) return rc;
    if (int rc = make_map_3(&mc, C, N / 2, M, 1, ldc, 0, 32)) return rc;
    return launch_gemm3x<256, ddsp::tc::EPI_GLU>(ma, mw, mwl, mc, mc, P, (cudaStream_t)stream);
}

int ddsp_b200_qkv_heads(const float* A, int64_t lda, const float* W, const float* W_lo, int64_t ldw, const float* bias,
                        float* q, float* k, float* vt, float* vt_lo, int B, int F, int Fp, int H, int K, void* stream) {
    g_launches = 0;
    if (!A || !W || !q || !k || !vt || B <= 0 || F <= 0 || H <= 0 || K <= 0 || lda < K || ldw < K || Fp < F)
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    if ((H & 3) || (Fp & 3) || (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) || (int64_t)B * F > 0x7fffffffLL)
        return DDSP_B200_ERR_UNSUPPORTED;
    ddsp::tc::GemmParams P = {};
    P.Z = 1; P.M = B * F; P.N = 3 * H * 64; P.K = K;
    P.w_presplit = W_lo ? 1 : 0;
    P.bias = bias; P.vt = vt; P.vt_lo = vt_lo; P.q = q; P.k = k; P.frames = F; P.frames_pad = Fp; P.heads = H;
    CUtensorMap ma, mw, mwl, mq, mk;
    if (int rc = make_map_3(&ma, A, K, P.M, 1, lda, 0, ddsp::tc::kBM)) return rc;
    if (int rc = make_map_3(&mw, W, K, P.N, 1, ldw, 0, 256)) return rc;
    if (int rc = make_map_3(&mwl, W_lo ? W_lo : W, K, P.N, 1, ldw, 0, 256)) return rc;
    const int64_t dims[4] = {64, F, H, B};
    const int64_t strides[3] = {64, (int64_t)F * 64, (int64_t)H * F * 64};
    if (int rc = make_map_nd(&mq, q, 4, dims, strides, 32)) return rc;
    if (int rc = make_map_nd(&mk, k, 4, dims, strides, 32)) return rc;
    return launch_gemm3x<256, ddsp::tc::EPI_QKV>(ma, mw, mwl, mq, mk, P, (cudaStream_t)stream);
}

int ddsp_b200_favor_features(const float* x, const float* proj_scaled, int n_features, int is_query, float eps, float* out,
                             int Z, int F, int Fp, void* stream) {
    g_launches = 0;
    if (!x || !proj_scaled || !out || Z <= 0 || F <= 0 || Fp < F) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (n_features != ddsp::tc::kFeat || (Fp & 3)) return DDSP_B200_ERR_UNSUPPORTED;
    ddsp::tc::FeatParams P = {};
    P.x = x; P.kt = out; P.Z = Z; P.F = F; P.Fp = Fp; P.eps = eps;
    P.tiles_m = (F + ddsp::tc::kBM - 1) / ddsp::tc::kBM;
    const int64_t tiles = (int64_t)Z * P.tiles_m;
    if (tiles > 0x7fffffffLL) return DDSP_B200_ERR_UNSUPPORTED;
    CUtensorMap ma, mw, mc;
    if (int rc = make_map_3(&ma, x, 64, F, Z, 64, (int64_t)F * 64, ddsp::tc::kBM)) return rc;
    if (int rc = make_map_3(&mw, proj_scaled, 64, n_features, 1, 64, 0, 136)) return rc;
    if (is_query) {
        if (int rc = make_map_3(&mc, out, ddsp::tc::kFeatPad, F, Z, ddsp::tc::kFeatPad, (int64_t)F * ddsp::tc::kFeatPad, 32)) return rc;
    } else {
        mc = ma;
    }
    const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_q[64] = {false}, attr_k[64] = {false};
    if (is_query) {
        if (int rc = set_smem_once(ddsp::tc::favor_features_kernel<true>, ddsp::tc::kFeatSmemBytes, attr_q)) return rc;
        ddsp::tc::favor_features_kernel<true><<<grid, ddsp::tc::feat_threads<true>(), ddsp::tc::kFeatSmemBytes, st>>>(ma, mw, mc, P);
    } else {
        if (int rc = set_smem_once(ddsp::tc::favor_features_kernel<false>, ddsp::tc::kFeatSmemBytes, attr_k)) return rc;
        ddsp::tc::favor_features_kernel<false><<<grid, ddsp::tc::feat_threads<false>(), ddsp::tc::kFeatSmemBytes, st>>>(ma, mw, mc, P);
    }
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_favor_context(const float* vt, const float* vt_lo, const float* kt, float* ctxT, float* ctxT_lo, int Z, int Fp,
                            void* stream) {
    g_launches = 0;
    if (!vt || !kt || !ctxT || Z <= 0 || Fp <= 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (Fp & 3) return DDSP_B200_ERR_UNSUPPORTED;
    using namespace ddsp::tc;
    GemmParams P = {};
    P.Z = Z; P.M = kVtRows; P.N = kFeatPad; P.K = Fp; P.w_batched = 1;
    CUtensorMap ma, mw, mc;
    if (int rc = make_map_3(&ma, vt, Fp, kVtRows, Z, Fp, (int64_t)kVtRows * Fp, kVtRows)) return rc;   // 80-row A tiles (GCfg: AM)
    if (int rc = make_map_3(&mw, kt, Fp, kFeatPad, Z, Fp, (int64_t)kFeatPad * Fp, 96)) return rc;
    if (int rc = make_map_3(&mc, ctxT, kFeatPad, kVtRows, Z, kFeatPad, (int64_t)kVtRows * kFeatPad, 32)) return rc;
    CUtensorMap mc2 = mc;
    if (ctxT_lo) {
        if (int rc = make_map_3(&mc2, ctxT_lo, kFeatPad, kVtRows, Z, kFeatPad, (int64_t)kVtRows * kFeatPad, 32)) return rc;
        P.split_out = 1;
    }
    CUtensorMap mal;
    if (vt_lo) {
        if (int rc = make_map_3(&mal, vt_lo, Fp, kVtRows, Z, Fp, (int64_t)kVtRows * Fp, kVtRows)) return rc;
        P.a_presplit = 1;
    }
    // measured and dropped: eight instead of four splitter warps for the two attention GEMMs (197.6 -> 197.4 us, 142.4 -> 140.7 us;
    // profiles/r02_gemm_splitw.txt); the output product transposed so that the frames are the wide N of the MMAs
    // (out^T = ctxT q'^T, 224 frames per tile, two 76-KB stages: 142 -> 210 us; profiles/r02_gemm_output_transposed.txt)
    // hi*hi and hi*lo(k') as one MMA of N = 192 over the adjacent k'_hi | k'_lo tiles: two instead of three reads of the A tile
    // per k-step through the shared-memory port, 197.9 -> 183.6 us (profiles/r02_gemm_nstack.txt)
    static const int nstack = [] { const char* e = getenv("DDSP_B200_ATTN_NSTACK"); return e ? atoi(e) : 1; }();   // experiments
    if (nstack)
        return launch_gemm3x<96, EPI_PLAIN, kVtRows, 2, kBK, true>(ma, mw, mw, mc, mc2, P, (cudaStream_t)stream, vt_lo ? &mal : nullptr);
    return launch_gemm3x<96, EPI_PLAIN, kVtRows>(ma, mw, mw, mc, mc2, P, (cudaStream_t)stream, vt_lo ? &mal : nullptr);   // 3 column tiles of 96 = 288 >= 272
}

int ddsp_b200_favor_output(const float* qf, const float* ctxT, const float* ctxT_lo, float* out, int B, int H, int F,
                           void* stream) {
    g_launches = 0;
    if (!qf || !ctxT || !out || B <= 0 || H <= 0 || F <= 0) return DDSP_B200_ERR_INVALID_ARGUMENT;
    using namespace ddsp::tc;
    const int Z = B * H;
    GemmParams P = {};
    P.Z = Z; P.M = F; P.N = kVtRows; P.K = kFeatPad; P.w_batched = 1; P.heads = H; P.frames = F;
    CUtensorMap ma, mw, mc;
    if (int rc = make_map_3(&ma, qf, kFeatPad, F, Z, kFeatPad, (int64_t)F * kFeatPad, kBM)) return rc;
    if (int rc = make_map_3(&mw, ctxT, kFeatPad, kVtRows, Z, kFeatPad, (int64_t)kVtRows * kFeatPad, kVtRows)) return rc;
    if (int rc = make_map_3(&mc, out, (int64_t)H * 64, F, B, (int64_t)H * 64, (int64_t)F * H * 64, 32)) return rc;
    CUtensorMap mwl = mw;
    if (ctxT_lo) {
        if (int rc = make_map_3(&mwl, ctxT_lo, kFeatPad, kVtRows, Z, kFeatPad, (int64_t)kVtRows * kFeatPad, kVtRows)) return rc;
        P.w_presplit = 1;
    }
    // (one staging slice per epilogue warp would buy a fourth stage here: measured 142 -> 150 us, not used)
    // (the stacked hi | lo form of the context GEMM measured here too: 142.0 -> 141.1 us, not used; profiles/r02_gemm_nstack.txt)
    static const int nstack = [] { const char* e = getenv("DDSP_B200_ATTN_NSTACK"); return e ? atoi(e) : 0; }();   // experiments
    if (nstack == 2) return launch_gemm3x<kVtRows, EPI_OUT, kBM, 2, kBK, true>(ma, mw, mwl, mc, mc, P, (cudaStream_t)stream);
    return launch_gemm3x<kVtRows, EPI_OUT>(ma, mw, mwl, mc, mc, P, (cudaStream_t)stream);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Downstream of the synthesizer: enhancer front-end and the GUI's SOLA splice (csrc/frontend.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" {

int ddsp_b200_mel_spectrogram(const float* audio, int B, int T, int n_fft, int win_size, int hop, const float* mel_basis,
                              const int* band_start, const int* band_end, int n_mels, float clip_val, float* out,
                              int n_frames, void* stream) {
    g_launches = 0;
    if (!audio || !mel_basis || !band_start || !band_end || !out || B <= 0 || T <= 0 || hop <= 0 || n_mels <= 0)
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (n_fft != ddsp::kMelFft || win_size != ddsp::kMelFft || hop > ddsp::kMelFft) return DDSP_B200_ERR_UNSUPPORTED;
    const int pad_left = (win_size - hop) / 2;
    int pad_right = (win_size - hop + 1) / 2;
    if (win_size - T - pad_left > pad_right) pad_right = win_size - T - pad_left;
    const int expect = 1 + (T + pad_left + pad_right - n_fft) / hop;
    if (n_frames != expect) return DDSP_B200_ERR_INVALID_ARGUMENT;
    ddsp::MelParams P;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = ensure_device_ready(st, &P.tw_tables, 1)) return rc;
    static bool attr_set[64] = {false};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(ddsp::mel_spectrogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ddsp::kMelSmemBytes));
        attr_set[dev] = true;
    }
    P.audio = audio; P.B = B; P.T = T; P.n_frames = n_frames; P.hop = hop; P.pad_left = pad_left;
    P.reflect = pad_right < T ? 1 : 0;
    P.mel_basis = mel_basis; P.band_start = band_start; P.band_end = band_end; P.n_mels = n_mels; P.clip_val = clip_val;
    P.out = out;
    int64_t grid = ((int64_t)B * n_frames + ddsp::kMelWarps - 1) / ddsp::kMelWarps;
    if (grid > 2 * sm_count()) grid = 2 * sm_count();
    ddsp::mel_spectrogram_kernel<<<(unsigned)grid, ddsp::kMelWarps * 32, ddsp::kMelSmemBytes, st>>>(P);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_sinc_resample(const float* x, int B, int T, const float* kernel_t, int orig, int nw, int width, float* y,
                            int T_out, void* stream) {
    g_launches = 0;
    if (!x || !kernel_t || !y || B <= 0 || T <= 0 || orig <= 0 || nw <= 0 || width < 0 || T_out <= 0)
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    const int K = 2 * width + orig;
    const size_t smem = (size_t)((ddsp::kResTileI - 1) * orig + K) * sizeof(float);
    if (smem > 48 * 1024 || B > 65535) return DDSP_B200_ERR_UNSUPPORTED;
    const int64_t n_i = ((int64_t)T_out + nw - 1) / nw;
    const dim3 grid((unsigned)((n_i + ddsp::kResTileI - 1) / ddsp::kResTileI), (unsigned)B);
    ddsp::sinc_resample_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, T, kernel_t, orig, nw, width, K, y, T_out);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_interp_frames(const float* f0, int64_t fB, int64_t fN, int B, int n, float scale, double hop_over_sr,
                            double real_factor, double dt_out, float* out, int n_out, void* stream) {
    g_launches = 0;
    if (!f0 || !out || B <= 0 || n <= 0 || n_out <= 0 || !(hop_over_sr > 0) || !(real_factor > 0) || !(dt_out > 0))
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    const int64_t tot = (int64_t)B * n_out;
    ddsp::interp_frames_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, (cudaStream_t)stream>>>(f0, fB, fN, B, n, scale, hop_over_sr,
                                                                                              real_factor, dt_out, out, n_out);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

int ddsp_b200_sola_splice(const float* x, int n, float* sola_buffer, const float* fade_in, const float* fade_out, int block,
                          int crossfade, int search, float* out, int* shift_out, void* stream) {
    g_launches = 0;
    if (!x || !sola_buffer || !fade_in || !fade_out || !out || !shift_out || block <= 0 || crossfade <= 0 || search < 0)
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (n < block + crossfade + search || block < crossfade) return DDSP_B200_ERR_INVALID_ARGUMENT;
    const size_t smem = (size_t)crossfade * sizeof(float) + ddsp::kSolaThreads * (sizeof(float) + sizeof(int));
    if (smem > 48 * 1024) return DDSP_B200_ERR_UNSUPPORTED;
    ddsp::sola_splice_kernel<<<1, ddsp::kSolaThreads, smem, (cudaStream_t)stream>>>(x, sola_buffer, fade_in, fade_out, block,
                                                                                   crossfade, search, out, shift_out);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

}  // extern "C"
