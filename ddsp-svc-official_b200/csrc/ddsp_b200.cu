// ddsp_b200.cu -- the C ABI (include/ddsp_b200.h): argument checks + kernel launches.
// No torch types, no allocation, no synchronisation; everything is enqueued on the caller's stream.
#include "../../include/ddsp_b200.h"

#include <cuda_runtime.h>

#include <mutex>

#include "combsubfast.cuh"
#include "phase.cuh"

namespace {

thread_local int g_last_cuda_error = 0;
thread_local int g_launches = 0;

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

// Immutable per-device tables (FFT twiddles + exact sqrt-Hann window), filled on first use.
__device__ __align__(16) float g_tables[ddsp::kTableBytes / 4];

std::mutex g_init_mutex;
bool g_device_ready[64] = {false};
const float* g_tables_ptr[64] = {nullptr};

// One-time per-device setup: opt-in shared memory sizes and the constant tables.  This is the
// only place the library synchronises (once per device, on the first call).
int ensure_device_ready(cudaStream_t st, const float** tables) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DDSP_B200_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> lock(g_init_mutex);
    if (!g_device_ready[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(ddsp::combsubfast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ddsp::kCsfSmemBytes));
        float* ptr = nullptr;
        CUDA_TRY(cudaGetSymbolAddress((void**)&ptr, g_tables));
        ddsp::fft_tables_kernel<<<4, 256, 0, st>>>(reinterpret_cast<float4*>(ptr), ptr + 2048);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(st));
        g_tables_ptr[dev] = ptr;
        g_device_ready[dev] = true;
    }
    *tables = g_tables_ptr[dev];
    return 0;
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

int ddsp_b200_phase(const float* f0_frames, int64_t fB, int64_t fF, int B, int F, int hop, double sr,
                    const float* initial_phase, int precise, float* phase_frames, double* prefix,
                    float* phase_full, void* stream) {
    g_launches = 0;
    (void)precise;   // the fused path always accumulates in fp64 (inference path, core.py:40)
    if (!f0_frames || !phase_frames || !prefix || B <= 0 || F <= 0 || !(sr > 0)) return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (hop != ddsp::kHop) return DDSP_B200_ERR_UNSUPPORTED;
    if ((int64_t)F * hop >= (1ll << 24)) return DDSP_B200_ERR_UNSUPPORTED;   // fp32-exact sample index (torch's own limit)
    cudaStream_t st = (cudaStream_t)stream;
    const double inv_sr = 1.0 / sr;
    const int64_t hops = (int64_t)B * F;
    // Small clips: one launch does totals + scan per clip.  Otherwise spread the totals over the chip.
    if (F <= 64 || (int64_t)B * 32 >= (int64_t)sm_count() * 64) {
        ddsp::phase_fused_kernel<<<B, 1024, 0, st>>>(f0_frames, fB, fF, F, inv_sr, initial_phase, prefix, phase_frames);
        LAUNCH_CHECK();
    } else {
        ddsp::hop_totals_kernel<<<(unsigned)((hops + 7) / 8), 256, 0, st>>>(f0_frames, fB, fF, B, F, prefix);
        LAUNCH_CHECK();
        ddsp::phase_scan_kernel<<<B, 1024, 0, st>>>(f0_frames, fB, fF, F, inv_sr, initial_phase, prefix, phase_frames);
        LAUNCH_CHECK();
    }
    if (phase_full) {
        ddsp::phase_full_kernel<<<(unsigned)((hops + 7) / 8), 256, 0, st>>>(f0_frames, fB, fF, B, F, inv_sr,
                                                                             initial_phase, prefix, phase_full);
        LAUNCH_CHECK();
    }
    return DDSP_B200_OK;
}

int ddsp_b200_combsubfast(const float* harmonic_magnitude, const float* harmonic_phase, const float* noise_magnitude,
                          int64_t cB, int64_t cF, const float* f0_frames, int64_t fB, int64_t fF,
                          const double* prefix, const float* initial_phase, const float* noise_u, uint64_t seed,
                          const float* window, int B, int F, int hop, double sr, float* signal, void* stream) {
    g_launches = 0;
    if (!harmonic_magnitude || !harmonic_phase || !noise_magnitude || !f0_frames || !prefix || !signal || B <= 0 ||
        F <= 0 || !(sr > 0))
        return DDSP_B200_ERR_INVALID_ARGUMENT;
    if (hop != ddsp::kHop) return DDSP_B200_ERR_UNSUPPORTED;
    if ((int64_t)F * hop >= (1ll << 24)) return DDSP_B200_ERR_UNSUPPORTED;
    ddsp::CsfParams P;
    if (int rc = ensure_device_ready((cudaStream_t)stream, &P.tables)) return rc;
    P.hm = harmonic_magnitude; P.hp = harmonic_phase; P.nm = noise_magnitude;
    P.cB = cB; P.cF = cF;
    P.f0_frames = f0_frames; P.fB = fB; P.fF = fF;
    P.prefix = prefix; (void)initial_phase; P.noise_u = noise_u; P.window = window;
    P.signal = signal; P.seed = seed; P.B = B; P.F = F;
    P.pairs_per_clip = (F + 2) / 2;                 // frames 0..F in pairs
    const int64_t slots = (int64_t)sm_count() * ddsp::kCsfWarps;
    const int64_t total_pairs = (int64_t)B * P.pairs_per_clip;
    int run_len = (int)((total_pairs + slots - 1) / slots);
    if (run_len < 1) run_len = 1;
    if (run_len > P.pairs_per_clip) run_len = P.pairs_per_clip;
    P.run_len = run_len;
    P.runs_per_clip = (P.pairs_per_clip + run_len - 1) / run_len;
    P.inv_sr = 1.0 / sr; P.sr = (float)sr;
    P.zero_unvoiced = 1;
    const int64_t runs = (int64_t)B * P.runs_per_clip;
    const unsigned grid = (unsigned)((runs + ddsp::kCsfWarps - 1) / ddsp::kCsfWarps);
    if (P.runs_per_clip > 1) {
        const int n_seams = B * (P.runs_per_clip - 1);
        ddsp::csf_zero_seams_kernel<<<n_seams, 128, 0, (cudaStream_t)stream>>>(signal, F, P.run_len, P.runs_per_clip,
                                                                                n_seams);
        LAUNCH_CHECK();
    }
    ddsp::combsubfast_kernel<<<grid, ddsp::kCsfThreads, ddsp::kCsfSmemBytes, (cudaStream_t)stream>>>(P);
    LAUNCH_CHECK();
    return DDSP_B200_OK;
}

}  // extern "C"
