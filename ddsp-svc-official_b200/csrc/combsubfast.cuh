// combsubfast.cuh -- stage B of CombSubFast.forward (ddsp/vocoder.py:455-492) as ONE kernel:
// excitation (combtooth sinc pulse train + uniform noise) -> sqrt-Hann framing -> 1024-point
// FFT -> per-bin filters exp(hm + j*pi*hp), exp(nm)/128 -> inverse FFT -> window -> overlap-add.
// Nothing but the control rows, f0 and the phase prefix is read from HBM and nothing but the
// finished signal is written.
//
// Work decomposition:
//   * a warp owns a run of consecutive frame PAIRS (2p, 2p+1) of one clip;
//   * frame m covers hops m-1 and m of the zero-padded excitation; each excitation hop is
//     generated once into a 2-slot shared-memory ring;
//   * per frame one complex FFT-1024 transforms comb (real part) and noise (imag part) together;
//     the two real spectra are separated with the conjugate-symmetric partner bin, which lives in
//     lane 32-l and is fetched with warp shuffles;
//   * per PAIR one complex inverse FFT-1024 returns both real output frames (V = Y_m + j*Y_{m+1});
//   * the hop shared by the two frames of a pair is finished in registers; the hop shared with
//     the next pair is carried in registers; at the start of a run one extra (halo) pair is
//     recomputed instead of synchronising with the neighbouring warp.
#pragma once
#include "fft32.cuh"
#include "phase.cuh"

namespace ddsp {

constexpr int kCsfWarps = 16;                       // warps per CTA (one CTA per SM)
constexpr int kCsfThreads = kCsfWarps * 32;
constexpr int kRingSlot = kHop + kHop / 32;         // 528 floats: one pad word per 32 samples
constexpr int kStashFloat2 = 17 * 32;               // Y_m stash: 16 bins/lane (+ bin 512 on lane 0)
constexpr int kCsfWarpBytes = kPlaneFloats * 4 + 2 * kRingSlot * 4 + kStashFloat2 * 8;
constexpr int kCsfSmemBytes = 1024 * 8 + 1024 * 4 + kCsfWarps * kCsfWarpBytes;

struct CsfParams {
    const float* hm; const float* hp; const float* nm;   // (B,F,513) views, strides (cB,cF,1)
    int64_t cB, cF;
    const float* f0_frames; int64_t fB, fF;              // (B,F)
    const double* prefix;                                 // (B,F) from stage A
    const float* initial_phase;                           // (B,) or null
    const float* noise_u;                                 // (B,T) or null
    const float* window;                                  // (1024,) module buffer or null
    float* signal;                                        // (B,T)
    uint64_t seed;
    int B, F;
    int pairs_per_clip, run_len, runs_per_clip;
    double inv_sr; float sr;
    int zero_unvoiced;                                    // CombSubFast: 1 (vocoder.py:460)
};

// Generate excitation hop h of clip b into a ring slot (zeros outside [0,F)).
__device__ __forceinline__ void csf_gen_hop(const CsfParams& P, int b, int h, float* __restrict__ slot,
                                            double init_rot, int lane) {
    if (h < 0 || h >= P.F) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { const int j = 16 * lane + i; slot[j + (j >> 5)] = 0.0f; }
        return;
    }
    const float* row = P.f0_frames + (int64_t)b * P.fB;
    const float x0 = __ldg(row + (int64_t)h * P.fF);
    const float x1 = __ldg(row + (int64_t)min(h + 1, P.F - 1) * P.fF);
    const double base = P.prefix[(int64_t)b * P.F + h];
    float f[16], rot[16];
    hop_rotation(x0, x1, base, P.inv_sr, init_rot, lane, f, rot);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        // vocoder.py:459  sinc(sr * rot / (f0 + 1e-3))
        const float x = __fdividef(__fmul_rn(P.sr, rot[i]), __fadd_rn(f[i], 1e-3f));
        float c = sinc_f(x);
        if (P.zero_unvoiced && f[i] <= 0.0f) c = 0.0f;     // vocoder.py:460
        const int j = 16 * lane + i;
        slot[j + (j >> 5)] = c;
    }
}

__global__ void __launch_bounds__(kCsfThreads, 1) combsubfast_kernel(const CsfParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    float* win = reinterpret_cast<float*>(smem_raw + 1024 * 8);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + 1024 * 8 + 1024 * 4 + wid * kCsfWarpBytes;
    float* plane = reinterpret_cast<float*>(wbase);
    float* ring = plane + kPlaneFloats;
    float2* stash = reinterpret_cast<float2*>(ring + 2 * kRingSlot);

    init_fft_tables(tw, win, threadIdx.x, blockDim.x);
    if (P.window) for (int e = threadIdx.x; e < 1024; e += blockDim.x) win[e] = __ldg(P.window + e);
    __syncthreads();

    const int64_t run = (int64_t)blockIdx.x * kCsfWarps + wid;
    if (run >= (int64_t)P.B * P.runs_per_clip) return;
    const int b = (int)(run / P.runs_per_clip);
    const int p_begin = (int)(run % P.runs_per_clip) * P.run_len;
    const int p_end = min(P.pairs_per_clip, p_begin + P.run_len);
    const int p_first = p_begin > 0 ? p_begin - 1 : 0;     // halo pair recomputed for its tail
    const int F = P.F;
    const int64_t T = (int64_t)F * kHop;
    const double init_rot = P.initial_phase ? ((double)P.initial_phase[b] / 2.0) / 3.14159265358979323846 : 0.0;
    const float* hm_b = P.hm + (int64_t)b * P.cB;
    const float* hp_b = P.hp + (int64_t)b * P.cB;
    const float* nm_b = P.nm + (int64_t)b * P.cB;
    const float* u_b = P.noise_u ? P.noise_u + (int64_t)b * T : nullptr;
    const uint32_t key = noise_key(P.seed, (uint32_t)b);
    float* out_b = P.signal + (int64_t)b * T;
    const int partner = (32 - lane) & 31;

    float re[32], im[32];
    float carry[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) carry[q] = 0.0f;

    // first hop of the first frame of the run
    csf_gen_hop(P, b, 2 * p_first - 1, ring + ((2 * p_first - 1) & 1) * kRingSlot, init_rot, lane);

    for (int p = p_first; p < p_end; ++p) {
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
            const int fm = 2 * p + s;                       // frame handled by steps 0 and 1
            if (s < 2) {
                // ---- excitation hop fm (second half of frame fm), then the windowed frame ----
                csf_gen_hop(P, b, fm, ring + (fm & 1) * kRingSlot, init_rot, lane);
                __syncwarp();
                const float* slotA = ring + ((fm - 1) & 1) * kRingSlot;   // hop fm-1
                const float* slotB = ring + (fm & 1) * kRingSlot;         // hop fm
                const int64_t tA = (int64_t)(fm - 1) * kHop;              // first sample of the frame
                const bool okA = (fm - 1 >= 0) && (fm - 1 < F), okB = fm < F;
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    const int i = 32 * n1 + lane;
                    const float w = win[i];
                    const float c = (n1 < 16 ? slotA : slotB)[(i & (kHop - 1)) + (n1 & 15)];
                    float nz = 0.0f;
                    if (n1 < 16 ? okA : okB) {
                        const int64_t t = tA + i;
                        const float u = u_b ? __ldg(u_b + t) : noise_uniform(key, (uint32_t)t);
                        nz = __fmaf_rn(2.0f, u, -1.0f);                   // vocoder.py:461
                    }
                    re[n1] = w * c;
                    im[n1] = w * nz;
                }
            }
            // s == 2: re/im already hold the packed spectrum of the pair (swapped for the inverse)

            warp_fft1024(re, im, plane, tw, lane);

            if (s < 2) {
                // ---- split the two real spectra, apply the filters (vocoder.py:472-481) ----
                const int mhat = min(fm, F - 1);                           // last filter frame repeated (:473,476)
                const float* hm_r = hm_b + (int64_t)mhat * P.cF;
                const float* hp_r = hp_b + (int64_t)mhat * P.cF;
                const float* nm_r = nm_b + (int64_t)mhat * P.cF;
                float yr[17], yi[17];
#pragma unroll
                for (int q = 0; q < 17; ++q) {
                    const int k = lane + 32 * q;
                    if (q == 16 && lane != 0) { yr[16] = 0.0f; yi[16] = 0.0f; continue; }
                    const float a = re[brev5(q & 31)], bb = im[brev5(q & 31)];
                    float c, d;
                    if (q < 16) {
                        c = __shfl_sync(kFullMask, re[brev5(31 - q)], partner);
                        d = __shfl_sync(kFullMask, im[brev5(31 - q)], partner);
                        if (lane == 0) { c = re[brev5((32 - q) & 31)]; d = im[brev5((32 - q) & 31)]; }
                    } else {        // bin 512 (lane 0 only): its own partner
                        c = a; d = bb;
                    }
                    const float Cr = a + c, Ci = bb - d, Nr = bb + d, Ni = c - a;
                    // H = exp(hm + j*pi*hp) (vocoder.py:472), N = exp(nm)/128 (:475); the 1/2 of the
                    // split and the 1/1024 of irfft are folded in as exact powers of two.
                    const float g = exp2f(fmaf(__ldg(hm_r + k), DDSP_LOG2E_F, -11.0f));
                    float sn, cs;
                    __sincosf(DDSP_PI_F * __ldg(hp_r + k), &sn, &cs);
                    const float Hr = g * cs, Hi = g * sn;
                    const float nf = exp2f(fmaf(__ldg(nm_r + k), DDSP_LOG2E_F, -18.0f));
                    yr[q] = fmaf(Cr, Hr, fmaf(-Ci, Hi, Nr * nf));
                    yi[q] = fmaf(Cr, Hi, fmaf(Ci, Hr, Ni * nf));
                    if (q == 16 || (q == 0 && lane == 0)) yi[q] = 0.0f;    // irfft ignores Im of DC / Nyquist
                }
                if (s == 0) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) stash[q * 32 + lane] = make_float2(yr[q], yi[q]);
                    if (lane == 0) stash[16 * 32] = make_float2(yr[16], yi[16]);
                } else {
                    // ---- V = Y_m + j*Y_{m+1}; pass (Im V, Re V) to the forward FFT = inverse ----
                    float xr[16], xi[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const float2 y = stash[q * 32 + lane];
                        // V[k]   = (y.x - yi) + j (y.y + yr);  V[N-k] = (y.x + yi) + j (yr - y.y)
                        re[q] = y.y + yr[q];       // swapped: "real" input = Im V
                        im[q] = y.x - yi[q];
                        xr[q] = y.x + yi[q];       // Re V[N-k]
                        xi[q] = yr[q] - y.y;       // Im V[N-k]
                    }
                    float v512r = 0.0f, v512i = 0.0f;
                    if (lane == 0) { v512r = stash[16 * 32].x; v512i = yr[16]; }
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        float tr = __shfl_sync(kFullMask, xr[q], partner);
                        float ti = __shfl_sync(kFullMask, xi[q], partner);
                        if (lane == 0) {
                            tr = (q == 15) ? v512r : xr[(q + 1) & 15];
                            ti = (q == 15) ? v512i : xi[(q + 1) & 15];
                        }
                        re[31 - q] = ti;           // swapped
                        im[31 - q] = tr;
                    }
                }
                __syncwarp();
            } else {
                // ---- window (vocoder.py:486), overlap-add (:485-487), crop (:490) ----
                // after the swapped FFT: im[] = Re v = frame 2p, re[] = Im v = frame 2p+1
                const int hopA = 2 * p - 1, hopB = 2 * p;
                const bool live = p >= p_begin;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int n = lane + 32 * q;
                    const float wa = win[n], wb = win[n + kHop];
                    const float m_first = im[brev5(q)] * wa, m_second = im[brev5(q + 16)] * wb;
                    const float n_first = re[brev5(q)] * wa, n_second = re[brev5(q + 16)] * wb;
                    if (live && hopA >= 0) out_b[(int64_t)hopA * kHop + n] = carry[q] + m_first;
                    if (live && hopB < F) out_b[(int64_t)hopB * kHop + n] = m_second + n_first;
                    carry[q] = n_second;
                }
            }
        }
    }
}

}  // namespace ddsp
