// combsubfast_bwd.cuh -- gradient of CombSubFast stage B (ddsp/vocoder.py:455-492) with respect to
// the three control tensors, as ONE kernel (makes the drop-in usable
// under autograd, solver.py:111).  f0 / phase carry no gradient (they are data).
//
// Forward, per frame m = 0..F with filter row r = min(m, F-1):
//     Y_m = rfft(w*comb_m) * exp(hm_r + j*pi*hp_r) + rfft(w*noise_m) * exp(nm_r)/128
//     signal = crop(overlap_add(w * irfft(Y_m)))
// With g = dL/dsignal and q_m[i] = w[i] * g[512(m-1)+i] (zero outside the clip):
//     Q_m = rfft(q_m),   G_m[k] = c_k/1024 * Q_m[k]   (c_k = 2, but 1 for k = 0 and 512 whose
//                                                      imaginary parts irfft ignores -> Im G = 0)
//     A = rfft(w*comb_m) * H,   Bn = rfft(w*noise_m) * N
//     dL/dhm_r[k] += Re G Re A + Im G Im A
//     dL/dhp_r[k] += pi * (Im G Re A - Re G Im A)
//     dL/dnm_r[k] += Re G Re Bn + Im G Im Bn
// (real-to-complex transforms are their own adjoints up to the c_k weights, so no inverse FFT is
// needed: three forward FFTs per frame pair, the same count as the forward kernel.)
//
// Work decomposition mirrors combsubfast.cuh: a warp owns a run of frame pairs; per pair one
// complex FFT of (q_2p + j*q_2p+1) gives both gradient spectra, which are parked in shared
// memory, then one complex FFT per frame regenerates (comb + j*noise) exactly as the forward pass
// did (same excitation code, same noise stream).  Rows r < F-1 receive exactly one frame and are
// written with plain stores; row F-1 receives frames F-1 and F and takes two atomic adds onto
// zeros (order-independent).
#pragma once
#include "combsubfast.cuh"

namespace ddsp {

constexpr int kCsbWarps = 12;                       // two parked spectra per warp -> fewer warps fit
constexpr int kCsbThreads = kCsbWarps * 32;
constexpr int kCsbWarpBytes = kPlaneFloats * 4 + 2 * kRingSlot * 4 + 2 * kStashFloat2 * 8 + kCsfCtxInts * 4;
constexpr int kCsbSmemBytes = kTableBytes + kCsbWarps * kCsbWarpBytes;
static_assert(kCsbSmemBytes <= 227 * 1024, "backward kernel shared memory");

struct CsbParams {
    CsfParams fwd;                                   // inputs of the forward pass (signal unused)
    const float* grad_signal;                        // (B,T)
    float* ghm; float* ghp; float* gnm;              // (B,F,513) views, strides (gB,gF,1)
    int64_t gB, gF;
};

template <bool HAS_U>
__global__ void __launch_bounds__(kCsbThreads, 1) combsubfast_backward_kernel(const CsbParams PB) {
    const CsfParams& P = PB.fwd;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const float4* tw4 = reinterpret_cast<const float4*>(smem_raw);
    float* win = reinterpret_cast<float*>(smem_raw + 512 * 16);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + kTableBytes + wid * kCsbWarpBytes;
    float* plane = reinterpret_cast<float*>(wbase);
    float* ring = plane + kPlaneFloats;
    float2* stash = reinterpret_cast<float2*>(ring + 2 * kRingSlot);      // [2][kStashFloat2]: Q_2p, Q_2p+1

    {
        const float4* src = reinterpret_cast<const float4*>(P.tables);
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        for (int e = threadIdx.x; e < kTableBytes / 16; e += kCsbThreads) dst[e] = __ldg(src + e);
        __syncthreads();
        if (P.window) {
            for (int e = threadIdx.x; e < 1024; e += kCsbThreads) {
                const float w = __ldg(P.window + e);
                win[win_analysis_index(e)] = w;
                win[win_synthesis_index(e)] = w;
            }
            __syncthreads();
        }
    }

    volatile int* ctx = reinterpret_cast<volatile int*>(stash + 2 * kStashFloat2);
    {
        const int64_t run = (int64_t)wid * gridDim.x + blockIdx.x;          // round-robin over the CTAs
        if (run >= (int64_t)P.B * P.runs_per_clip) return;
        const int b0 = (int)(run / P.runs_per_clip);
        const int r = (int)(run % P.runs_per_clip);
        const int pb = csf_run_begin(r, P.run_len, P.run_rem);
        if (lane == 0) {
            ctx[0] = b0;
            ctx[1] = pb;
            ctx[2] = pb + P.run_len + (r < P.run_rem ? 1 : 0);
            const uint64_t k64 = noise_key64(P.seed, (uint32_t)b0);
            ctx[3] = (int)(uint32_t)k64;
            ctx[5] = (int)(uint32_t)(k64 >> 32);
        }
        __syncwarp();
    }
    const int F = P.F;
    const int partner = (32 - lane) & 31;
    const bool lane0 = lane == 0;

    Pts32 X;
    // steps per pair: s=0 gradient spectra of frames 2p, 2p+1; s=1 frame 2p; s=2 frame 2p+1.
    // s=-1 (first iteration only) generates the excitation hop preceding the run.
    int p = ctx[1], s = -1;
#pragma unroll 1
    for (;;) {
        const int b = ctx[0];
        const int fm = 2 * p + (s <= 0 ? s : s - 1);         // s=-1: 2p-1, s=1: 2p, s=2: 2p+1
        if (s != 0) {
            csf_gen_hop(P, fm, csf_load_hop(P, b, fm), ring + (fm & 1) * kRingSlot, lane);
            __syncwarp();
            if (s < 0) { s = 0; continue; }
            {   // pull this frame's three control rows into L2 while the FFT runs (lanes 0..16: one line each)
                const int64_t rp = (int64_t)b * P.cB + (int64_t)min(fm, F - 1) * P.cF + 32 * lane;
                if (lane <= 16) { prefetch_l2(P.hm + rp); prefetch_l2(P.hp + rp); prefetch_l2(P.nm + rp); }
            }
            csf_load_frame<HAS_U>(P, X, ring, win, fm, b, (uint32_t)ctx[3], (uint32_t)ctx[5], lane);
        } else {
            // ---- q_2p -> real part, q_2p+1 -> imaginary part --------------------------------------
            // frame m spans output hops m-1 (first half) and m (second half); hops outside [0,F) carry no gradient
            const int hA = 2 * p - 1, hB = 2 * p, hC = 2 * p + 1;
            const float* g_b = PB.grad_signal + (int64_t)b * F * kHop + lane;
            const bool vA = hA >= 0 && hA < F, vB = hB < F, vC = hC < F;
            const float* gA = g_b + (vA ? (int64_t)hA * kHop : 0);
            const float* gB_ = g_b + (vB ? (int64_t)hB * kHop : 0);
            const float* gC = g_b + (vC ? (int64_t)hC * kHop : 0);
            const float okA = vA ? 1.0f : 0.0f, okB = vB ? 1.0f : 0.0f, okC = vC ? 1.0f : 0.0f;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const float2 wp = reinterpret_cast<const float2*>(win + 1024)[n1 * 32 + lane];   // synthesis-order pairs
                const float w0 = wp.x, w1 = wp.y;
                const float a = __ldg(gA + 32 * n1) * okA, bm = __ldg(gB_ + 32 * n1) * okB, c = __ldg(gC + 32 * n1) * okC;
                DDSP_RE(X, brev5(n1)) = w0 * a;              // frame 2p,   first half
                DDSP_RE(X, brev5(n1 + 16)) = w1 * bm;        // frame 2p,   second half
                DDSP_IM(X, brev5(n1)) = w0 * bm;             // frame 2p+1, first half
                DDSP_IM(X, brev5(n1 + 16)) = w1 * c;         // frame 2p+1, second half
            }
        }

        warp_fft1024(X, plane, tw4, lane);

        // ---- split the packed spectrum into the two real-input spectra (x2) ---------------------
        // first = spectrum of the real part, second = spectrum of the imaginary part
        if (s == 0) {
            // gradient spectra of frames 2p / 2p+1 -> the two parked slots; bins 0 and 512 (lane 0,
            // q = 0 / 16) drop their imaginary part (irfft ignores it, so it has no gradient)
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float a = DDSP_RE(X, q), bb = DDSP_IM(X, q);
                float c = __shfl_sync(kFullMask, DDSP_RE(X, 31 - q), partner);
                float d = __shfl_sync(kFullMask, DDSP_IM(X, 31 - q), partner);
                const float c0 = DDSP_RE(X, (32 - q) & 31), d0 = DDSP_IM(X, (32 - q) & 31);
                c = lane0 ? c0 : c;
                d = lane0 ? d0 : d;
                const bool edge = lane0 && q == 0;
                stash[q * 32 + lane] = make_float2(a + c, edge ? 0.0f : bb - d);
                stash[kStashFloat2 + q * 32 + lane] = make_float2(bb + d, edge ? 0.0f : c - a);
            }
            if (lane0) {
                stash[16 * 32] = make_float2(2.0f * DDSP_RE(X, 16), 0.0f);
                stash[kStashFloat2 + 16 * 32] = make_float2(2.0f * DDSP_IM(X, 16), 0.0f);
            }
            s = 1;
            continue;
        }
        if (fm <= F) {          // (frame F+1 of an odd-length pair list does not exist)
            const int row = min(fm, F - 1);
            const float* hm_r = P.hm + ((int64_t)b * P.cB + (int64_t)row * P.cF + lane);
            const float* hp_r = P.hp + ((int64_t)b * P.cB + (int64_t)row * P.cF + lane);
            const float* nm_r = P.nm + ((int64_t)b * P.cB + (int64_t)row * P.cF + lane);
            float* ghm_r = PB.ghm + ((int64_t)b * PB.gB + (int64_t)row * PB.gF + lane);
            float* ghp_r = PB.ghp + ((int64_t)b * PB.gB + (int64_t)row * PB.gF + lane);
            float* gnm_r = PB.gnm + ((int64_t)b * PB.gB + (int64_t)row * PB.gF + lane);
            const bool shared_row = fm >= F - 1;             // row F-1 also serves frame F
            const float2* Qs = stash + (s == 2 ? kStashFloat2 : 0);
            // control loads run kLook bins ahead of their use (bin 512 lives on lane 0 only; others read a dummy)
            constexpr int kLook = 4;
            const int k16 = lane0 ? 512 : 0;
            float chm[kLook], chp[kLook], cnm[kLook];
#pragma unroll
            for (int q = 0; q < kLook; ++q) {
                chm[q] = __ldg(hm_r + 32 * q); chp[q] = __ldg(hp_r + 32 * q); cnm[q] = __ldg(nm_r + 32 * q);
            }
#pragma unroll
            for (int q = 0; q < 17; ++q) {
                const float vhm = chm[q % kLook], vhp = chp[q % kLook], vnm = cnm[q % kLook];
                if (q + kLook < 17) {
                    const int offn = (q + kLook < 16) ? 32 * (q + kLook) : k16;
                    chm[q % kLook] = __ldg(hm_r + offn); chp[q % kLook] = __ldg(hp_r + offn); cnm[q % kLook] = __ldg(nm_r + offn);
                }
                float a, bb, c, d;
                if (q < 16) {
                    a = DDSP_RE(X, q); bb = DDSP_IM(X, q);
                    c = __shfl_sync(kFullMask, DDSP_RE(X, 31 - q), partner);
                    d = __shfl_sync(kFullMask, DDSP_IM(X, 31 - q), partner);
                    const float c0 = DDSP_RE(X, (32 - q) & 31), d0 = DDSP_IM(X, (32 - q) & 31);
                    c = lane0 ? c0 : c;
                    d = lane0 ? d0 : d;
                } else {
                    a = DDSP_RE(X, 16); bb = DDSP_IM(X, 16); c = a; d = bb;
                }
                const float Cr = a + c, Ci = bb - d, Nr = bb + d, Ni = c - a;
                const float2 G = Qs[(q < 16) ? q * 32 + lane : 16 * 32];
                // 1/2 (split of C) * 1/2 (split of Q) * c_k/1024 folded into the exponentials;
                // c_k = 1 instead of 2 for bins 0 and 512 (lane 0, q = 0 / 16)
                const float esh = (lane0 && (q == 0 || q == 16)) ? 1.0f : 0.0f;
                const float gmag = ex2_approx(fmaf(vhm, DDSP_LOG2E_F, -11.0f - esh));
                const float ang = DDSP_PI_F * vhp;
                const float Hr = gmag * cos_approx(ang), Hi = gmag * sin_approx(ang);
                const float nf = ex2_approx(fmaf(vnm, DDSP_LOG2E_F, -18.0f - esh));
                const float Ar = fmaf(Cr, Hr, -Ci * Hi), Ai = fmaf(Cr, Hi, Ci * Hr);
                const float d_hm = fmaf(G.x, Ar, G.y * Ai);
                const float d_hp = DDSP_PI_F * fmaf(G.y, Ar, -G.x * Ai);
                const float d_nm = nf * fmaf(G.x, Nr, G.y * Ni);
                if (q < 16 || lane0) {
                    const int off = (q < 16) ? 32 * q : 512;
                    if (shared_row) {
                        atomicAdd(ghm_r + off, d_hm);
                        atomicAdd(ghp_r + off, d_hp);
                        atomicAdd(gnm_r + off, d_nm);
                    } else {
                        ghm_r[off] = d_hm;
                        ghp_r[off] = d_hp;
                        gnm_r[off] = d_nm;
                    }
                }
            }
        }
        __syncwarp();
        if (s == 2) {
            if (++p >= ctx[2]) break;
            s = 0;
        } else {
            s = 2;
        }
    }
}

}  // namespace ddsp
