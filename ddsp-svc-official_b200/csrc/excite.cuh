// excite.cuh -- excitation generators that write a (B,T) buffer:
//   * combtooth sinc pulse train            vocoder.py:539 (CombSub-old; no unvoiced zeroing)
//   * Sins harmonic oscillator bank          vocoder.py:397,402-412 (+ core.py:24-28 Nyquist mask)
//   * elementwise add                        vocoder.py:421,548
#pragma once
#include "fft32.cuh"
#include "phase.cuh"

namespace ddsp {

// One warp per hop; lane owns 16 consecutive samples (128-bit stores).
__global__ void __launch_bounds__(256) combtooth_kernel(const float* __restrict__ f0_frames, int64_t fB, int64_t fF,
                                                        int B, int F, double inv_sr, float sr,
                                                        const double* __restrict__ prefix, int zero_unvoiced,
                                                        float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= (int64_t)B * F) return;
    const int b = (int)(warp / F), h = (int)(warp % F);
    const float* row = f0_frames + (int64_t)b * fB;
    const float x0 = __ldg(row + (int64_t)h * fF);
    const float x1 = __ldg(row + (int64_t)min(h + 1, F - 1) * fF);
    // the arithmetic of the CombSubFast excitation (csf_gen_hop): fixed-point phase inside the lane on top of the exact fp64
    // hop prefix, packed fp32x2 sinc -- two samples per instruction instead of an fp64 running sum per sample
    float2 f2[8], rot2[8];
    float c[16];
    hop_rotation_q32(x0, x1, prefix[warp], inv_sr, lane, f2, rot2);
    const float sr_scale = sr * 2.3283064365386963e-10f;      // rot2 holds rot * 2^32
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 den = add2(f2[j], bc2(1e-3f));                                   // vocoder.py:539
        const float2 xs = fma2(mul2(bc2(sr_scale), rot2[j]), make_float2(rcp_approx(den.x), rcp_approx(den.y)), bc2(1e-30f));
        const float2 v = sinc2_xs(xs);
        c[2 * j] = (zero_unvoiced && f2[j].x <= 0.0f) ? 0.0f : v.x;
        c[2 * j + 1] = (zero_unvoiced && f2[j].y <= 0.0f) ? 0.0f : v.y;
    }
    float4* dst = reinterpret_cast<float4*>(out + warp * kHop + 16 * lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_float4(c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3]);
}

// Sins oscillator bank.  One CTA of kOscThreads threads per hop (512 samples, 2*kOscPairs per thread).
//   A[m,k] = fl(exp(a[m,k]) / 128) * ((f0[m]*k < sr/2) + 1e-7)            (vocoder.py:397,402)
//   s[t]   = sum_k lerp(A[:,k])[t] * sin(k * theta[t])                     (vocoder.py:406-412)
// sin(k*theta) comes from a two-term recurrence (see below); two samples ride in one packed
// fp32x2 register pair.  n_harm must be even.
constexpr int kSinsMaxHarm = 512;
#ifndef SINS_PAIRS
#define SINS_PAIRS 4                               // packed sample pairs per thread
#endif
#ifndef SINS_SKIP_MASKED
#define SINS_SKIP_MASKED 1                         // skip harmonics that are above Nyquist in both frames of a hop
#endif
constexpr int kOscPairs = SINS_PAIRS;
constexpr int kOscThreads = kHop / (2 * kOscPairs);   // one CTA per hop

__global__ void __launch_bounds__(kOscThreads) sins_osc_kernel(const float* __restrict__ amp_ctrl, int64_t cB, int64_t cF,
                                                       int n_harm, const float* __restrict__ f0_frames, int64_t fB,
                                                       int64_t fF, int F, float fmax,
                                                       const float* __restrict__ phase_full,
                                                       float* __restrict__ out) {
    __shared__ float2 amps[kSinsMaxHarm];      // (A0, dA) per harmonic: broadcast operands of the packed interpolation
    __shared__ int k_live;                      // harmonics below Nyquist in at least one of the hop's two frames
    if (threadIdx.x == 0) k_live = 0;
    __syncthreads();
    const int hop = blockIdx.x % F, b = blockIdx.x / F;
    const int m1 = min(hop + 1, F - 1);         // hold-last (core.py:17)
    const float f0a = __ldg(f0_frames + (int64_t)b * fB + (int64_t)hop * fF);
    const float f0b = __ldg(f0_frames + (int64_t)b * fB + (int64_t)m1 * fF);
    const float* ra = amp_ctrl + (int64_t)b * cB + (int64_t)hop * cF;
    const float* rb = amp_ctrl + (int64_t)b * cB + (int64_t)m1 * cF;
    for (int k = threadIdx.x; k < n_harm; k += blockDim.x) {
        const float lvl = (float)(k + 1);
        const float ma = __fadd_rn((__fmul_rn(f0a, lvl) < fmax) ? 1.0f : 0.0f, 1e-7f);     // core.py:26-27
        const float mb = __fadd_rn((__fmul_rn(f0b, lvl) < fmax) ? 1.0f : 0.0f, 1e-7f);
        const float A0 = __fmul_rn(__fmul_rn(expf(__ldg(ra + k)), 0.0078125f), ma);         // exp()/128 then mask
        const float A1 = __fmul_rn(__fmul_rn(expf(__ldg(rb + k)), 0.0078125f), mb);
        const float dA = A1 - A0;
        amps[k] = make_float2(A0, dA);
#if SINS_SKIP_MASKED
        if (ma > 0.5f || mb > 0.5f) atomicMax(&k_live, k + 1);
#endif
    }
    __syncthreads();
#if SINS_SKIP_MASKED
    // Harmonics above Nyquist in BOTH frames of the hop carry the mask weight 1e-7 (core.py:27) throughout the hop:
    // their sum is below 1.3e-7 of full scale for any amplitudes a trained network emits (sum_k exp(a_k)/128 ~ 1)
    // and the loop stops at the last harmonic that is live in either frame (f0 = 800 Hz: 27 of 128).  The mask
    // values of the live harmonics -- including the partially masked ones at a frame boundary -- are the
    // reference's, bit for bit.
    const int n_loop = min(n_harm, (k_live + 1) & ~1);
#else
    const int n_loop = n_harm;
#endif
    const int64_t base = ((int64_t)b * F + hop) * kHop;
    const int t = threadIdx.x;
    // sin(k*theta) by Reinsch's stable recurrence:  d_{k+1} = d_k + delta*s_k,  s_{k+1} = s_k + d_{k+1},
    // delta = -4 sin^2(phi/2), which needs |phi| <= pi/2: theta in the outer half of [-pi,pi] is shifted by
    // +-pi, sin(k*theta) = (-1)^k sin(k*phi), i.e. the odd harmonics change sign (separate accumulators).
    float2 lam[kOscPairs], dl[kOscPairs], sk[kOscPairs], dk[kOscPairs], acc_e[kOscPairs], acc_o[kOscPairs], sgn[kOscPairs];
#pragma unroll
    for (int p = 0; p < kOscPairs; ++p) {
        const int i0 = t + 2 * kOscThreads * p, i1 = i0 + kOscThreads;
        lam[p] = make_float2((float)i0 * (1.0f / kHop), (float)i1 * (1.0f / kHop));
        float th[2] = {__ldg(phase_full + base + i0), __ldg(phase_full + base + i1)};
        float hx[2], gv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const bool shift = fabsf(th[e]) > 0.5f * DDSP_PI_F;
            const float phi = shift ? th[e] - copysignf(DDSP_PI_F, th[e]) : th[e];
            hx[e] = 0.5f * phi;                   // |hx| <= pi/4: no range reduction needed
            gv[e] = shift ? -1.0f : 1.0f;
        }
        // sin / cos of the half angle on [-pi/4, pi/4] by the minimax polynomials of the single-precision
        // libm kernels (errors below 1 ulp there), both samples of the pair in packed arithmetic:
        // a dozen instructions per pair instead of two sincosf calls
        const float2 x = make_float2(hx[0], hx[1]);
        const float2 z = mul2(x, x);
        float2 ps = fma2(z, bc2(-1.9515295891e-4f), bc2(8.3321608736e-3f));
        ps = fma2(z, ps, bc2(-1.6666654611e-1f));
        const float2 sh = fma2(mul2(z, x), ps, x);                              // x + x^3 P(z)
        float2 pc = fma2(z, bc2(2.443315711809948e-5f), bc2(-1.388731625493765e-3f));
        pc = fma2(z, pc, bc2(4.166664568298827e-2f));
        const float2 ch = fma2(mul2(z, z), pc, fma2(z, bc2(-0.5f), bc2(1.0f)));   // 1 - z/2 + z^2 Q(z)
        dl[p] = mul2(mul2(sh, sh), bc2(-4.0f));
        sk[p] = mul2(mul2(sh, ch), bc2(2.0f));   // s_1
        dk[p] = sk[p];                            // d_1 = s_1 - s_0
        sgn[p] = make_float2(gv[0], gv[1]);
        acc_e[p] = make_float2(0.0f, 0.0f);
        acc_o[p] = make_float2(0.0f, 0.0f);
    }
#pragma unroll 4
    for (int k = 0; k < n_loop; k += 2) {        // k, k+1 are harmonics k+1 (odd) and k+2 (even); n_harm is even
        // (A0, dA) enter as single-register broadcast operands: a packed FMA with three register-PAIR operands takes
        // 3 cycles of register reads instead of 2 (profiles/ubench/ffma2_issue.cu)
        const float4 aa = *reinterpret_cast<const float4*>(&amps[k]);       // harmonics k+1, k+2
#pragma unroll
        for (int p = 0; p < kOscPairs; ++p) {
            acc_o[p] = fma2(fma2(lam[p], bc2(aa.y), bc2(aa.x)), sk[p], acc_o[p]);
            dk[p] = fma2(dl[p], sk[p], dk[p]);
            sk[p] = add2(sk[p], dk[p]);
            acc_e[p] = fma2(fma2(lam[p], bc2(aa.w), bc2(aa.z)), sk[p], acc_e[p]);
            dk[p] = fma2(dl[p], sk[p], dk[p]);
            sk[p] = add2(sk[p], dk[p]);
        }
    }
#pragma unroll
    for (int p = 0; p < kOscPairs; ++p) {
        const float2 r = fma2(sgn[p], acc_o[p], acc_e[p]);
        out[base + t + 2 * kOscThreads * p] = r.x;
        out[base + t + 2 * kOscThreads * p + kOscThreads] = r.y;
    }
}

// Caller-side epilogue of main.py:116,159 / gui.py:112,127 fused:  signal *= upsample(mask_frames, hop)
// without materialising the (B,T) mask.  One thread per 4 consecutive samples (same frame).
__global__ void __launch_bounds__(256) apply_frame_mask_kernel(float4* __restrict__ signal,
                                                               const float* __restrict__ mask_frames, int64_t mB,
                                                               int64_t mF, int B, int F) {
    const int64_t n4 = (int64_t)B * F * (kHop / 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hop_id = i / (kHop / 4);
        const int j = (int)(i % (kHop / 4)) * 4;
        const int b = (int)(hop_id / F), h = (int)(hop_id % F);
        const float* row = mask_frames + (int64_t)b * mB;
        const float x0 = __ldg(row + (int64_t)h * mF), x1 = __ldg(row + (int64_t)min(h + 1, F - 1) * mF);
        float4 v = signal[i];
        v.x = __fmul_rn(v.x, lerp_torch(x0, x1, (float)(j + 0) * (1.0f / kHop)));
        v.y = __fmul_rn(v.y, lerp_torch(x0, x1, (float)(j + 1) * (1.0f / kHop)));
        v.z = __fmul_rn(v.z, lerp_torch(x0, x1, (float)(j + 2) * (1.0f / kHop)));
        v.w = __fmul_rn(v.w, lerp_torch(x0, x1, (float)(j + 3) * (1.0f / kHop)));
        signal[i] = v;
    }
}

// The whole silence-mask epilogue of main.py:112-116,159 / gui.py:108-112,127 in one in-place pass:
//   mask = (volume > threshold)  ->  edge-padded 9-frame maximum  ->  upsample  ->  signal *= mask
// (the comparison is made in double like numpy's float32-array > python-float; edge padding == index clamping).
__global__ void __launch_bounds__(256) apply_volume_mask_kernel(float4* __restrict__ signal, const float* __restrict__ volume,
                                                                int64_t vB, int64_t vF, double threshold, int B, int F) {
    const int64_t n4 = (int64_t)B * F * (kHop / 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hop_id = i / (kHop / 4);
        const int j = (int)(i % (kHop / 4)) * 4;
        const int b = (int)(hop_id / F), h = (int)(hop_id % F);
        const float* row = volume + (int64_t)b * vB;
        // frames h-4 .. h+5 cover the 9-frame windows of mask[h] and mask[h+1]
        bool any_lo = false, any_hi = false;          // window of h: d = -4..4; window of h+1: d = -3..5
#pragma unroll
        for (int d = -4; d <= 5; ++d) {
            const bool on = (double)__ldg(row + (int64_t)min(max(h + d, 0), F - 1) * vF) > threshold;
            if (d <= 4) any_lo |= on;
            if (d >= -3) any_hi |= on;
        }
        // mask[h+1] with hold-last at the end (core.py:17): frame F is frame F-1, whose window is d = -4..4 around F-1
        if (h + 1 > F - 1) any_hi = any_lo;
        const float x0 = any_lo ? 1.0f : 0.0f, x1 = any_hi ? 1.0f : 0.0f;
        float4 v = signal[i];
        v.x = __fmul_rn(v.x, lerp_torch(x0, x1, (float)(j + 0) * (1.0f / kHop)));
        v.y = __fmul_rn(v.y, lerp_torch(x0, x1, (float)(j + 1) * (1.0f / kHop)));
        v.z = __fmul_rn(v.z, lerp_torch(x0, x1, (float)(j + 2) * (1.0f / kHop)));
        v.w = __fmul_rn(v.w, lerp_torch(x0, x1, (float)(j + 3) * (1.0f / kHop)));
        signal[i] = v;
    }
}

__global__ void __launch_bounds__(256) add_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                  float4* __restrict__ out, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 x = __ldg(a + i), y = __ldg(b + i);
        out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
}

}  // namespace ddsp
