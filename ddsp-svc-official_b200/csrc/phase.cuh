// phase.cuh -- stage A (f0 upsample + fp64 phase accumulation) and the standalone core ops
// upsample / fo_to_rot / remove_above_fmax.
//
// Reference: ddsp/core.py:7-21 (upsample), :31-51 (fo_to_rot), :24-28 (remove_above_fmax);
// callers ddsp/vocoder.py:391-393, 449-451, 515-517.
#pragma once
#include "common.cuh"

namespace ddsp {

// ---------------------------------------------------------------------------------------------
// A1: per-hop totals.  A warp sums the 512 upsampled fp32 f0 samples of a hop exactly (the total is
// an fp64 number without rounding for any realistic f0: 24-bit mantissas, 512 terms).
// ---------------------------------------------------------------------------------------------
constexpr int kHopsPerWarp = 8;   // hops handled by one warp (amortises the interpolation weights)

__global__ void __launch_bounds__(256) hop_totals_kernel(const float* __restrict__ f0_frames, int64_t fB,
                                                         int64_t fF, int B, int F,
                                                         double* __restrict__ totals) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t first = warp * kHopsPerWarp, n_hops = (int64_t)B * F;
    if (first >= n_hops) return;
    // lane owns the 16 consecutive samples 16*lane .. 16*lane+15 of a hop, as 8 packed fp32x2 pairs;
    // the weights j/512 are exact, so lambda_0 + i/512 equals (16 lane + i)/512 bit for bit
    float2 w0[8], w1[8];
    {
        const float lam0 = (float)(16 * lane) * (1.0f / kHop);
        float2 lam = make_float2(lam0, lam0 + 1.0f / kHop);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            w1[j] = lam;
            w0[j] = sub2(bc2(1.0f), lam);
            lam = add2(lam, bc2(2.0f / kHop));
        }
    }
    // lane g (< kHopsPerWarp) fetches the two frame values of hop first+g
    float x0 = 0.0f, x1 = 0.0f;
    if (lane < kHopsPerWarp && first + lane < n_hops) {
        const int b = (int)((first + lane) / F), h = (int)((first + lane) % F);
        const float* row = f0_frames + (int64_t)b * fB;
        x0 = __ldg(row + (int64_t)h * fF);
        x1 = __ldg(row + (int64_t)min(h + 1, F - 1) * fF);
    }
#pragma unroll 2
    for (int g = 0; g < kHopsPerWarp; ++g) {
        const float a = __shfl_sync(kFullMask, x0, g), c = __shfl_sync(kFullMask, x1, g);
        double s;
        if (fabsf(c - a) < 2.0f * fminf(a, c)) {
            // Both frames voiced and within a factor 3 (every hop but onsets / offsets; warp-uniform test).
            // The lane's 16 samples then differ by less than min(a,c)/16, so each difference f_i - f_0 is
            // exact in fp32 (Sterbenz) and so is any sum of them: all are multiples of ulp(f_min) and stay
            // below 2^24 ulps.  Two fp32->fp64 conversions per lane instead of 16, packed fp32x2 arithmetic,
            // and the lane total is the same exact number.
            const float2 f0p = fma2(w0[0], bc2(a), mul2(w1[0], bc2(c)));        // lerp_torch on both halves
            float2 gs = make_float2(0.0f, __fsub_rn(f0p.y, f0p.x));
#pragma unroll
            for (int j = 1; j < 8; ++j)
                gs = add2(gs, sub2(fma2(w0[j], bc2(a), mul2(w1[j], bc2(c))), bc2(f0p.x)));
            s = fma(16.0, (double)f0p.x, (double)__fadd_rn(gs.x, gs.y));
        } else {
            s = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 f = fma2(w0[j], bc2(a), mul2(w1[j], bc2(c)));
                s += (double)f.x;
                s += (double)f.y;
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(kFullMask, s, d);
        if (lane == 0 && first + g < n_hops) totals[first + g] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// A2: per-clip exclusive scan of the hop totals (in place: totals -> prefix) and the frame-rate
// phase  phase_frames[b,h] = fl32(2*pi) * wrap((prefix[h] + f0[h]) / sr + init/2/pi).
// One CTA of 1024 threads per clip; each thread owns a contiguous chunk of hops.
// ---------------------------------------------------------------------------------------------
template <typename Acc>
__device__ __forceinline__ Acc block_exclusive_scan(Acc v, Acc* smem /* >= 32 */, Acc& block_total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    Acc inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Acc t = __shfl_up_sync(kFullMask, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        Acc w = (lane < nw) ? smem[lane] : Acc(0);
        Acc winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const Acc t = __shfl_up_sync(kFullMask, winc, d);
            if (lane >= d) winc += t;
        }
        smem[lane] = winc - w;            // exclusive warp offsets
        if (lane == 31) smem[32] = winc;  // grand total
    }
    __syncthreads();
    const Acc off = smem[wid];
    block_total = smem[32];
    __syncthreads();
    return off + (inc - v);
}

// Streaming: the prefix (sum of upsampled f0 in Hz*samples) a previous block ended with; must not
// alias this call's `prefix` output.  Only its value modulo sr matters for the phase; it is folded
// once it exceeds 2^40 (hours of audio) so that an endless stream never runs out of fp64 mantissa,
// and left untouched below that so that a stream of blocks stays bit-identical to one call over
// the concatenated frames.
__device__ __forceinline__ double stream_carry(const double* __restrict__ carry, int64_t carry_stride, int b,
                                               double inv_sr) {
    if (!carry) return 0.0;
    double c = carry[(int64_t)b * carry_stride];
    if (fabs(c) >= 0x1p40) {
        const double sr = 1.0 / inv_sr, sri = rint(sr);
        c = fmod(c, fabs(sr - sri) < 1e-9 * sr ? sri : sr);
    }
    return c;
}

__global__ void __launch_bounds__(1024) phase_scan_kernel(const float* __restrict__ f0_frames, int64_t fB,
                                                          int64_t fF, int F, double inv_sr,
                                                          const float* __restrict__ initial_phase,
                                                          const double* __restrict__ carry, int64_t carry_stride,
                                                          double* __restrict__ prefix /* in: totals */,
                                                          float* __restrict__ phase_frames) {
    __shared__ double sm[33];
    const int b = blockIdx.x;
    double* pf = prefix + (int64_t)b * F;
    const float* row = f0_frames + (int64_t)b * fB;
    const int per = (F + blockDim.x - 1) / blockDim.x;
    const int h0 = min(F, (int)threadIdx.x * per), h1 = min(F, h0 + per);
    cudaGridDependencySynchronize();          // launched with programmatic stream serialisation: totals come from A1
    double s = 0.0;
    for (int h = h0; h < h1; ++h) s += pf[h];
    double total;
    double run = block_exclusive_scan<double>(s, sm, total);
    // initial_phase/2/pi rotations (core.py:45), carried inside the prefix in Hz*samples
    run += initial_phase ? (((double)initial_phase[b] / 2.0) / 3.14159265358979323846) / inv_sr : 0.0;
    run += stream_carry(carry, carry_stride, b, inv_sr);
    for (int h = h0; h < h1; ++h) {
        const double t = pf[h];
        pf[h] = run;
        const double c = (run + (double)__ldg(row + (int64_t)h * fF)) * inv_sr;
        phase_frames[(int64_t)b * F + h] = __fmul_rn(DDSP_TWO_PI_F, wrap_rot(c));
        run += t;
    }
}

// Small-problem variant: one CTA per clip does A1 and A2 in one launch (streaming sizes).
__global__ void __launch_bounds__(1024) phase_fused_kernel(const float* __restrict__ f0_frames, int64_t fB,
                                                           int64_t fF, int F, double inv_sr,
                                                           const float* __restrict__ initial_phase,
                                                           const double* __restrict__ carry, int64_t carry_stride,
                                                           double* __restrict__ prefix,
                                                           float* __restrict__ phase_frames) {
    __shared__ double sm[33];
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* pf = prefix + (int64_t)b * F;
    const float* row = f0_frames + (int64_t)b * fB;
    for (int h = wid; h < F; h += nw) {
        const float x0 = __ldg(row + (int64_t)h * fF);
        const float x1 = __ldg(row + (int64_t)min(h + 1, F - 1) * fF);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < kHop / 32; ++i) s += (double)lerp_torch(x0, x1, (float)(lane + 32 * i) * (1.0f / kHop));
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(kFullMask, s, d);
        if (lane == 0) pf[h] = s;
    }
    __syncthreads();
    const int per = (F + blockDim.x - 1) / blockDim.x;
    const int h0 = min(F, (int)threadIdx.x * per), h1 = min(F, h0 + per);
    double s = 0.0;
    for (int h = h0; h < h1; ++h) s += pf[h];
    double total;
    double run = block_exclusive_scan<double>(s, sm, total);
    // initial_phase/2/pi rotations (core.py:45), carried inside the prefix in Hz*samples
    run += initial_phase ? (((double)initial_phase[b] / 2.0) / 3.14159265358979323846) / inv_sr : 0.0;
    run += stream_carry(carry, carry_stride, b, inv_sr);
    for (int h = h0; h < h1; ++h) {
        const double t = pf[h];
        pf[h] = run;
        const double c = (run + (double)__ldg(row + (int64_t)h * fF)) * inv_sr;
        phase_frames[(int64_t)b * F + h] = __fmul_rn(DDSP_TWO_PI_F, wrap_rot(c));
        run += t;
    }
}

// ---------------------------------------------------------------------------------------------
// One hop of per-sample rotation, computed by a warp: lane owns 16 consecutive samples.
// f[i] = upsampled f0, rot[i] = wrapped rotation (fp32) of sample 16*lane + i of hop h.
// ---------------------------------------------------------------------------------------------
// Packed form: sample pairs (2j, 2j+1) share fp32x2 registers; the interpolation weights are exact
// multiples of 1/512, so lambda_0 + i/512 equals (16 lane + i)/512 bit for bit.
__device__ __forceinline__ void hop_rotation2(float x0, float x1, double base, double inv_sr, int lane,
                                              float2 (&f2)[8], float2 (&rot2)[8]) {
    double sl[16];
    double s = 0.0;
    const float lam0 = (float)(16 * lane) * (1.0f / kHop);
    float2 lam = make_float2(lam0, lam0 + 1.0f / kHop);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // lerp_torch on both halves: fma(1 - lambda, x0, fl(lambda * x1))
        f2[j] = fma2(sub2(bc2(1.0f), lam), bc2(x0), mul2(lam, bc2(x1)));
        lam = add2(lam, bc2(2.0f / kHop));
        s += (double)f2[j].x;
        sl[2 * j] = s;
        s += (double)f2[j].y;
        sl[2 * j + 1] = s;
    }
    double inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(kFullMask, inc, d);
        if (lane >= d) inc += t;
    }
    const double off = (base + (inc - s)) * inv_sr;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        rot2[j] = make_float2(wrap_rot(fma(sl[2 * j], inv_sr, off)), wrap_rot(fma(sl[2 * j + 1], inv_sr, off)));
}

// Fixed-point variant for the fused CombSubFast kernels (forward and gradient).  The lane's first sample
// takes its rotation from the exact fp64 hop prefix plus the closed-form sum of the interpolated f0 over
// the samples before the lane; inside the lane the rotation advances in 32-bit fixed point (unit 2^-32
// rotation; integer wrap-around is the reference's `rot - round(rot)`, core.py:46) by rn(f * 2^32/sr)
// per sample.  What is ignored: the fp32 roundings of the individual upsampled samples inside ONE hop
// (<= 512 * 2^-25 * f0 Hz*samples, i.e. < 3e-7 rotations worst case, ~1e-8 typical) -- they do not
// accumulate because `base` is the exact prefix of every hop -- and the 2^-24 relative rounding of the 16
// increments.  In the sinc argument sr*rot/f0 both stay below 1e-6, under the 3e-5 fp32 spacing the
// reference's own argument has at low f0.  Returns f (bit-exact upsample) and rot * 2^32 as fp32.
__device__ __forceinline__ void hop_rotation_q32(float x0, float x1, double base, double inv_sr, int lane,
                                                 float2 (&f2)[8], float2 (&roti2)[8]) {
    const double n = (double)(16 * lane);
    const double dx = (double)x1 - (double)x0;
    // sum_{k < n} (x0 + dx k/512) = n x0 + dx n(n-1)/1024   (n(n-1) and dx/1024 are exact)
    const double excl = fma(dx * (1.0 / 1024.0), n * (n - 1.0), n * (double)x0);
    double c = (base + excl) * inv_sr;
    c -= rint(c);
    uint32_t p = (uint32_t)__double2ll_rn(c * 4294967296.0);
    const float k32 = (float)(4294967296.0 * inv_sr);
    const float lam0 = (float)(16 * lane) * (1.0f / kHop);
    float2 lam = make_float2(lam0, lam0 + 1.0f / kHop);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        f2[j] = fma2(sub2(bc2(1.0f), lam), bc2(x0), mul2(lam, bc2(x1)));      // lerp_torch on both halves
        lam = add2(lam, bc2(2.0f / kHop));
        const float2 d = mul2(f2[j], bc2(k32));
        p += __float2uint_rn(d.x);
        const float r0 = (float)(int32_t)p;
        p += __float2uint_rn(d.y);
        roti2[j] = make_float2(r0, (float)(int32_t)p);
    }
}

__device__ __forceinline__ void hop_rotation(float x0, float x1, double base, double inv_sr, int lane,
                                             float (&f)[16], float (&rot)[16]) {
    float2 f2[8], rot2[8];
    hop_rotation2(x0, x1, base, inv_sr, lane, f2, rot2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        f[2 * j] = f2[j].x; f[2 * j + 1] = f2[j].y;
        rot[2 * j] = rot2[j].x; rot[2 * j + 1] = rot2[j].y;
    }
}

// A3 (Sins): full-rate phase = fl32(2*pi)*rot (vocoder.py:392), one warp per hop.
__global__ void __launch_bounds__(256) phase_full_kernel(const float* __restrict__ f0_frames, int64_t fB,
                                                         int64_t fF, int B, int F, double inv_sr,
                                                         const float* __restrict__ initial_phase,
                                                         const double* __restrict__ prefix,
                                                         float* __restrict__ phase_full) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= (int64_t)B * F) return;
    const int b = (int)(warp / F), h = (int)(warp % F);
    const float* row = f0_frames + (int64_t)b * fB;
    const float x0 = __ldg(row + (int64_t)h * fF);
    const float x1 = __ldg(row + (int64_t)min(h + 1, F - 1) * fF);
    float f[16], rot[16];
    hop_rotation(x0, x1, prefix[warp], inv_sr, lane, f, rot);
    float4* out = reinterpret_cast<float4*>(phase_full + warp * kHop + 16 * lane);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        out[i] = make_float4(__fmul_rn(DDSP_TWO_PI_F, rot[4 * i]), __fmul_rn(DDSP_TWO_PI_F, rot[4 * i + 1]),
                             __fmul_rn(DDSP_TWO_PI_F, rot[4 * i + 2]), __fmul_rn(DDSP_TWO_PI_F, rot[4 * i + 3]));
}

// ---------------------------------------------------------------------------------------------
// Standalone upsample (core.py:7-21): (B,F,C) strided -> (B,F*factor,C) contiguous.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_kernel(const float* __restrict__ x, int64_t sB, int64_t sF,
                                                       int64_t sC, int B, int F, int C, int factor,
                                                       float rwidth, float* __restrict__ y) {
    const int64_t total = (int64_t)B * F * factor * C;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const int64_t bt = idx / C;
        const int64_t t = bt % ((int64_t)F * factor);
        const int b = (int)(bt / ((int64_t)F * factor));
        // torch: w1r = rwidth * w2 (fp32), w1 = (int)w1r, lambda = w1r - w1      (UpSampleLinear1d.cu)
        const float w1r = __fmul_rn(rwidth, (float)t);
        int m = (int)w1r;
        const float lam = __fsub_rn(w1r, (float)m);
        const int m1 = min(m + 1, F - 1);          // x[F] := x[F-1] (core.py:17 hold-last)
        m = min(m, F - 1);
        const float* base = x + (int64_t)b * sB + (int64_t)c * sC;
        y[idx] = lerp_torch(__ldg(base + (int64_t)m * sF), __ldg(base + (int64_t)m1 * sF), lam);
    }
}

// ---------------------------------------------------------------------------------------------
// Standalone remove_above_fmax (core.py:24-28), bit-exact fp32.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) remove_above_fmax_kernel(const float* __restrict__ amp, int64_t aB,
                                                                int64_t aF, const float* __restrict__ pitch,
                                                                int64_t pB, int64_t pF, float fmax,
                                                                int level_start, int B, int F, int K,
                                                                float* __restrict__ out) {
    const int64_t total = (int64_t)B * F * K;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % K);
        const int64_t bf = idx / K;
        const int f = (int)(bf % F), b = (int)(bf / F);
        const float p = __ldg(pitch + (int64_t)b * pB + (int64_t)f * pF);
        const float pk = __fmul_rn(p, (float)(k + level_start));
        const float aa = __fadd_rn((pk < fmax) ? 1.0f : 0.0f, 1e-7f);
        out[idx] = __fmul_rn(__ldg(amp + (int64_t)b * aB + (int64_t)f * aF + k), aa);
    }
}

// ---------------------------------------------------------------------------------------------
// Standalone fo_to_rot (core.py:31-51) over an arbitrary (B,T) fp32 contour: three-kernel
// chunked scan (chunk sums -> per-row scan of chunk sums -> apply).
// ---------------------------------------------------------------------------------------------
constexpr int kRotChunk = 2048;   // samples per CTA chunk (256 threads x 8)

template <typename Acc>
__global__ void __launch_bounds__(256) rot_chunk_sums_kernel(const float* __restrict__ fo, int64_t T,
                                                             int nchunks, Acc sr, Acc* __restrict__ sums) {
    __shared__ Acc sm[33];
    const int b = blockIdx.y, ch = blockIdx.x;
    const float* row = fo + (int64_t)b * T;
    const int64_t t0 = (int64_t)ch * kRotChunk + (int64_t)threadIdx.x * 8;
    Acc s = Acc(0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t t = t0 + i;
        if (t < T) s += (Acc)row[t] / sr;
    }
    Acc total;
    (void)block_exclusive_scan<Acc>(s, sm, total);
    if (threadIdx.x == 0) sums[(int64_t)b * nchunks + ch] = total;
}

template <typename Acc>
__global__ void __launch_bounds__(1024) rot_scan_sums_kernel(int nchunks, Acc* __restrict__ sums) {
    __shared__ Acc sm[33];
    Acc* row = sums + (int64_t)blockIdx.x * nchunks;
    const int per = (nchunks + blockDim.x - 1) / blockDim.x;
    const int c0 = min(nchunks, (int)threadIdx.x * per), c1 = min(nchunks, c0 + per);
    Acc s = Acc(0);
    for (int c = c0; c < c1; ++c) s += row[c];
    Acc total;
    Acc run = block_exclusive_scan<Acc>(s, sm, total);
    for (int c = c0; c < c1; ++c) { const Acc t = row[c]; row[c] = run; run += t; }
}

template <typename Acc>
__global__ void __launch_bounds__(256) rot_apply_kernel(const float* __restrict__ fo, int64_t T, int nchunks,
                                                        Acc sr, const float* __restrict__ initial_phase,
                                                        const Acc* __restrict__ sums, float* __restrict__ rot) {
    __shared__ Acc sm[33];
    const int b = blockIdx.y, ch = blockIdx.x;
    const float* row = fo + (int64_t)b * T;
    const int64_t t0 = (int64_t)ch * kRotChunk + (int64_t)threadIdx.x * 8;
    Acc v[8];
    Acc s = Acc(0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t t = t0 + i;
        v[i] = (t < T) ? (Acc)row[t] / sr : Acc(0);     // core.py:43  _fo / sr
        s += v[i];
        v[i] = s;
    }
    Acc total;
    const Acc off = block_exclusive_scan<Acc>(s, sm, total) + sums[(int64_t)b * nchunks + ch];
    Acc init = Acc(0);
    if (initial_phase) init = (Acc)initial_phase[b] / Acc(2) / Acc(3.14159265358979323846);   // core.py:45
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t t = t0 + i;
        if (t < T) {
            const Acc c = off + v[i] + init;
            rot[(int64_t)b * T + t] = (float)(c - rint(c));   // core.py:46,49
        }
    }
}

}  // namespace ddsp
