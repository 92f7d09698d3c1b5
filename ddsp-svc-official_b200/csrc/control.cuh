// control.cuh -- fused elementwise stages of the control network (SURVEY.md section 8f rank 1:
// ddsp/unit2control.py + ddsp/pcmer.py).  The GEMMs of the network stay library calls (cuBLAS, fp32);
// these two kernels replace the chains of broadcast/elementwise/reduction passes between them, each of
// which streamed a (B,H,N,266) or (B,N,1024) tensor through HBM:
//   * performer_features_kernel: the FAVOR+ softmax-kernel feature map (pcmer.py:124-160) from the raw
//     projections: one read of `dash`, one write of the features (was ~10 passes);
//   * glu_dwconv_silu_kernel: GLU -> depthwise Conv1d(k=31, 'same') -> SiLU of the conformer
//     convolution module (pcmer.py:41-63) in channels-last layout, no transposes.
// They are streaming kernels in fp32; the one-pass feature kernel and the conv stage use full-precision
// expf, the projection-fused kernel (whose exp argument already carries the fp32 rounding of a 64-term
// dot product) uses __expf.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace ddsp {

constexpr int kPerfDim = 64;       // head dimension (pcmer.py:166 dim_head = 64)

// One warp per row r = (b, n, h) of the projections.
//   dash : (B, N, H, M) contiguous = (normalizer * x) @ projection^T            pcmer.py:146
//   x    : (B, N, H, 64) contiguous (the to_q / to_k output before the head split)
//   out  : (B, H, N, M) contiguous                                                 (head-major, as the einsums want)
// query: ratio * (exp(dash - diag - max_j dash) + eps);  key: ratio * exp(dash - diag + eps)   pcmer.py:154-157
// with diag = sum(x^2)/2 * normalizer^2                                                        pcmer.py:149-152
template <bool IS_QUERY>
__global__ void __launch_bounds__(256) performer_features_kernel(const float* __restrict__ dash,
                                                                 const float* __restrict__ x, float* __restrict__ out,
                                                                 int B, int N, int H, int M, float normalizer2_half,
                                                                 float ratio, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t rows = (int64_t)B * N * H;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const int h = (int)(r % H);
        const int64_t bn = r / H;
        const int n = (int)(bn % N), b = (int)(bn / N);
        const float2 xv = __ldg(reinterpret_cast<const float2*>(x + r * kPerfDim) + lane);
        float ss = fmaf(xv.x, xv.x, xv.y * xv.y);
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float diag = ss * normalizer2_half;
        const float* drow = dash + r * M;
        float v[12];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int j = lane + 32 * i;
            v[i] = (j < M) ? __ldg(drow + j) : -INFINITY;
            mx = fmaxf(mx, v[i]);
        }
        float shift = diag;
        if (IS_QUERY) {
#pragma unroll
            for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        float* orow = out + (((int64_t)b * H + h) * N + n) * M;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int j = lane + 32 * i;
            if (j < M) {
                float y;
                if (IS_QUERY) y = ratio * (expf(v[i] - shift - mx) + eps);
                else y = ratio * expf(v[i] - shift + eps);
                orow[j] = y;
            }
        }
    }
}

// Projection + feature map in one kernel: dash never touches HBM.
//   x    : (B, N, H, 64) contiguous;  proj : (M, 64) random-feature matrix (pcmer.py:80-120 buffer)
//   out  : (B, H, N, M)
// dash[r][j] = sum_d fl(normalizer * x[r][d]) * proj[j][d]  (fp32 FMA chain over d), then the feature
// map above.  A CTA keeps proj^T (64 x 288, zero padded) in shared memory; a warp owns kPpfRows rows
// at a time (72 accumulators per lane: 8 rows x 9 column groups), x is staged per warp in shared
// memory, transposed and row-pair interleaved, and read back as broadcast float4s: 9 + 2 shared loads
// per 36 packed FFMA2 (= 72 FMAs; the projection value is the broadcast operand).
constexpr int kPpfRows = 8, kPpfCols = 288, kPpfWarps = 8;
constexpr int kPpfSmemBytes = (kPerfDim * kPpfCols + kPpfWarps * kPpfRows * kPerfDim) * 4;

template <bool IS_QUERY>
__global__ void __launch_bounds__(kPpfWarps * 32, 2) performer_project_features_kernel(
    const float* __restrict__ x, const float* __restrict__ xbias, const float* __restrict__ proj,
    float* __restrict__ out, int B, int N, int H, int M, float normalizer, float normalizer2_half, float ratio,
    float eps) {
    extern __shared__ __align__(16) float ppf_smem[];
    float* PT = ppf_smem;                                       // [64][288]: PT[d][j] = proj[j][d]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* xs = ppf_smem + kPerfDim * kPpfCols + wid * (kPpfRows * kPerfDim);   // [8][64] scaled rows of this warp
    for (int e = threadIdx.x; e < kPerfDim * kPpfCols; e += kPpfWarps * 32) {
        const int j = e / kPerfDim, d = e % kPerfDim;           // read proj row-major (coalesced), write transposed
        PT[d * kPpfCols + j] = (j < M) ? __ldg(proj + (int64_t)j * kPerfDim + d) : 0.0f;
    }
    __syncthreads();
    const int64_t rows = (int64_t)B * N * H;
    const int64_t groups = (rows + kPpfRows - 1) / kPpfRows;
    for (int64_t g = (int64_t)blockIdx.x * kPpfWarps + wid; g < groups; g += (int64_t)gridDim.x * kPpfWarps) {
        const int64_t r0 = g * kPpfRows;
        // stage 8 rows (512 floats): lane holds 16 consecutive floats = a quarter of row lane/4
        float ss = 0.0f;
        {
            const int64_t row = r0 + (lane >> 2);
            const float4* src = reinterpret_cast<const float4*>(x + row * kPerfDim) + (lane & 3) * 4;
            // optional bias of the producing Linear (H*64 entries, head h = row % H), added here so that the
            // GEMM runs without a separate bias epilogue
            const float4* bsrc = xbias ? reinterpret_cast<const float4*>(xbias + (row % H) * kPerfDim) + (lane & 3) * 4 : nullptr;
            // staged transposed and row-pair interleaved: xs[(d*4 + rp)*2 + e] = x[2 rp + e][d], so that the main
            // loop reads the same d of two rows as one packed fp32x2 operand
            float* dst = xs + ((lane & 3) * 16 * 4 + (lane >> 3)) * 2 + ((lane >> 2) & 1);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < rows) {
                    t = __ldg(src + v);
                    if (bsrc) {
                        const float4 bb = __ldg(bsrc + v);
                        t.x += bb.x; t.y += bb.y; t.z += bb.z; t.w += bb.w;
                    }
                }
                ss = fmaf(t.x, t.x, fmaf(t.y, t.y, fmaf(t.z, t.z, fmaf(t.w, t.w, ss))));
                dst[(4 * v + 0) * 8] = normalizer * t.x;
                dst[(4 * v + 1) * 8] = normalizer * t.y;
                dst[(4 * v + 2) * 8] = normalizer * t.z;
                dst[(4 * v + 3) * 8] = normalizer * t.w;
            }
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);           // lanes 4r..4r+3 hold |x_r|^2
        }
        __syncwarp();
        float2 acc[kPpfRows / 2][9];                             // .x: row 2 rp, .y: row 2 rp + 1
#pragma unroll
        for (int rp = 0; rp < kPpfRows / 2; ++rp)
#pragma unroll
            for (int i = 0; i < 9; ++i) acc[rp][i] = make_float2(0.0f, 0.0f);
        const float4* xs4 = reinterpret_cast<const float4*>(xs);
#pragma unroll 4
        for (int d = 0; d < kPerfDim; ++d) {
            const float4 xa = xs4[2 * d], xb = xs4[2 * d + 1];   // rows (0,1),(2,3) and (4,5),(6,7) at this d
            float pv[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) pv[i] = PT[d * kPpfCols + lane + 32 * i];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                acc[0][i] = fma2(make_float2(xa.x, xa.y), bc2(pv[i]), acc[0][i]);
                acc[1][i] = fma2(make_float2(xa.z, xa.w), bc2(pv[i]), acc[1][i]);
                acc[2][i] = fma2(make_float2(xb.x, xb.y), bc2(pv[i]), acc[2][i]);
                acc[3][i] = fma2(make_float2(xb.z, xb.w), bc2(pv[i]), acc[3][i]);
            }
        }
        __syncwarp();                                            // xs is rewritten by the next group
#define ACC(r, i) ((r) & 1 ? acc[(r) >> 1][i].y : acc[(r) >> 1][i].x)
        // feature map as one FFMA + ex2 per element:  ratio*exp(a - c) = 2^(a*log2e + k),
        // k = log2(ratio) - c*log2e per row (ex2.approx of a 64-term fp32 dot product: same error order)
        const float log2_ratio = log2f(ratio);
        const bool tail_ok = lane + 256 < M;                 // column group i = 8 is the only partial one (M > 256)
#pragma unroll
        for (int r = 0; r < kPpfRows; ++r) {
            const int64_t row = r0 + r;
            const float diag = __shfl_sync(0xffffffffu, ss, 4 * r) * normalizer2_half;
            float mx = 0.0f;
            if (IS_QUERY) {
                mx = tail_ok ? ACC(r, 8) : -INFINITY;
#pragma unroll
                for (int i = 0; i < 8; ++i) mx = fmaxf(mx, ACC(r, i));
#pragma unroll
                for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            const float k = IS_QUERY ? fmaf(-(diag + mx), 1.4426950408889634f, log2_ratio)
                                     : fmaf(eps - diag, 1.4426950408889634f, log2_ratio);
            const float add = IS_QUERY ? ratio * eps : 0.0f;
            if (row < rows) {
                const int h = (int)(row % H);
                const int64_t bn = row / H;
                const int n = (int)(bn % N), b = (int)(bn / N);
                float* orow = out + (((int64_t)b * H + h) * N + n) * M + lane;
#pragma unroll
                for (int i = 0; i < 8; ++i) orow[32 * i] = ex2_approx(fmaf(ACC(r, i), 1.4426950408889634f, k)) + add;
                if (tail_ok) orow[256] = ex2_approx(fmaf(ACC(r, 8), 1.4426950408889634f, k)) + add;
            }
        }
#undef ACC
    }
}

// Whole non-causal Performer attention of one (clip, head) for streaming-sized blocks
// (pcmer.py:69-78,124-160,191-251 after the q/k/v projections): feature maps of q and k, k_sum,
// context = k'^T v, out = (q' context) / (q' k_sum + 1e-8), written head-merged as (B, N, H*64).
// Replaces ~16 small kernels per layer (scales, two projection GEMMs, two feature maps, sum, gemv,
// reciprocal, two batched GEMMs, scaling, two layout copies) whose launch latencies dominate the
// graph-replay time of a GUI block.  One CTA of 288 threads per (head, clip); thread j owns feature
// column j (< M <= 288) in the projection and context phases.  Everything is laid out for latency:
// the CTA is alone on its SM, so loads are batched and dependent chains are split.
//   q, k, v : (B, N, H*64) contiguous, optional biases (H*64) of the producing Linears added on load
//   proj    : (M, 64);  out : (B, N, H*64)
constexpr int kPasThreads = 288, kPasRows = 8, kPasWarps = kPasThreads / 32;
constexpr int kPasPStride = kPerfDim + 1;                      // proj rows padded: thread j reads P[j*65 + d] conflict-free
constexpr int kPasSmemFloats = kPpfCols * kPasPStride          // P[j][d]
                               + kPpfCols * kPasPStride        // ctx[j][0..63], ctx[j][64] = k_sum[j]
                               + kPasRows * kPpfCols           // feature tile
                               + 2 * kPasRows * kPerfDim       // x tile (scaled), v tile
                               + 3 * kPasRows + kPasWarps * kPasRows;   // per-row diag / max / denominators, warp maxima
constexpr int kPasSmemBytes = kPasSmemFloats * 4;

// Feature tile of rows n0 .. n0+rows-1 of `src` (q or k of this head) into ft[r][j].
__device__ __forceinline__ void pas_features(const float* __restrict__ src, const float* __restrict__ bias, int h,
                                             int64_t rs, int n0, int rows, const float* P, float* xt, float* ft,
                                             float* rowstat, float* wmax, int M, bool is_query, float ratio, float eps) {
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    // stage the tile's rows scaled by normalizer = 64^-0.25 (pcmer.py:137,146): 512 floats, 2 per thread (first 256)
    if (t < kPasRows * kPerfDim / 2) {
        const int r = t / (kPerfDim / 2), d = 2 * (t % (kPerfDim / 2));
        float2 v = make_float2(0.0f, 0.0f);
        if (r < rows) {
            v = __ldg(reinterpret_cast<const float2*>(src + (int64_t)(n0 + r) * rs + h * kPerfDim + d));
            if (bias) { v.x += __ldg(bias + h * kPerfDim + d); v.y += __ldg(bias + h * kPerfDim + d + 1); }
        }
        *reinterpret_cast<float2*>(xt + r * kPerfDim + d) = make_float2(0.35355339059327373f * v.x, 0.35355339059327373f * v.y);
    }
    __syncthreads();
    if (wid < kPasRows) {                                   // warp r: |x_r|^2 / 16 = |normalizer x_r|^2 / 2  (pcmer.py:149-152)
        const float a = xt[wid * kPerfDim + lane], b = xt[wid * kPerfDim + 32 + lane];
        float ss = fmaf(a, a, b * b);
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) rowstat[wid] = ss * 0.5f;
    }
    // dash[r][j] = sum_d (normalizer * x[r][d]) * proj[j][d], thread j = t; 4 values of d per step
    float acc[kPasRows];
#pragma unroll
    for (int r = 0; r < kPasRows; ++r) acc[r] = 0.0f;
    const float* prow = P + t * kPasPStride;
#pragma unroll 4
    for (int d = 0; d < kPerfDim; d += 4) {
        const float p0 = prow[d], p1 = prow[d + 1], p2 = prow[d + 2], p3 = prow[d + 3];
#pragma unroll
        for (int r = 0; r < kPasRows; ++r) {
            const float4 xv = *reinterpret_cast<const float4*>(xt + r * kPerfDim + d);     // broadcast
            acc[r] = fmaf(xv.w, p3, fmaf(xv.z, p2, fmaf(xv.y, p1, fmaf(xv.x, p0, acc[r]))));
        }
    }
    if (is_query) {                                         // row maxima over the M valid columns: warp, then CTA
#pragma unroll
        for (int r = 0; r < kPasRows; ++r) {
            float m = (t < M) ? acc[r] : -INFINITY;
#pragma unroll
            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) wmax[wid * kPasRows + r] = m;
        }
    }
    __syncthreads();                                        // rowstat / wmax visible
#pragma unroll
    for (int r = 0; r < kPasRows; ++r) {
        float y = 0.0f;
        if (t < M && r < rows) {
            if (is_query) {
                float mm = wmax[r];
#pragma unroll
                for (int w = 1; w < kPasWarps; ++w) mm = fmaxf(mm, wmax[w * kPasRows + r]);
                y = ratio * (expf(acc[r] - rowstat[r] - mm) + eps);
            } else {
                y = ratio * expf(acc[r] - rowstat[r] + eps);
            }
        }
        ft[r * kPpfCols + t] = y;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPasThreads) performer_attention_small_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const float* __restrict__ qb, const float* __restrict__ kb, const float* __restrict__ vb,
    const float* __restrict__ proj, float* __restrict__ out, int N, int H, int M, int64_t rs, float ratio, float eps) {
    extern __shared__ __align__(16) float pas_smem[];
    float* P = pas_smem;
    float* ctx = P + kPpfCols * kPasPStride;
    float* ft = ctx + kPpfCols * kPasPStride;
    float* xt = ft + kPasRows * kPpfCols;
    float* vt = xt + kPasRows * kPerfDim;
    float* rowstat = vt + kPasRows * kPerfDim;
    float* wmax = rowstat + 3 * kPasRows;
    const int t = threadIdx.x, h = blockIdx.x, b = blockIdx.y;
    const int64_t clip = (int64_t)b * N * rs;                 // q / k / v: rs elements between consecutive frames
    const int64_t oclip = (int64_t)b * N * H * kPerfDim;      // out: head-merged, contiguous
    {   // proj (M x 64, contiguous) -> P[j*65 + d]; all loads of a thread in flight before the first store
        const int n4 = M * kPerfDim / 4;
        constexpr int kPer = (kPpfCols * kPerfDim / 4 + kPasThreads - 1) / kPasThreads;      // 16
        float4 buf[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int f = t + i * kPasThreads;
            buf[i] = (f < n4) ? __ldg(reinterpret_cast<const float4*>(proj) + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int f = t + i * kPasThreads;
            if (f < kPpfCols * kPerfDim / 4) {
                float* dst = P + (f / (kPerfDim / 4)) * kPasPStride + 4 * (f % (kPerfDim / 4));
                dst[0] = buf[i].x; dst[1] = buf[i].y; dst[2] = buf[i].z; dst[3] = buf[i].w;
            }
        }
    }
    __syncthreads();
    // ---- keys: context[j][e] = sum_n k'[n][j] v[n][e],  k_sum[j] = sum_n k'[n][j] ----------------
    float c[kPerfDim + 1];
#pragma unroll
    for (int e = 0; e <= kPerfDim; ++e) c[e] = 0.0f;
    for (int n0 = 0; n0 < N; n0 += kPasRows) {
        const int rows = min(kPasRows, N - n0);
        if (t >= kPasThreads - kPasRows * kPerfDim / 16) {   // the last 32 threads stage the v tile (512 floats, float4 each x4)
            const int u = t - (kPasThreads - kPasRows * kPerfDim / 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int f = u + 32 * i, r = f / (kPerfDim / 4), d = 4 * (f % (kPerfDim / 4));
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < rows) {
                    x = __ldg(reinterpret_cast<const float4*>(v + clip + (int64_t)(n0 + r) * rs + h * kPerfDim + d));
                    if (vb) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(vb + h * kPerfDim + d));
                        x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
                    }
                }
                *reinterpret_cast<float4*>(vt + r * kPerfDim + d) = x;
            }
        }
        pas_features(k + clip, kb, h, rs, n0, rows, P, xt, ft, rowstat, wmax, M, false, ratio, eps);
#pragma unroll 2
        for (int r = 0; r < kPasRows; ++r) {                // rows beyond the tile hold zero features
            const float kf = ft[r * kPpfCols + t];
            const float4* vr = reinterpret_cast<const float4*>(vt + r * kPerfDim);
#pragma unroll
            for (int e4 = 0; e4 < kPerfDim / 4; ++e4) {
                const float4 vv = vr[e4];
                c[4 * e4 + 0] = fmaf(kf, vv.x, c[4 * e4 + 0]);
                c[4 * e4 + 1] = fmaf(kf, vv.y, c[4 * e4 + 1]);
                c[4 * e4 + 2] = fmaf(kf, vv.z, c[4 * e4 + 2]);
                c[4 * e4 + 3] = fmaf(kf, vv.w, c[4 * e4 + 3]);
            }
            c[kPerfDim] += kf;
        }
        __syncthreads();
    }
#pragma unroll
    for (int e = 0; e <= kPerfDim; ++e) ctx[t * kPasPStride + e] = c[e];
    __syncthreads();
    // ---- queries: out[n][e] = sum_j q'[n][j] context[j][e] / (sum_j q'[n][j] k_sum[j] + 1e-8) ----
    for (int n0 = 0; n0 < N; n0 += kPasRows) {
        const int rows = min(kPasRows, N - n0);
        pas_features(q + clip, qb, h, rs, n0, rows, P, xt, ft, rowstat, wmax, M, true, ratio, eps);
        // 8 rows x 65 columns (64 outputs + the denominator): thread -> (row, column), four partial sums
        for (int o = t; o < kPasRows * kPasPStride; o += kPasThreads) {
            const int r = o / kPasPStride, e = o % kPasPStride;
            const float* fr = ft + r * kPpfCols;
            const float* cc = ctx + e;
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
            int j = 0;
#pragma unroll 4
            for (; j + 4 <= M; j += 4) {                    // (unrolled x4: 16 loads in flight per step)
                const float4 f4 = *reinterpret_cast<const float4*>(fr + j);
                a0 = fmaf(f4.x, cc[j * kPasPStride], a0);
                a1 = fmaf(f4.y, cc[(j + 1) * kPasPStride], a1);
                a2 = fmaf(f4.z, cc[(j + 2) * kPasPStride], a2);
                a3 = fmaf(f4.w, cc[(j + 3) * kPasPStride], a3);
            }
            for (; j < M; ++j) a0 = fmaf(fr[j], cc[j * kPasPStride], a0);
            const float a = (a0 + a1) + (a2 + a3);
            if (e == kPerfDim) rowstat[2 * kPasRows + r] = a;
            else xt[r * kPerfDim + e] = a;                  // xt is free again: numerators
        }
        __syncthreads();
        for (int o = t; o < rows * kPerfDim; o += kPasThreads) {
            const int r = o / kPerfDim;
            out[oclip + ((int64_t)(n0 + r) * H + h) * kPerfDim + (o % kPerfDim)] = xt[o] * (1.0f / (rowstat[2 * kPasRows + r] + 1e-8f));
        }
        __syncthreads();
    }
}

// The same attention for longer blocks, spread over the chip: one CTA per (head, clip, 8-frame tile).
//   1. performer_context_partial_kernel: key features of the tile and its 288 x 65 partial context
//      (last column: partial k_sum) -> workspace[(clip*H + head)*tiles + tile]
//   2. performer_context_reduce_kernel: sums the tiles in a fixed order (deterministic) -> context
//   3. performer_output_kernel: query features of the tile, (q' context) / (q' k_sum + 1e-8), head merge
__device__ __forceinline__ void pas_stage_projection(const float* __restrict__ proj, int M, float* P) {
    const int t = threadIdx.x;
    const int n4 = M * kPerfDim / 4;
    constexpr int kPer = (kPpfCols * kPerfDim / 4 + kPasThreads - 1) / kPasThreads;      // 16
    float4 buf[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const int f = t + i * kPasThreads;
        buf[i] = (f < n4) ? __ldg(reinterpret_cast<const float4*>(proj) + f) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const int f = t + i * kPasThreads;
        if (f < kPpfCols * kPerfDim / 4) {
            float* dst = P + (f / (kPerfDim / 4)) * kPasPStride + 4 * (f % (kPerfDim / 4));
            dst[0] = buf[i].x; dst[1] = buf[i].y; dst[2] = buf[i].z; dst[3] = buf[i].w;
        }
    }
}

constexpr int kPctxFloats = kPpfCols * kPasPStride;            // one (partial) context incl. the k_sum column

__global__ void __launch_bounds__(kPasThreads) performer_context_partial_kernel(
    const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ kb, const float* __restrict__ vb,
    const float* __restrict__ proj, float* __restrict__ partial, int N, int H, int M, int64_t rs, float ratio, float eps) {
    extern __shared__ __align__(16) float pas_smem[];
    float* P = pas_smem;
    float* ft = P + kPpfCols * kPasPStride;
    float* xt = ft + kPasRows * kPpfCols;
    float* vt = xt + kPasRows * kPerfDim;
    float* rowstat = vt + kPasRows * kPerfDim;
    float* wmax = rowstat + 3 * kPasRows;
    const int t = threadIdx.x, h = blockIdx.x, b = blockIdx.y, tile = blockIdx.z;
    const int64_t clip = (int64_t)b * N * rs;
    const int64_t oclip = (int64_t)b * N * H * kPerfDim;
    (void)oclip;
    const int n0 = tile * kPasRows, rows = min(kPasRows, N - n0);
    pas_stage_projection(proj, M, P);
    if (t >= kPasThreads - 32) {                             // the last warp stages the v tile
        const int u = t - (kPasThreads - 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = u + 32 * i, r = f / (kPerfDim / 4), d = 4 * (f % (kPerfDim / 4));
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) {
                x = __ldg(reinterpret_cast<const float4*>(v + clip + (int64_t)(n0 + r) * rs + h * kPerfDim + d));
                if (vb) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(vb + h * kPerfDim + d));
                    x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
                }
            }
            *reinterpret_cast<float4*>(vt + r * kPerfDim + d) = x;
        }
    }
    __syncthreads();
    pas_features(k + clip, kb, h, rs, n0, rows, P, xt, ft, rowstat, wmax, M, false, ratio, eps);
    float c[kPerfDim + 1];
#pragma unroll
    for (int e = 0; e <= kPerfDim; ++e) c[e] = 0.0f;
#pragma unroll 2
    for (int r = 0; r < kPasRows; ++r) {
        const float kf = ft[r * kPpfCols + t];
        const float4* vr = reinterpret_cast<const float4*>(vt + r * kPerfDim);
#pragma unroll
        for (int e4 = 0; e4 < kPerfDim / 4; ++e4) {
            const float4 vv = vr[e4];
            c[4 * e4 + 0] = fmaf(kf, vv.x, c[4 * e4 + 0]);
            c[4 * e4 + 1] = fmaf(kf, vv.y, c[4 * e4 + 1]);
            c[4 * e4 + 2] = fmaf(kf, vv.z, c[4 * e4 + 2]);
            c[4 * e4 + 3] = fmaf(kf, vv.w, c[4 * e4 + 3]);
        }
        c[kPerfDim] += kf;
    }
    float* dst = partial + (((int64_t)b * H + h) * gridDim.z + tile) * kPctxFloats + t * kPasPStride;
#pragma unroll
    for (int e = 0; e <= kPerfDim; ++e) dst[e] = c[e];
}

__global__ void __launch_bounds__(256) performer_context_reduce_kernel(const float* __restrict__ partial,
                                                                       float* __restrict__ context, int tiles) {
    const int64_t bh = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kPctxFloats) return;
    const float* src = partial + bh * tiles * kPctxFloats + i;
    float a = 0.0f;
    for (int tl = 0; tl < tiles; ++tl) a += __ldg(src + (int64_t)tl * kPctxFloats);      // fixed order
    context[bh * kPctxFloats + i] = a;
}

__global__ void __launch_bounds__(kPasThreads) performer_output_kernel(
    const float* __restrict__ q, const float* __restrict__ qb, const float* __restrict__ proj,
    const float* __restrict__ context, float* __restrict__ out, int N, int H, int M, int64_t rs, float ratio, float eps) {
    extern __shared__ __align__(16) float pas_smem[];
    float* P = pas_smem;
    float* ctx = P + kPpfCols * kPasPStride;
    float* ft = ctx + kPpfCols * kPasPStride;
    float* xt = ft + kPasRows * kPpfCols;
    float* vt = xt + kPasRows * kPerfDim;
    float* rowstat = vt + kPasRows * kPerfDim;
    float* wmax = rowstat + 3 * kPasRows;
    const int t = threadIdx.x, h = blockIdx.x, b = blockIdx.y, tile = blockIdx.z;
    const int64_t clip = (int64_t)b * N * rs;
    const int64_t oclip = (int64_t)b * N * H * kPerfDim;
    (void)oclip;
    const int n0 = tile * kPasRows, rows = min(kPasRows, N - n0);
    pas_stage_projection(proj, M, P);
    {
        const float4* src = reinterpret_cast<const float4*>(context + ((int64_t)b * H + h) * kPctxFloats);
        float4* dst = reinterpret_cast<float4*>(ctx);
        for (int e = t; e < kPctxFloats / 4; e += kPasThreads) dst[e] = __ldg(src + e);
    }
    __syncthreads();
    pas_features(q + clip, qb, h, rs, n0, rows, P, xt, ft, rowstat, wmax, M, true, ratio, eps);
    for (int o = t; o < kPasRows * kPasPStride; o += kPasThreads) {
        const int r = o / kPasPStride, e = o % kPasPStride;
        const float* fr = ft + r * kPpfCols;
        const float* cc = ctx + e;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        int j = 0;
#pragma unroll 4
        for (; j + 4 <= M; j += 4) {
            const float4 f4 = *reinterpret_cast<const float4*>(fr + j);
            a0 = fmaf(f4.x, cc[j * kPasPStride], a0);
            a1 = fmaf(f4.y, cc[(j + 1) * kPasPStride], a1);
            a2 = fmaf(f4.z, cc[(j + 2) * kPasPStride], a2);
            a3 = fmaf(f4.w, cc[(j + 3) * kPasPStride], a3);
        }
        for (; j < M; ++j) a0 = fmaf(fr[j], cc[j * kPasPStride], a0);
        const float a = (a0 + a1) + (a2 + a3);
        if (e == kPerfDim) rowstat[2 * kPasRows + r] = a;
        else xt[r * kPerfDim + e] = a;
    }
    __syncthreads();
    for (int o = t; o < rows * kPerfDim; o += kPasThreads) {
        const int r = o / kPerfDim;
        out[oclip + ((int64_t)(n0 + r) * H + h) * kPerfDim + (o % kPerfDim)] = xt[o] * (1.0f / (rowstat[2 * kPasRows + r] + 1e-8f));
    }
}

// Input embedding sum of Unit2Control.forward (unit2control.py:80-95) in one pass:
//   out[b,n,c] = x[b,n,c] + (wf[c]*log(1 + f0/700) + bf[c]) + (wp[c]*(phase/pi) + bp[c]) + (wv[c]*vol + bv[c]) + spk[b?,c]
// x may be any strided (B,N,C) view (the pre-net output is a transposed view); out is contiguous.
// spk is (1,C) or (B,C): the (possibly mixed) speaker embedding row.  Replaces ~12 tiny kernels.
__global__ void __launch_bounds__(256) embed_sum_kernel(const float* __restrict__ x, int64_t xB, int64_t xN, int64_t xC,
                                                        const float* __restrict__ f0, int64_t fB, int64_t fN,
                                                        const float* __restrict__ phase, int64_t pB, int64_t pN,
                                                        const float* __restrict__ vol, int64_t vB, int64_t vN,
                                                        const float* __restrict__ wf, const float* __restrict__ bf,
                                                        const float* __restrict__ wp, const float* __restrict__ bp,
                                                        const float* __restrict__ wv, const float* __restrict__ bv,
                                                        const float* __restrict__ spk, int64_t sB, int B, int N, int C,
                                                        float* __restrict__ out) {
    const int64_t total = (int64_t)B * N * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t bn = i / C;
        const int n = (int)(bn % N), b = (int)(bn / N);
        const float lf0 = logf(__fadd_rn(1.0f, __fdiv_rn(__ldg(f0 + b * fB + n * fN), 700.0f)));
        const float ph = __fdiv_rn(__ldg(phase + b * pB + n * pN), 3.14159265358979323846f);
        const float vv = __ldg(vol + b * vB + n * vN);
        float v = __ldg(x + b * xB + n * xN + c * xC);
        v = __fadd_rn(v, fmaf(lf0, __ldg(wf + c), __ldg(bf + c)));
        v = __fadd_rn(v, fmaf(ph, __ldg(wp + c), __ldg(bp + c)));
        v = __fadd_rn(v, fmaf(vv, __ldg(wv + c), __ldg(bv + c)));
        v = __fadd_rn(v, __ldg(spk + b * sB + c));
        out[i] = v;
    }
}

// GLU + depthwise conv (k=31, zero 'same' padding) + SiLU, channels last.
//   u   : (B, T, 2C) contiguous (output of the first pointwise conv, bias included)
//   w   : (C, 31) depthwise taps (Conv1d weight (C,1,31)), bias (C)
//   out : (B, T, C):  silu(bias[c] + sum_j w[c][j] * g[b][t + j - 15][c]),  g = u[..., :C] * sigmoid(u[..., C:])
constexpr int kDwTaps = 31, kDwPad = 15;
constexpr int kDwTileT = 64, kDwTileC = 64, kDwPerThread = 16;      // 256 threads: 64 channels x 4 time groups

__global__ void __launch_bounds__(256) glu_dwconv_silu_kernel(const float* __restrict__ u, const float* __restrict__ ubias,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              float* __restrict__ out, int B, int T, int C) {
    __shared__ float g[(kDwTileT + kDwTaps - 1) * kDwTileC];
    const int c0 = blockIdx.x * kDwTileC, t0 = blockIdx.y * kDwTileT, b = blockIdx.z;
    const int cl = threadIdx.x & (kDwTileC - 1), tg = threadIdx.x / kDwTileC;
    const int c = c0 + cl;
    const float* ub = u + (int64_t)b * T * 2 * C;
    for (int e = threadIdx.x; e < (kDwTileT + kDwTaps - 1) * kDwTileC; e += 256) {
        const int tt = e / kDwTileC, cc = e % kDwTileC;
        const int t = t0 + tt - kDwPad;
        float val = 0.0f;
        if (t >= 0 && t < T && c0 + cc < C) {
            float a = __ldg(ub + (int64_t)t * 2 * C + c0 + cc);
            float gate = __ldg(ub + (int64_t)t * 2 * C + C + c0 + cc);
            if (ubias) { a += __ldg(ubias + c0 + cc); gate += __ldg(ubias + C + c0 + cc); }   // bias of the producing Linear
            val = a / (1.0f + expf(-gate));
        }
        g[e] = val;
    }
    __syncthreads();
    if (c >= C) return;
    float wr[kDwTaps];
#pragma unroll
    for (int j = 0; j < kDwTaps; ++j) wr[j] = __ldg(w + (int64_t)c * kDwTaps + j);
    const float bs = __ldg(bias + c);
    float acc[kDwPerThread];
#pragma unroll
    for (int o = 0; o < kDwPerThread; ++o) acc[o] = bs;
    const float* gp = g + (tg * kDwPerThread) * kDwTileC + cl;
#pragma unroll
    for (int i = 0; i < kDwPerThread + kDwTaps - 1; ++i) {
        const float val = gp[i * kDwTileC];
#pragma unroll
        for (int o = 0; o < kDwPerThread; ++o) {
            const int j = i - o;
            if (j >= 0 && j < kDwTaps) acc[o] = fmaf(wr[j], val, acc[o]);
        }
    }
    float* ob = out + (int64_t)b * T * C + c;
#pragma unroll
    for (int o = 0; o < kDwPerThread; ++o) {
        const int t = t0 + tg * kDwPerThread + o;
        if (t < T) ob[(int64_t)t * C] = acc[o] / (1.0f + expf(-acc[o]));
    }
}

// Depthwise Conv1d(k=31, 'same') + SiLU on a gated channels-last tensor (the GLU already applied by the epilogue of the
// producing GEMM, csrc/gemm_attn.cuh EPI_GLU).  A thread owns one channel and kDw2Run consecutive frames: the 31 taps and
// a sliding window of inputs stay in registers, every input is loaded once per thread (lanes = consecutive channels:
// 128-byte coalesced rows; the 30-frame halo of neighbouring runs comes out of L1 / L2).
#ifndef DW2_RUN
#define DW2_RUN 24      // measured on B200 at 64 x 862 x 512: 16 -> 77 us, 20 -> 68, 24 -> 65, 32 -> 71, 40 -> 70, 48 -> 82
#endif
constexpr int kDw2Run = DW2_RUN;

// C_T: compile-time channel count (0 = runtime C).  With the shipped width (512) every load / store address is the thread's
// base pointer plus an immediate offset, and runs that lie inside the clip (all but the first and last one) carry no range
// tests: the checked form spent 14 address / predicate instructions per load (2420 warp instructions per run for 992 FMAs,
// ncu), this one ~1.4 k.
template <int C_T, bool CHECKED>
__device__ __forceinline__ void dwconv_silu_run(const float* __restrict__ gp, const float (&wr)[kDwTaps], float bs,
                                                float* __restrict__ op, int t0, int T, int C) {
    const int64_t Cc = C_T ? C_T : C;
    float acc[kDw2Run];
#pragma unroll
    for (int o = 0; o < kDw2Run; ++o) acc[o] = bs;
#pragma unroll
    for (int i = 0; i < kDw2Run + kDwTaps - 1; ++i) {
        float val;
        if (CHECKED) {
            const int t = t0 + i - kDwPad;
            val = (t >= 0 && t < T) ? __ldg(gp + i * Cc) : 0.0f;
        } else {
            val = __ldg(gp + i * Cc);
        }
#pragma unroll
        for (int o = 0; o < kDw2Run; ++o) {
            const int j = i - o;
            if (j >= 0 && j < kDwTaps) acc[o] = fmaf(wr[j], val, acc[o]);
        }
    }
#pragma unroll
    for (int o = 0; o < kDw2Run; ++o) {
        if (!CHECKED || t0 + o < T) op[o * Cc] = __fdividef(acc[o], 1.0f + __expf(-acc[o]));
    }
}

template <int C_T>
__global__ void __launch_bounds__(256) dwconv_silu_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ out, int T, int C) {
    const int Cc = C_T ? C_T : C;
    const int c = blockIdx.x * 256 + threadIdx.x, t0 = blockIdx.y * kDw2Run, b = blockIdx.z;
    if (c >= Cc) return;
    float wr[kDwTaps];
#pragma unroll
    for (int j = 0; j < kDwTaps; ++j) wr[j] = __ldg(w + (int64_t)c * kDwTaps + j);
    const float bs = __ldg(bias + c);
    // (the pointer of the first window frame may lie in front of the clip: it is only dereferenced where the frame exists)
    const float* gp = g + ((int64_t)b * T + (t0 - kDwPad)) * Cc + c;
    float* op = out + ((int64_t)b * T + t0) * Cc + c;
    if (t0 >= kDwPad && t0 + kDw2Run + kDwPad <= T) dwconv_silu_run<C_T, false>(gp, wr, bs, op, t0, T, C);   // block-uniform
    else dwconv_silu_run<C_T, true>(gp, wr, bs, op, t0, T, C);
}

// ---- unit pre-net on the tensor cores (unit2control.py:38-45) -----------------------------------------------------------
// Conv1d(k=3, 'same') over channels-last frames is a GEMM whose row m is the window of three consecutive frames: with one
// zero frame in front of and behind every clip the window of frame n is the 3*C contiguous floats that start at padded
// frame n, so the A operand is the padded buffer itself read with row stride C and K = 3*C (overlapping rows; the two
// rows per clip that straddle a clip boundary produce garbage that lands exactly on the pad frames of the next buffer).

// (B,N,C) strided -> (B,N+2,C) contiguous with zero frames 0 and N+1.  One thread per float4.
__global__ void __launch_bounds__(256) pad_frames_kernel(const float* __restrict__ x, int64_t xB, int64_t xN, int B, int N,
                                                         int C, float4* __restrict__ out) {
    const int c4 = C >> 2;
    const int64_t total = (int64_t)B * (N + 2) * c4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4);
        const int64_t r = i / c4;
        const int n = (int)(r % (N + 2)) - 1, b = (int)(r / (N + 2));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n >= 0 && n < N) v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)b * xB + (int64_t)n * xN) + c);
        out[i] = v;
    }
}

// GroupNorm statistics of a padded (B,N+2,C) buffer over the N real frames: sums[b][g] = (sum, sum of squares) in fp64.
// grid (chunks, B); a thread owns 4 consecutive channels (one group: C/G is a multiple of 4) and strides over the frames.
constexpr int kGnThreads = 256;
__global__ void __launch_bounds__(kGnThreads) groupnorm_stats_kernel(const float* __restrict__ hp, int N, int C, int G,
                                                                      double* __restrict__ sums) {
    __shared__ double sh[2 * 32];                         // up to 32 groups
    const int b = blockIdx.y, c4 = C >> 2;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const int col = threadIdx.x % c4, row0 = threadIdx.x / c4, rows_per_pass = blockDim.x / c4;
    const int per = (N + gridDim.x - 1) / gridDim.x, n0 = blockIdx.x * per, n1 = min(N, n0 + per);
    float s = 0.f, q = 0.f;
    const float4* base = reinterpret_cast<const float4*>(hp + ((int64_t)b * (N + 2) + 1) * C) + col;
    for (int n = n0 + row0; n < n1; n += rows_per_pass) {
        const float4 v = __ldg(base + (int64_t)n * c4);
        s += (v.x + v.y) + (v.z + v.w);
        q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, q))));
    }
    const int g = (4 * col) / (C / G);
    atomicAdd(&sh[2 * g], (double)s);
    atomicAdd(&sh[2 * g + 1], (double)q);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(&sums[(int64_t)b * 2 * G + i], sh[i]);
}

// In place: hp[b,1+n,c] = leaky_relu(gamma[c] * (h - mean[b,g]) * rstd[b,g] + beta[c]); pad frames 0 and N+1 := 0.
__global__ void __launch_bounds__(256) groupnorm_leaky_kernel(float4* __restrict__ hp, const double* __restrict__ sums,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, float slope, int B, int N, int C, int G) {
    const int c4 = C >> 2;
    const int64_t total = (int64_t)B * (N + 2) * c4;
    const double inv_cnt = 1.0 / ((double)N * (C / G));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        const int64_t r = i / c4;
        const int n = (int)(r % (N + 2)) - 1, b = (int)(r / (N + 2));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n >= 0 && n < N) {
            const int g = c / (C / G);
            const double m = sums[((int64_t)b * G + g) * 2] * inv_cnt;
            const double var = fmax(sums[((int64_t)b * G + g) * 2 + 1] * inv_cnt - m * m, 0.0);      // biased, as GroupNorm
            const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)eps));
            const float4 h = hp[i];
            const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c)), be = __ldg(reinterpret_cast<const float4*>(beta + c));
            v.x = fmaf((h.x - mean) * rstd, ga.x, be.x); v.y = fmaf((h.y - mean) * rstd, ga.y, be.y);
            v.z = fmaf((h.z - mean) * rstd, ga.z, be.z); v.w = fmaf((h.w - mean) * rstd, ga.w, be.w);
            v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
            v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
        }
        hp[i] = v;
    }
}

// Input embedding sum (embed_sum_kernel above) + the first LayerNorm of PCmer (pcmer.py:25) for C = 256: one warp per frame,
// lane owns channels 4l..4l+3 and 128+4l..128+4l+3 (two 128-bit accesses per tensor); the per-frame scalars
// log(1 + f0/700), phase/pi, volume are computed once per frame instead of once per channel.
__global__ void __launch_bounds__(256) embed_sum_ln256_kernel(const float* __restrict__ x, int64_t xB, int64_t xN,
                                                              const float* __restrict__ f0, int64_t fB, int64_t fN,
                                                              const float* __restrict__ phase, int64_t pB, int64_t pN,
                                                              const float* __restrict__ vol, int64_t vB, int64_t vN,
                                                              const float* __restrict__ wf, const float* __restrict__ bf,
                                                              const float* __restrict__ wp, const float* __restrict__ bp,
                                                              const float* __restrict__ wv, const float* __restrict__ bv,
                                                              const float* __restrict__ spk, int64_t sB,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, int B, int N, float* __restrict__ out,
                                                              float* __restrict__ out_ln) {
    const int lane = threadIdx.x & 31;
    const int64_t rows = (int64_t)B * N;
    const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), wstep = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = w0; r < rows; r += wstep) {
        const int n = (int)(r % N), b = (int)(r / N);
        const float lf0 = logf(__fadd_rn(1.0f, __fdiv_rn(__ldg(f0 + b * fB + n * fN), 700.0f)));
        const float ph = __fdiv_rn(__ldg(phase + b * pB + n * pN), 3.14159265358979323846f);
        const float vv = __ldg(vol + b * vB + n * vN);
        float v[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 128 * h + 4 * lane;
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + b * xB + n * xN + c));
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(wf + c)), a1 = __ldg(reinterpret_cast<const float4*>(bf + c));
            const float4 p0 = __ldg(reinterpret_cast<const float4*>(wp + c)), p1 = __ldg(reinterpret_cast<const float4*>(bp + c));
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(wv + c)), v1 = __ldg(reinterpret_cast<const float4*>(bv + c));
            const float4 sp = __ldg(reinterpret_cast<const float4*>(spk + b * sB + c));
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, wfs[4] = {a0.x, a0.y, a0.z, a0.w}, bfs[4] = {a1.x, a1.y, a1.z, a1.w};
            const float wps[4] = {p0.x, p0.y, p0.z, p0.w}, bps[4] = {p1.x, p1.y, p1.z, p1.w};
            const float wvs[4] = {v0.x, v0.y, v0.z, v0.w}, bvs[4] = {v1.x, v1.y, v1.z, v1.w}, sps[4] = {sp.x, sp.y, sp.z, sp.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {                 // the operation order of embed_sum_kernel
                float t = __fadd_rn(xs[e], fmaf(lf0, wfs[e], bfs[e]));
                t = __fadd_rn(t, fmaf(ph, wps[e], bps[e]));
                t = __fadd_rn(t, fmaf(vv, wvs[e], bvs[e]));
                v[4 * h + e] = __fadd_rn(t, sps[e]);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[e];
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / 256.0f);
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) q = fmaf(v[e] - mean, v[e] - mean, q);
#pragma unroll
        for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / 256.0f) + eps);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 128 * h + 4 * lane;
            const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c)), be = __ldg(reinterpret_cast<const float4*>(beta + c));
            *reinterpret_cast<float4*>(out + r * 256 + c) = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
            *reinterpret_cast<float4*>(out_ln + r * 256 + c) =
                make_float4(fmaf((v[4 * h] - mean) * rstd, ga.x, be.x), fmaf((v[4 * h + 1] - mean) * rstd, ga.y, be.y),
                            fmaf((v[4 * h + 2] - mean) * rstd, ga.z, be.z), fmaf((v[4 * h + 3] - mean) * rstd, ga.w, be.w));
        }
    }
}

}  // namespace ddsp
