// control.cuh -- fused elementwise stages of the control network (SURVEY.md section 8f rank 1:
// ddsp/unit2control.py + ddsp/pcmer.py).  The GEMMs of the network stay library calls (cuBLAS, fp32);
// these two kernels replace the chains of broadcast/elementwise/reduction passes between them, each of
// which streamed a (B,H,N,266) or (B,N,1024) tensor through HBM:
//   * performer_features_kernel: the FAVOR+ softmax-kernel feature map (pcmer.py:124-160) from the raw
//     projections: one read of `dash`, one write of the features (was ~10 passes);
//   * glu_dwconv_silu_kernel: GLU -> depthwise Conv1d(k=31, 'same') -> SiLU of the conformer
//     convolution module (pcmer.py:41-63) in channels-last layout, no transposes.
// Both are HBM-bound streaming kernels in fp32 with full-precision expf (they feed exp() of the
// synthesizer, so no approximate transcendentals here).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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

// GLU + depthwise conv (k=31, zero 'same' padding) + SiLU, channels last.
//   u   : (B, T, 2C) contiguous (output of the first pointwise conv, bias included)
//   w   : (C, 31) depthwise taps (Conv1d weight (C,1,31)), bias (C)
//   out : (B, T, C):  silu(bias[c] + sum_j w[c][j] * g[b][t + j - 15][c]),  g = u[..., :C] * sigmoid(u[..., C:])
constexpr int kDwTaps = 31, kDwPad = 15;
constexpr int kDwTileT = 64, kDwTileC = 64, kDwPerThread = 16;      // 256 threads: 64 channels x 4 time groups

__global__ void __launch_bounds__(256) glu_dwconv_silu_kernel(const float* __restrict__ u, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              int B, int T, int C) {
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
            const float a = __ldg(ub + (int64_t)t * 2 * C + c0 + cc);
            const float gate = __ldg(ub + (int64_t)t * 2 * C + C + c0 + cc);
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

}  // namespace ddsp
