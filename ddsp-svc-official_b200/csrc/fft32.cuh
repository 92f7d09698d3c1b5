// fft32.cuh -- register-resident 32-point complex FFT and the warp-level 1024-point FFT
// (32 x 32 four-step: each lane owns 32 points; one transpose through shared memory).
//
// The butterflies use Blackwell's packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2, sm_100+):
// register-index r and r+16 of the 32-point transform share one 64-bit register pair, so the
// stages with span 1, 2, 4, 8 process two butterflies per instruction (the twiddle is the same
// for both halves and rides as a broadcast immediate); only the last stage (span 16) works
// inside the pairs with scalar FFMAs.  The kernel is issue-bound, so halving the FP instruction
// count of the FFT is worth more than anything the (identical) FLOP rate could give.
#pragma once
#include "common.cuh"

namespace ddsp {

__host__ __device__ constexpr int brev5(int i) {
    return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// cos/sin(2*pi*j/32), j = 0..15 (compile-time constants once the loops are unrolled)
__device__ __forceinline__ constexpr float cos32(int j) {
    return j == 0 ? 1.0f : j == 1 ? 0.98078528040323044913f : j == 2 ? 0.92387953251128675613f
         : j == 3 ? 0.83146961230254523708f : j == 4 ? 0.70710678118654752440f
         : j == 5 ? 0.55557023301960222474f : j == 6 ? 0.38268343236508977173f
         : j == 7 ? 0.19509032201612826785f : j == 8 ? 0.0f
         : j == 9 ? -0.19509032201612826785f : j == 10 ? -0.38268343236508977173f
         : j == 11 ? -0.55557023301960222474f : j == 12 ? -0.70710678118654752440f
         : j == 13 ? -0.83146961230254523708f : j == 14 ? -0.92387953251128675613f
         : -0.98078528040323044913f;
}
__device__ __forceinline__ constexpr float sin32(int j) {
    return j == 0 ? 0.0f : j == 1 ? 0.19509032201612826785f : j == 2 ? 0.38268343236508977173f
         : j == 3 ? 0.55557023301960222474f : j == 4 ? 0.70710678118654752440f
         : j == 5 ? 0.83146961230254523708f : j == 6 ? 0.92387953251128675613f
         : j == 7 ? 0.98078528040323044913f : j == 8 ? 1.0f
         : j == 9 ? 0.98078528040323044913f : j == 10 ? 0.92387953251128675613f
         : j == 11 ? 0.83146961230254523708f : j == 12 ? 0.70710678118654752440f
         : j == 13 ? 0.55557023301960222474f : j == 14 ? 0.38268343236508977173f
         : 0.19509032201612826785f;
}

// (packed fp32x2 helpers add2 / mul2 / fma2 / neg2 / bc2 / sub2 live in common.cuh)

// The 32 points of one lane: element with register-index r lives in R[r & 15] (.x for r < 16,
// .y for r >= 16), real parts in R, imaginary parts in I.
struct Pts32 {
    float2 R[16], I[16];
};
#define DDSP_RE(P, r) ((r) < 16 ? (P).R[(r) & 15].x : (P).R[(r) & 15].y)
#define DDSP_IM(P, r) ((r) < 16 ? (P).I[(r) & 15].x : (P).I[(r) & 15].y)

// One decimation-in-time stage with butterfly span HALF in {1,2,4,8} on the 16 packed positions:
// x0' = a + w b, x1' = a - w b, two butterflies per instruction.
template <int HALF>
__device__ __forceinline__ void dit_stage2(Pts32& P) {
    constexpr float Rq = 0.70710678118654752440f;
    constexpr int tstep = 16 / HALF;
#pragma unroll
    for (int g = 0; g < 16; g += 2 * HALF) {
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
            const int i0 = g + j, i1 = i0 + HALF;
            const float2 ar = P.R[i0], ai = P.I[i0], br = P.R[i1], bi = P.I[i1];
            const int tw = j * tstep;                 // w = W32^tw = cos - j sin
            if (tw == 0) {
                P.R[i0] = add2(ar, br); P.I[i0] = add2(ai, bi);
                P.R[i1] = sub2(ar, br); P.I[i1] = sub2(ai, bi);
            } else if (tw == 8) {                      // w = -j : w b = (bi, -br)
                P.R[i0] = add2(ar, bi); P.I[i0] = sub2(ai, br);
                P.R[i1] = sub2(ar, bi); P.I[i1] = add2(ai, br);
            } else if (tw == 4) {                      // w = (1-j)/sqrt2 : w b = q(br+bi) + j q(bi-br)
                const float2 t1 = add2(br, bi), t2 = sub2(bi, br);
                P.R[i0] = fma2(t1, bc2(Rq), ar); P.I[i0] = fma2(t2, bc2(Rq), ai);
                P.R[i1] = fma2(t1, bc2(-Rq), ar); P.I[i1] = fma2(t2, bc2(-Rq), ai);
            } else if (tw == 12) {                     // w = (-1-j)/sqrt2 : w b = q(bi-br) - j q(br+bi)
                const float2 t1 = sub2(bi, br), t2 = add2(br, bi);
                P.R[i0] = fma2(t1, bc2(Rq), ar); P.I[i0] = fma2(t2, bc2(-Rq), ai);
                P.R[i1] = fma2(t1, bc2(-Rq), ar); P.I[i1] = fma2(t2, bc2(Rq), ai);
            } else {                                    // w b = (c br + s bi) + j(c bi - s br)
                const float c = cos32(tw), s = sin32(tw);
                const float2 xr = fma2(br, bc2(c), fma2(bi, bc2(s), ar));
                const float2 xi = fma2(bi, bc2(c), fma2(br, bc2(-s), ai));
                P.R[i0] = xr; P.I[i0] = xi;
                P.R[i1] = fma2(ar, bc2(2.0f), neg2(xr));      // a - w b = 2a - (a + w b)
                P.I[i1] = fma2(ai, bc2(2.0f), neg2(xi));
            }
        }
    }
}

// Last stage (span 16): butterflies between the two halves of each pair, twiddle W32^r.
__device__ __forceinline__ void dit_stage_last(Pts32& P) {
    constexpr float Rq = 0.70710678118654752440f;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const float ar = P.R[r].x, ai = P.I[r].x, br = P.R[r].y, bi = P.I[r].y;
        if (r == 0) {
            P.R[r] = make_float2(ar + br, ar - br); P.I[r] = make_float2(ai + bi, ai - bi);
        } else if (r == 8) {
            P.R[r] = make_float2(ar + bi, ar - bi); P.I[r] = make_float2(ai - br, ai + br);
        } else if (r == 4) {
            const float t1 = br + bi, t2 = bi - br;
            P.R[r] = make_float2(fmaf(Rq, t1, ar), fmaf(-Rq, t1, ar));
            P.I[r] = make_float2(fmaf(Rq, t2, ai), fmaf(-Rq, t2, ai));
        } else if (r == 12) {
            const float t1 = bi - br, t2 = br + bi;
            P.R[r] = make_float2(fmaf(Rq, t1, ar), fmaf(-Rq, t1, ar));
            P.I[r] = make_float2(fmaf(-Rq, t2, ai), fmaf(Rq, t2, ai));
        } else {
            const float c = cos32(r), s = sin32(r);
            const float xr = fmaf(c, br, fmaf(s, bi, ar));
            const float xi = fmaf(c, bi, fmaf(-s, br, ai));
            P.R[r] = make_float2(xr, fmaf(2.0f, ar, -xr));
            P.I[r] = make_float2(xi, fmaf(2.0f, ai, -xi));
        }
    }
}

// 32-point forward FFT, decimation in time: element n must sit at register-index brev5(n) on
// entry; on exit register-index k holds X[k] (natural order).
__device__ __forceinline__ void fft32_dit(Pts32& P) {
    dit_stage2<1>(P);
    dit_stage2<2>(P);
    dit_stage2<4>(P);
    dit_stage2<8>(P);
    dit_stage_last(P);
}

// Same, but every element n is first multiplied by this lane's twiddle W1024^(n*lane).  The
// packed position i holds elements (e, e+1), e = brev5(i); `tw4[(e/2)*32 + lane]` delivers
// (cos e, cos e+1, -sin e, -sin e+1) in one 128-bit load.  The multiply is folded into the span-1
// butterflies (10 packed ops per two butterflies).
__device__ __forceinline__ void fft32_dit_twiddled(Pts32& P, const float4* __restrict__ tw4, int lane) {
#pragma unroll
    for (int i0 = 0; i0 < 16; i0 += 2) {
        const int e = brev5(i0);                        // even, bit 4 clear; position i0+1 holds (e+16, e+17)
        const float4 wa = tw4[(e >> 1) * 32 + lane];
        const float4 wb = tw4[((e >> 1) + 8) * 32 + lane];
        const float2 wax = make_float2(wa.x, wa.y), way = make_float2(wa.z, wa.w);
        const float2 wbx = make_float2(wb.x, wb.y), wby = make_float2(wb.z, wb.w);
        const float2 Ar = P.R[i0], Ai = P.I[i0], Br = P.R[i0 + 1], Bi = P.I[i0 + 1];
        const float2 ar = fma2(neg2(Ai), way, mul2(Ar, wax));       // a = wa * A
        const float2 ai = fma2(Ar, way, mul2(Ai, wax));
        const float2 xr = fma2(Br, wbx, fma2(neg2(Bi), wby, ar));   // a + wb * B
        const float2 xi = fma2(Br, wby, fma2(Bi, wbx, ai));
        P.R[i0] = xr;
        P.I[i0] = xi;
        P.R[i0 + 1] = fma2(ar, bc2(2.0f), neg2(xr));                // a - wb * B = 2a - (a + wb * B)
        P.I[i0 + 1] = fma2(ai, bc2(2.0f), neg2(xi));
    }
    dit_stage2<2>(P);
    dit_stage2<4>(P);
    dit_stage2<8>(P);
    dit_stage_last(P);
}

constexpr int kPlaneStride = 34;                 // even (64-bit reads) and conflict-free for both access patterns
constexpr int kPlaneFloats = 32 * kPlaneStride;  // 1088 floats = 4352 B per warp

// Forward complex FFT of 1024 points spread over one warp.
//   in : element x[32*n1 + lane] at register-index brev5(n1)
//   out: X[lane + 32*k2] at register-index k2 (natural order)
// `plane` is this warp's private 32x34-float transpose buffer, `tw4` the CTA-wide twiddle table.
// The inverse transform is obtained by calling it with real and imaginary parts swapped.
// `mid()` runs between the two passes, after the last access to `plane`: the buffer is free from there until the next
// transform (combsubfast_kernel starts the bulk copy of the next filter rows into it at that point).
template <class Mid>
__device__ __forceinline__ void warp_fft1024(Pts32& P, float* __restrict__ plane, const float4* __restrict__ tw4,
                                             int lane, Mid&& mid) {
    fft32_dit(P);                                    // register-index k1 = A[k1] for column n2 = lane
    float* wr = plane + lane;
    const float2* rd = reinterpret_cast<const float2*>(plane + lane * kPlaneStride);
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) wr[k1 * kPlaneStride] = DDSP_RE(P, k1);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) P.R[i] = rd[brev5(i) >> 1];       // elements (e, e+1), e = brev5(i)
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) wr[k1 * kPlaneStride] = DDSP_IM(P, k1);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) P.I[i] = rd[brev5(i) >> 1];
    __syncwarp();
    mid();
    fft32_dit_twiddled(P, tw4, lane);                // row k1 = lane: X[lane + 32*k2]
}
__device__ __forceinline__ void warp_fft1024(Pts32& P, float* __restrict__ plane, const float4* __restrict__ tw4,
                                             int lane) {
    warp_fft1024(P, plane, tw4, lane, [] {});
}

// Device-wide constant tables (filled once per device by fft_tables_kernel):
//   [0, 512) float4 twiddles tw4[(e/2)*32 + lane] = (cos(e l), cos((e+1) l), -sin(e l), -sin((e+1) l)),
//            angles 2 pi e l / 1024, e even
//   then the window w[i] = sin(pi i / 1024) = sqrt(hann_periodic(1024))[i] twice, as float2 pairs in the
//   two orders the kernels consume it (one 64-bit shared load feeds one packed fp32x2 multiply):
//     analysis order  A[j*32 + lane] = (w[64j + lane], w[64j + 32 + lane])        j < 16
//     synthesis order B[q*32 + lane] = (w[32q + lane], w[512 + 32q + lane])       q < 16
constexpr int kTableBytes = 512 * 16 + 2048 * 4;

__host__ __device__ __forceinline__ int win_analysis_index(int i) {      // float index of w[i] in table A
    return 2 * ((i >> 6) * 32 + (i & 31)) + ((i >> 5) & 1);
}
__host__ __device__ __forceinline__ int win_synthesis_index(int i) {     // float index of w[i] in table B (after A)
    return 1024 + 2 * (((i & 511) >> 5) * 32 + (i & 31)) + (i >> 9);
}

__global__ void fft_tables_kernel(float4* __restrict__ tw4, float* __restrict__ win) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 1024) return;
    double s, c;
    if (t < 512) {
        const int e = 2 * (t >> 5), l = t & 31;
        double s1, c1;
        sincospi((double)(e * l) / 512.0, &s, &c);
        sincospi((double)((e + 1) * l) / 512.0, &s1, &c1);
        tw4[t] = make_float4((float)c, (float)c1, (float)(-s), (float)(-s1));
    }
    sincospi((double)t / 1024.0, &s, &c);
    win[win_analysis_index(t)] = (float)s;
    win[win_synthesis_index(t)] = (float)s;
}

}  // namespace ddsp
