// fft32.cuh -- register-resident 32-point complex FFT and the warp-level 1024-point FFT
// (32 x 32 four-step: each lane owns 32 points; one transpose through shared memory).
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

// One decimation-in-time stage (butterfly span `half`) on 32 register-resident points.
// x0' = a + w b, x1' = a - w b = 2a - x0' : a general butterfly is 6 FFMA.
template <int HALF>
__device__ __forceinline__ void dit_stage(float (&re)[32], float (&im)[32]) {
    constexpr float R = 0.70710678118654752440f;
    constexpr int tstep = 16 / HALF;
#pragma unroll
    for (int g = 0; g < 32; g += 2 * HALF) {
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
            const int i0 = g + j, i1 = i0 + HALF;
            const float ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
            const int tw = j * tstep;                 // w = W32^tw = cos - j sin
            if (tw == 0) {
                re[i0] = ar + br; im[i0] = ai + bi; re[i1] = ar - br; im[i1] = ai - bi;
            } else if (tw == 8) {                      // w = -j : w b = (bi, -br)
                re[i0] = ar + bi; im[i0] = ai - br; re[i1] = ar - bi; im[i1] = ai + br;
            } else if (tw == 4) {                      // w = (1-j)/sqrt2 : w b = R(br+bi) + jR(bi-br)
                const float t1 = br + bi, t2 = bi - br;
                re[i0] = fmaf(R, t1, ar); im[i0] = fmaf(R, t2, ai);
                re[i1] = fmaf(-R, t1, ar); im[i1] = fmaf(-R, t2, ai);
            } else if (tw == 12) {                     // w = (-1-j)/sqrt2 : w b = R(bi-br) - jR(br+bi)
                const float t1 = bi - br, t2 = br + bi;
                re[i0] = fmaf(R, t1, ar); im[i0] = fmaf(-R, t2, ai);
                re[i1] = fmaf(-R, t1, ar); im[i1] = fmaf(R, t2, ai);
            } else {                                    // w b = (c br + s bi) + j(c bi - s br)
                const float c = cos32(tw), s = sin32(tw);
                const float xr = fmaf(c, br, fmaf(s, bi, ar));
                const float xi = fmaf(c, bi, fmaf(-s, br, ai));
                re[i0] = xr; im[i0] = xi;
                re[i1] = fmaf(2.0f, ar, -xr); im[i1] = fmaf(2.0f, ai, -xi);
            }
        }
    }
}

// 32-point forward FFT, decimation in time: element n must sit in register brev5(n) on entry;
// on exit register k holds X[k] (natural order).
__device__ __forceinline__ void fft32_dit(float (&re)[32], float (&im)[32]) {
    dit_stage<1>(re, im);
    dit_stage<2>(re, im);
    dit_stage<4>(re, im);
    dit_stage<8>(re, im);
    dit_stage<16>(re, im);
}

// Same, but element n (register brev5(n)) is first multiplied by the per-lane twiddle tw[n*32+lane];
// the multiply is folded into the span-1 butterflies (10 ops per pair instead of 8 + 4).
__device__ __forceinline__ void fft32_dit_twiddled(float (&re)[32], float (&im)[32],
                                                   const float2* __restrict__ tw, int lane) {
#pragma unroll
    for (int g = 0; g < 16; ++g) {
        const int na = brev5(2 * g);                   // element index of register 2g (< 16)
        const int i0 = 2 * g, i1 = 2 * g + 1;          // register 2g+1 holds element na + 16
        float ar = re[i0], ai = im[i0];
        if (na != 0) {
            const float2 wa = tw[na * 32 + lane];
            const float tr = fmaf(-ai, wa.y, ar * wa.x);
            const float ti = fmaf(ar, wa.y, ai * wa.x);
            ar = tr; ai = ti;
        }
        const float2 wb = tw[(na + 16) * 32 + lane];
        const float br = re[i1], bi = im[i1];
        const float xr = fmaf(br, wb.x, fmaf(-bi, wb.y, ar));
        const float xi = fmaf(br, wb.y, fmaf(bi, wb.x, ai));
        re[i0] = xr; im[i0] = xi;
        re[i1] = fmaf(2.0f, ar, -xr); im[i1] = fmaf(2.0f, ai, -xi);
    }
    dit_stage<2>(re, im);
    dit_stage<4>(re, im);
    dit_stage<8>(re, im);
    dit_stage<16>(re, im);
}

constexpr int kPlaneStride = 33;                 // padded row stride of the transpose plane
constexpr int kPlaneFloats = 32 * kPlaneStride;  // 1056 floats = 4224 B per warp

// Forward complex FFT of 1024 points spread over one warp.
//   in : re/im[brev5(n1)] = x[32*n1 + lane]          (bit-reversed register order)
//   out: re/im[k2]        = X[lane + 32*k2]          (natural register order)
// `plane` is this warp's private 32x33-float transpose buffer, `tw` the CTA-wide table
// tw[a*32 + b] = (cos, -sin)(2*pi*a*b/1024).
// The inverse transform is obtained by calling it with the two arrays swapped.
__device__ __forceinline__ void warp_fft1024(float (&re)[32], float (&im)[32], float* __restrict__ plane,
                                             const float2* __restrict__ tw, int lane) {
    fft32_dit(re, im);                               // register k1 = A[k1] for column n2 = lane
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) plane[k1 * kPlaneStride + lane] = re[k1];
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) re[brev5(n2)] = plane[lane * kPlaneStride + n2];
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) plane[k1 * kPlaneStride + lane] = im[k1];
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) im[brev5(n2)] = plane[lane * kPlaneStride + n2];
    __syncwarp();
    fft32_dit_twiddled(re, im, tw, lane);            // row k1 = lane: X[lane + 32*k2]
}

// Device-wide constant tables (filled once per device by fft_tables_kernel):
//   [0, 1024) float2 twiddles tw[a*32+b] = (cos, -sin)(2 pi a b / 1024)
//   then 1024 floats sin(pi i / 1024) = sqrt(hann_periodic(1024))[i]
constexpr int kTableBytes = 1024 * 8 + 1024 * 4;

__global__ void fft_tables_kernel(float2* __restrict__ tw, float* __restrict__ win) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 1024) return;
    const int a = e >> 5, b = e & 31;
    double s, c;
    sincospi((double)(a * b) / 512.0, &s, &c);
    tw[e] = make_float2((float)c, (float)(-s));
    sincospi((double)e / 1024.0, &s, &c);
    win[e] = (float)s;
}

}  // namespace ddsp
