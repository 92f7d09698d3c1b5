// fft32.cuh -- register-resident 32-point complex FFT and the warp-level 1024-point FFT
// (32 x 32 four-step: each lane owns 32 points; one transpose through shared memory).
#pragma once
#include "common.cuh"

namespace ddsp {

__host__ __device__ constexpr int brev5(int i) {
    return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// cos/sin(2*pi*j/32), j = 0..15 (compile-time after unrolling)
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

// sin(2*pi*j/32), j = 0..15
__device__ __forceinline__ constexpr float sinw32(int j) {
    return j == 0 ? 0.0f : j == 1 ? 0.19509032201612826785f : j == 2 ? 0.38268343236508977173f
         : j == 3 ? 0.55557023301960222474f : j == 4 ? 0.70710678118654752440f
         : j == 5 ? 0.83146961230254523708f : j == 6 ? 0.92387953251128675613f
         : j == 7 ? 0.98078528040323044913f : j == 8 ? 1.0f
         : j == 9 ? 0.98078528040323044913f : j == 10 ? 0.92387953251128675613f
         : j == 11 ? 0.83146961230254523708f : j == 12 ? 0.70710678118654752440f
         : j == 13 ? 0.55557023301960222474f : j == 14 ? 0.38268343236508977173f
         : 0.19509032201612826785f;
}

// In-place radix-2 decimation-in-frequency FFT of 32 complex points held in registers.
// Forward transform (e^{-j...}); natural-order input, BIT-REVERSED output: after the call
// element i holds X[brev5(i)].  All indices are compile-time constants once unrolled, so the
// arrays never leave the register file.
__device__ __forceinline__ void fft32_dif(float (&re)[32], float (&im)[32]) {
    constexpr float R = 0.70710678118654752440f;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const int tstep = 16 / half;
#pragma unroll
        for (int g = 0; g < 32; g += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int i0 = g + j, i1 = i0 + half;
                const float ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
                re[i0] = ar + br;
                im[i0] = ai + bi;
                const float dr = ar - br, di = ai - bi;
                const int tw = j * tstep;   // twiddle W32^tw
                if (tw == 0) {
                    re[i1] = dr; im[i1] = di;
                } else if (tw == 8) {        // * (-j)
                    re[i1] = di; im[i1] = -dr;
                } else if (tw == 4) {        // * (1-j)/sqrt2
                    re[i1] = (dr + di) * R; im[i1] = (di - dr) * R;
                } else if (tw == 12) {       // * (-1-j)/sqrt2
                    re[i1] = (di - dr) * R; im[i1] = -(dr + di) * R;
                } else {                      // * (c - j s)
                    const float c = cos32(tw), s = sinw32(tw);
                    re[i1] = fmaf(di, s, dr * c);
                    im[i1] = fmaf(-dr, s, di * c);
                }
            }
        }
    }
}

constexpr int kPlaneStride = 33;                 // padded row stride of the transpose plane
constexpr int kPlaneFloats = 32 * kPlaneStride;  // 1056 floats = 4224 B per warp

// Forward complex FFT of 1024 points spread over one warp.
//   in : re[n1], im[n1] = x[32*n1 + lane]            (natural register order)
//   out: re[i],  im[i]  = X[lane + 32*brev5(i)]      (bit-reversed register order)
// `plane` is this warp's private 32x33-float transpose buffer, `tw` the CTA-wide table
// tw[k1*32 + lane] = (cos, -sin)(2*pi*k1*lane/1024).
// The inverse transform is obtained by calling it with the two arrays swapped.
__device__ __forceinline__ void warp_fft1024(float (&re)[32], float (&im)[32], float* __restrict__ plane,
                                             const float2* __restrict__ tw, int lane) {
    fft32_dif(re, im);
#pragma unroll
    for (int i = 1; i < 32; ++i) {
        const float2 w = tw[brev5(i) * 32 + lane];
        const float r = re[i], q = im[i];
        re[i] = fmaf(-q, w.y, r * w.x);
        im[i] = fmaf(r, w.y, q * w.x);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) plane[brev5(i) * kPlaneStride + lane] = re[i];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) re[j] = plane[lane * kPlaneStride + j];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; ++i) plane[brev5(i) * kPlaneStride + lane] = im[i];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) im[j] = plane[lane * kPlaneStride + j];
    __syncwarp();
    fft32_dif(re, im);
}

// CTA-wide tables: twiddles (32x32 float2) and the sqrt-Hann window sin(pi*i/1024) (1024 floats).
__device__ __forceinline__ void init_fft_tables(float2* tw, float* win, int tid, int nthreads) {
    for (int e = tid; e < 1024; e += nthreads) {
        const int k1 = e >> 5, l = e & 31;
        double s, c;
        sincospi((double)(k1 * l) / 512.0, &s, &c);     // 2*pi*k1*l/1024
        tw[e] = make_float2((float)c, (float)(-s));
        double ws, wc;
        sincospi((double)e / 1024.0, &ws, &wc);          // sqrt(hann_periodic(1024))[e] = sin(pi e/1024)
        win[e] = (float)ws;
    }
}

}  // namespace ddsp
