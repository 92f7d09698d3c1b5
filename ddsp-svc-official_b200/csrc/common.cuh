// common.cuh -- shared device helpers for the ddsp_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddsp {

constexpr int kHop = 512;          // block_size of every shipped config (configs/*.yaml:6)
constexpr int kFrame = 2 * kHop;   // 50 %-overlap analysis frame
constexpr unsigned kFullMask = 0xffffffffu;

// fl32(2*pi): `2 * np.pi * rot` multiplies an fp32 tensor by a python float (vocoder.py:392,451,517)
#define DDSP_TWO_PI_F 6.28318530717958647692f
#define DDSP_PI_F 3.14159265358979323846f
#define DDSP_LOG2E_F 1.44269504088896340736f

// torch upsample_linear1d(align_corners=True) arithmetic for one sample (core.py:17):
// fma(w0, x0, fl32(w1*x1)) with w1 = j/hop (exact for power-of-two hop), w0 = 1 - w1.
__device__ __forceinline__ float lerp_torch(float x0, float x1, float w1) {
    const float w0 = __fsub_rn(1.0f, w1);
    return __fmaf_rn(w0, x0, __fmul_rn(w1, x1));
}

// Counter-based uniform noise used when no `noise_u` tensor is injected.
// lowbias32 finaliser over (sample index, per-clip key); returns U in [0,1) on a 2^-24 grid.
__host__ __device__ __forceinline__ uint32_t noise_key(uint64_t seed, uint32_t clip) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(clip + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32) ^ (uint32_t)z;
}
__host__ __device__ __forceinline__ float noise_uniform(uint32_t key, uint32_t t) {
    uint32_t x = t * 0x9E3779B1u + key;
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return (float)(x >> 8) * 5.9604644775390625e-8f;   // 2^-24
}

// torch.sinc (1 at 0, sin(pi x)/(pi x)) evaluated from the fp32 argument; sinpif is exact-range-
// reduced so large |x| (up to sr*0.5/f0) costs nothing extra.
__device__ __forceinline__ float sinc_f(float x) {
    const float px = DDSP_PI_F * x;
    const float s = sinpif(x);
    return (x == 0.0f) ? 1.0f : __fdividef(s, px);
}

// rot (fp32, wrapped to [-0.5,0.5], half-to-even) from an fp64 rotation count (core.py:46-49)
__device__ __forceinline__ float wrap_rot(double c) { return (float)(c - rint(c)); }

}  // namespace ddsp
