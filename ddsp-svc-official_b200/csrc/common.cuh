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

// Packed fp32x2 arithmetic (sm_100+: FADD2 / FMUL2 / FFMA2, one issue slot for two IEEE-rn results).
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }   // folds into an operand modifier
__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }            // folds into a broadcast immediate
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return fma2(b, bc2(-1.0f), a); }

// Run partition shared by the run-based kernels (combsubfast_kernel, ltv_conv*_kernel): a clip is cut into `R` runs of
// consecutive units (frame pairs / frames), the first `rem` of them one unit longer than the rest.
__host__ __device__ __forceinline__ int run_begin(int r, int len, int rem) { return r * len + (r < rem ? r : rem); }
// Warp slot -> run.  Slots are numbered wid * gridDim.x + blockIdx.x, so the four warps a scheduler owns (wid, wid + 4,
// wid + 8, wid + 12) hold slots a multiple of 4 * gridDim.x apart.  All LONG runs come first in slot order and the short
// ones after them: every scheduler then gets its long runs in its low warps and no scheduler carries more than
// ceil(4 * long share) of them.  (Dealing run = clip * R + r straight onto the slots put 16 long runs on 100 of the 148
// CTAs and 16 short ones on the other 48 at the headline shape, because 37 runs per clip divides 148.)
__device__ __forceinline__ void slot_to_run(int64_t slot, int B, int R, int rem, int& b, int& r) {
    const int64_t n_long = (int64_t)B * rem;
    if (slot < n_long) {
        b = (int)(slot / rem); r = (int)(slot % rem);
    } else {
        const int64_t i = slot - n_long;
        const int S = R - rem;
        b = (int)(i / S); r = rem + (int)(i % S);
    }
}

// ---- bulk asynchronous copies global -> shared (cp.async.bulk: the TMA engine without a tensor map) -----------------
// One lane issues a copy of a 16-byte-aligned range; completion is counted in bytes on an mbarrier the consumers wait on
// (phase parity = number of completed uses & 1).  No register is held while the data is in flight and no lane spends
// issue slots on it, which is what the instruction-bound FFT kernels need: the copies are started where the warp's
// transpose plane falls idle (between the two passes of a transform, fft32.cuh) and land while the second pass runs.
__device__ __forceinline__ uint32_t smem_addr(const volatile void* p) {
    return (uint32_t)__cvta_generic_to_shared(const_cast<const void*>(p));
}
__device__ __forceinline__ void bulk_mbar_init(uint32_t bar) {         // one arrival (the issuing lane's expect_tx) per phase
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "BULK_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra BULK_DONE_%=;\n\t"
        "bra BULK_WAIT_%=;\n\t"
        "BULK_DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// generic-proxy accesses of the warp (made visible to the issuing lane by a __syncwarp) before async-proxy writes
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// torch upsample_linear1d(align_corners=True) arithmetic for one sample (core.py:17):
// fma(w0, x0, fl32(w1*x1)) with w1 = j/hop (exact for power-of-two hop), w0 = 1 - w1.
__device__ __forceinline__ float lerp_torch(float x0, float x1, float w1) {
    const float w0 = __fsub_rn(1.0f, w1);
    return __fmaf_rn(w0, x0, __fmul_rn(w1, x1));
}

// Counter-based uniform noise used when no `noise_u` tensor is injected.
// A clip owns a 64-bit key (splitmix64 of seed and clip index): the low word enters the per-(hop, lane) stream seed
// linearly (so that a streaming hop offset folds into it on the host), the high word is added between the two
// multiply-xorshift rounds of the hash.  Two clips draw the same (shifted) streams only if BOTH words line up,
// i.e. with probability ~2^-64 per pair instead of ~2^-32 * frames with a single additive key.
__host__ __device__ __forceinline__ uint64_t noise_key64(uint64_t seed, uint32_t clip) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(clip + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// Noise stream layout: the 16 samples {hop*512 + 32*i + lane, i = 0..15} that one lane feeds into
// the FFT form one short multiplicative-congruential stream seeded by a strong hash of (clip key, hop,
// lane).  The state keeps its low byte clear: state = 256*s with s an odd 24-bit number, and
// state <- 747796405 state mod 2^32 is s <- 747796405 s mod 2^24.  The draw is s itself (top 24 bits of
// the state), so the centred value is the state read as a signed integer -- one IMAD + one exact
// int->float conversion per sample, no shift.
__host__ __device__ __forceinline__ uint32_t noise_seed(uint32_t key, uint32_t key2, uint32_t hop, uint32_t lane) {
    uint32_t x = (hop * 32u + lane) * 0x9E3779B1u + key;
    x ^= x >> 16; x *= 0x7feb352du;
    x += key2;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return (x & 0xffffff00u) | 0x100u;
}
__host__ __device__ __forceinline__ uint32_t noise_next(uint32_t state) { return state * 747796405u; }
// uniform integer in [0, 2^24) from a state: its top 24 bits; U = value * 2^-24
__host__ __device__ __forceinline__ uint32_t noise_u24(uint32_t state) { return (state >> 8) ^ 0x800000u; }
// the same draw as a centred integer v = u24 - 2^23 in [-2^23, 2^23):  2U - 1 == v * 2^-23 exactly
__host__ __device__ __forceinline__ int32_t noise_s24(uint32_t state) { return (int32_t)state >> 8; }
// 2^8 * v as fp32 (exact: the low byte of the state is clear):  2U - 1 == noise_f31(state) * 2^-31
__device__ __forceinline__ float noise_f31(uint32_t state) { return (float)(int32_t)state; }

// Single-MUFU approximations (flush-to-zero forms: no denormal fix-up code around them).
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sin_approx(float x) {
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cos_approx(float x) {
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// torch.sinc (1 at 0, sin(pi x)/(pi x)) evaluated from the fp32 argument.
// x = n + r with n = rint(x), |r| <= 0.5 (exact in fp32):  sinc(x) = (-1)^n * (r/x) * S(r^2),
// S(u) = sin(pi r)/(pi r) as a degree-4 interpolant in u = r^2 (|error| < 5e-9 on |r| <= 0.5).
// For |x| < 0.5 the quotient r/x is exactly 1, so there is no cancellation at the pulse peak.
__device__ __forceinline__ float sinc_f(float x) {
    // n = rint(x) through the 1.5*2^23 magic constant: the low mantissa bit of t is the parity of n
    // (exact for |x| < 2^22; beyond that sinc(x) < 8e-8 and the result stays below 1e-6).
    const float xs = x + 1e-30f;                   // x == 0 -> tiny, so that r/x below is 1 instead of 0/0
    const float t = xs + 12582912.0f;
    const float n = t - 12582912.0f;
    const float r = xs - n;
    const float u = r * r;
    float p = 0.024718644097447395f;
    p = fmaf(p, u, -0.19044175744056702f);
    p = fmaf(p, u, 0.8117148280143738f);
    p = fmaf(p, u, -1.6449332237243652f);
    p = fmaf(p, u, 1.0f);
    const float v = p * (r * rcp_approx(xs));
    return __int_as_float(__float_as_int(v) ^ (__float_as_int(t) << 31));   // (-1)^n
}

// Two samples at once: the same arithmetic as sinc_f on both halves of packed registers (the
// reciprocals, the sign flip and nothing else stay scalar).
__device__ __forceinline__ float2 sinc2_xs(float2 xs) {          // xs = x + 1e-30 already formed
    const float2 t = add2(xs, bc2(12582912.0f));
    const float2 n = add2(t, bc2(-12582912.0f));
    const float2 r = sub2(xs, n);
    const float2 u = mul2(r, r);
    float2 p = fma2(bc2(0.024718644097447395f), u, bc2(-0.19044175744056702f));
    p = fma2(p, u, bc2(0.8117148280143738f));
    p = fma2(p, u, bc2(-1.6449332237243652f));
    p = fma2(p, u, bc2(1.0f));
    const float2 v = mul2(p, mul2(r, make_float2(rcp_approx(xs.x), rcp_approx(xs.y))));
    return make_float2(__int_as_float(__float_as_int(v.x) ^ (__float_as_int(t.x) << 31)),
                       __int_as_float(__float_as_int(v.y) ^ (__float_as_int(t.y) << 31)));
}

// rot (fp32, wrapped to [-0.5,0.5], half-to-even) from an fp64 rotation count (core.py:46-49)
__device__ __forceinline__ float wrap_rot(double c) { return (float)(c - rint(c)); }

}  // namespace ddsp
