// ltvfir.cuh -- `frequency_filter` (ddsp/core.py:331-336): per-frame linear-phase LTV-FIR applied
// through the frequency domain, as ONE kernel per filter:
//
//   _frequency_impulse_response  core.py:306-328   magnitudes -> irfft_L -> roll -> window
//   _fft_convolve                core.py:185-239   Bartlett framing, FFT convolution, overlap-add,
//                                                   crop with L//2 delay compensation
//
// Two kernels per filter, five 1024-point complex FFTs per frame in total (all through the
// warp_fft1024 of fft32.cuh).  `ltv_ir_kernel` (one warp per frame, frames independent) does steps
// 2-4 and leaves the raw tap spectrum (1024 complex) in a workspace; `ltv_conv_kernel` (a warp owns
// a run of consecutive frames of one clip) does steps 1 and 5.  Splitting keeps both kernels at 16
// warps per SM (shared memory per warp 4.3 KB / 12.5 KB) and halves the code each warp walks through:
//   1. audio frame (1024 Bartlett-windowed samples, zero-padded to 2048): real FFT-2048 computed as
//      the complex FFT-1024 of even/odd samples -> kept raw in shared memory;
//   2./3. impulse response: the L-point inverse real DFT (L = 510 or 1022, NOT a power of two) is
//      evaluated exactly with Bluestein's chirp-z identity as a 1024-point circular convolution
//      (forward FFT, multiply with the precomputed chirp spectrum, inverse FFT); then the roll and
//      the window (none / Hann / dynamic cosine) are applied while writing the taps to shared memory;
//   4. the taps' real FFT-2048 (again as complex FFT-1024 of even/odd taps);
//   5. spectra multiplied in the even/odd domain (so no second conjugate-pair exchange is needed)
//      and one inverse FFT-1024 yields the 2048 output samples of the frame, which are overlap-added
//      in a shared-memory ring; every frame retires 512 finished samples to HBM.
// The reference's FFT size (1533 / 2045) only has to cover the linear convolution; 2048 gives the
// identical result.
// L = 510 filters take a shorter route on the convolution side (ltv_conv510_kernel below): the two
// halves of the Bartlett frame ride as real/imaginary part of one FFT-1024 and the taps' spectrum is
// stored as H[k] itself.
#pragma once
#include "fft32.cuh"

namespace ddsp {

#ifndef LTV_STAGGER_NS
#define LTV_STAGGER_NS 2000u
#endif
constexpr int kLtvWarps = 16;                      // warps per CTA, both kernels
constexpr int kLtvThreads = kLtvWarps * 32;
constexpr int kLtvRing = 2048;                     // floats: overlap-add ring of the convolution kernel
constexpr int kLtvCtxInts = 8;                     // cold per-warp scalars parked in shared memory (see combsubfast.cuh)
constexpr int kLtvChirpBytes = (512 * 2 + 1024 * 2) * 4;   // c[m] (512 complex) + chirp spectrum (1024 complex), parked in shared memory
constexpr int kLtvIrSmemBytes = 512 * 16 + kLtvChirpBytes + kLtvWarps * (kPlaneFloats * 4 + kLtvCtxInts * 4);
constexpr int kLtvConvWarpBytes = kPlaneFloats * 4 + kLtvRing * 4 + kLtvCtxInts * 4;
constexpr int kLtvConvSmemBytes = 512 * 16 + kLtvWarps * kLtvConvWarpBytes;
constexpr int kLtvSpecFloat2 = 1024;               // workspace per frame: raw FFT-1024 of the even/odd-packed taps

// per-L chirp tables: c[m] = exp(i*pi*m^2/L), m < 512; dhat = FFT_1024 of the wrapped conjugate chirp
constexpr int kChirpFloats = 512 * 2 + 1024 * 2;

struct LtvParams {
    const float* audio;                 // (B,T) contiguous, or U when audio_mode==1, unused when 2
    int audio_mode;                     // 0: samples, 1: uniform U -> 2U-1 (vocoder.py:418,545), 2: in-kernel noise
    uint64_t seed;
    uint32_t key_offset;                // streaming: noise key shift of the block's first hop (0 otherwise)
    const float* mags; int64_t mB, mF;  // (B,F,n_mag) control / magnitude rows
    int n_mag, encoding; float mag_scale;
    int window_mode;
    const float* f0_frames; int64_t fB, fF; float sr15;   // dynamic window: hw = 1.5*sr/(f0+1e-3)
    const float* tw_tables;             // twiddles (fft32.cuh)
    const float* chirp;                 // tables for this L
    float2* spec;                       // workspace (B,F,1024) complex: tap spectra
    float* out;                         // (B,T)
    const float* add_in; float* sum_out; // optional (L = 510 convolution): sum_out = add_in + out (vocoder.py:421,548)
    int B, F, run_len, run_rem, runs_per_clip;   // run r of a clip: frames [run_begin(r), + run_len + (r < run_rem))
};

__device__ __forceinline__ float bartlett1024(int i) {      // torch.bartlett_window(1024), periodic (core.py:221)
    return (i <= 512) ? (float)i * (1.0f / 512.0f) : (float)(1024 - i) * (1.0f / 512.0f);
}

// Conjugate-symmetric partner of bin k = lane + 32 q of a 1024-point spectrum held in registers:
// bin 1024-k lives in lane 32-l, register-index 31-q (lane 0: its own register (32-q)&31).
#define LTV_PARTNER(X, q, outr, outi)                                              \
    {                                                                              \
        outr = __shfl_sync(kFullMask, DDSP_RE(X, 31 - (q)), partner);              \
        outi = __shfl_sync(kFullMask, DDSP_IM(X, 31 - (q)), partner);              \
        const float r0_ = DDSP_RE(X, (32 - (q)) & 31), i0_ = DDSP_IM(X, (32 - (q)) & 31); \
        outr = lane0 ? r0_ : outr;                                                 \
        outi = lane0 ? i0_ : outi;                                                 \
    }

// ---------------------------------------------------------------------------------------------
// Kernel 1: magnitudes -> windowed causal impulse response -> its spectrum.  One warp per frame.
// ENC / WIN >= 0 fix the magnitude encoding / window mode at compile time (-1: read from P).
// ---------------------------------------------------------------------------------------------
template <int ENC, int WIN, int NMAG = 0>     // NMAG > 0 fixes n_mag at compile time (drops the per-element range tests)
__global__ void __launch_bounds__(kLtvThreads, 1) ltv_ir_kernel(const LtvParams P) {
    const int enc = ENC >= 0 ? ENC : P.encoding;
    const int win_mode = WIN >= 0 ? WIN : P.window_mode;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const float4* tw4 = reinterpret_cast<const float4*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float2* chirp_s = reinterpret_cast<const float2*>(smem_raw + 512 * 16);      // [0,512): c[m]; [512,1536): chirp spectrum
    float* plane = reinterpret_cast<float*>(smem_raw + 512 * 16 + kLtvChirpBytes) + wid * (kPlaneFloats + kLtvCtxInts);
    volatile int* ctx = reinterpret_cast<volatile int*>(plane + kPlaneFloats);
    {
        const float4* src = reinterpret_cast<const float4*>(P.tw_tables);
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        for (int e = threadIdx.x; e < 512; e += kLtvThreads) dst[e] = __ldg(src + e);
        const float4* csrc = reinterpret_cast<const float4*>(P.chirp);
        float4* cdst = reinterpret_cast<float4*>(smem_raw + 512 * 16);
        for (int e = threadIdx.x; e < kLtvChirpBytes / 16; e += kLtvThreads) cdst[e] = __ldg(csrc + e);
        __syncthreads();
    }
#define IR_NMAG (NMAG > 0 ? NMAG : P.n_mag)
#define IR_L (2 * (IR_NMAG - 1))
#define IR_D (IR_NMAG - 1)
    const int64_t n_frames = (int64_t)P.B * P.F;

    __nanosleep((unsigned)(wid >> 2) * LTV_STAGGER_NS);
    Pts32 X;
    for (int64_t fr0 = (int64_t)blockIdx.x * kLtvWarps + wid; fr0 < n_frames; fr0 += (int64_t)gridDim.x * kLtvWarps) {
        if (lane == 0) { ctx[0] = (int)(fr0 / P.F); ctx[1] = (int)(fr0 % P.F); }
        __syncwarp();
#pragma unroll 1
        for (int phase = 0; phase < 3; ++phase) {
            if (phase == 0) {
                // a'[k] = w_k X[k] c[k] / (L*1024), k = 32 n1 + lane < n_mag   (irfft, core.py:316)
                const int n_mag = IR_NMAG, L = IR_L;
                const float2* chirp_c = chirp_s;
                const float* row = P.mags + (int64_t)ctx[0] * P.mB + (int64_t)ctx[1] * P.mF;
                const float scale = 1.0f / ((float)L * 1024.0f);
                float carry = 0.0f;                                      // allpass: running phase in turns
                // the whole control row first (all loads in flight at once), then the arithmetic
                float craw[16];
                float2 ccplx[16];
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    craw[n1] = 0.0f; ccplx[n1] = make_float2(0.0f, 0.0f);
                    if (n1 * 32 < n_mag) {
                        if (enc == DDSP_B200_MAG_COMPLEX) ccplx[n1] = __ldg(reinterpret_cast<const float2*>(row) + 32 * n1 + lane);
                        else craw[n1] = __ldg(row + 32 * n1 + lane);
                    }
                }
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    float xr = 0.0f, xi = 0.0f;
                    if (n1 < 16 && n1 * 32 < n_mag) {
                        const int k = 32 * n1 + lane;
                        if (enc == DDSP_B200_MAG_ALLPASS_TANH) {
                            // exp(j*cumsum(pi*tanh(c)))  (vocoder.py:398,415 / 521,540), phase kept in turns
                            float g = 0.5f * tanhf(craw[n1 & 15]);
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                const float t = __shfl_up_sync(kFullMask, g, d);
                                if (lane >= d) g += t;
                            }
                            g += carry;
                            carry = __shfl_sync(kFullMask, g, 31);
                            carry -= rintf(carry);
                            g -= rintf(g);
                            xr = cos_approx(DDSP_TWO_PI_F * g);
                            xi = sin_approx(DDSP_TWO_PI_F * g);
                        } else if (enc == DDSP_B200_MAG_EXP) {
                            xr = ex2_approx(craw[n1 & 15] * DDSP_LOG2E_F) * P.mag_scale;   // vocoder.py:399,475,522-523
                        } else if (enc == DDSP_B200_MAG_COMPLEX) {
                            xr = ccplx[n1 & 15].x; xi = ccplx[n1 & 15].y;
                        } else {
                            xr = craw[n1 & 15];
                        }
                        // DC and Nyquist: imaginary part ignored, weight 1; interior bins weight 2
                        const bool edge = (k == 0) || (k == n_mag - 1);
                        const float wgt = edge ? scale : 2.0f * scale;
                        xi = edge ? 0.0f : xi;
                        const float2 c = chirp_c[k];
                        const float ar = xr * wgt, ai = xi * wgt;
                        xr = ar * c.x - ai * c.y;
                        xi = ar * c.y + ai * c.x;
                    }
                    DDSP_RE(X, brev5(n1)) = xr;
                    DDSP_IM(X, brev5(n1)) = xi;
                }
            } else if (phase == 2) {
                const int D = IR_D;
                if (IR_L == 510) {
                    // L = 510: z[n] = h[n] (n < 512, imaginary part 0) -> the stored spectrum is H[k] itself,
                    // which ltv_conv510_kernel multiplies with the packed half-frames (no even/odd split)
#pragma unroll
                    for (int n1 = 0; n1 < 32; ++n1) {
                        DDSP_RE(X, brev5(n1)) = (n1 < 16) ? plane[32 * (n1 & 15) + lane] : 0.0f;
                        DDSP_IM(X, brev5(n1)) = 0.0f;
                    }
                } else {
                    // z[n] = h[2n] + j h[2n+1] from the taps in `plane`
                    const float2* h2 = reinterpret_cast<const float2*>(plane);
#pragma unroll
                    for (int n1 = 0; n1 < 32; ++n1) {
                        float2 v = make_float2(0.0f, 0.0f);
                        if (n1 < 16 && 32 * n1 < D + 1) v = h2[32 * n1 + lane];
                        DDSP_RE(X, brev5(n1)) = v.x;
                        DDSP_IM(X, brev5(n1)) = v.y;
                    }
                }
                __syncwarp();
            }

            warp_fft1024(X, plane, tw4, lane);

            if (phase == 0) {
                // times the chirp spectrum, then inverse FFT (real/imag swapped through the forward FFT)
                const float2* chirp_d = chirp_s + 512;
                float pr[32], pi[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const float2 dh = chirp_d[lane + 32 * q];
                    const float ur = DDSP_RE(X, q), ui = DDSP_IM(X, q);
                    pr[q] = ur * dh.x - ui * dh.y;
                    pi[q] = ur * dh.y + ui * dh.x;
                }
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    DDSP_RE(X, brev5(q)) = pi[q];
                    DDSP_IM(X, brev5(q)) = pr[q];
                }
            } else if (phase == 1) {
                // conv[n] (Re in X.im, Im in X.re after the swapped FFT), n = lane + 32 q < n_out:
                // ir_zero_phase[n] = Re(c[n] * conv[n]); causal form + window (core.py:242-303,326)
                const int L = IR_L, D = IR_D;
                const int n_out = (L == 510) ? 510 : 512;          // IR samples produced by the chirp convolution
                const bool sym = L != 510;                          // L=1022: real magnitudes -> IR symmetric, mirror it
                const float2* chirp_c = chirp_s;
                const float two_pi_over_L = DDSP_TWO_PI_F / (float)L;
                float hw_inv = 0.0f;
                if (win_mode == DDSP_B200_WINDOW_DYNAMIC) {
                    const float f0 = __ldg(P.f0_frames + (int64_t)ctx[0] * P.fB + (int64_t)ctx[1] * P.fF);
                    hw_inv = __fdiv_rn(1.0f, __fdiv_rn(P.sr15, __fadd_rn(f0, 1e-3f)));
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int n = lane + 32 * q;
                    if (n < n_out) {
                        const float2 c = chirp_c[n];
                        const float ir = DDSP_IM(X, q) * c.x - DDSP_RE(X, q) * c.y;   // Re(c * conv)
                        // tap positions: i = (n + D) mod L holds lag +n; for symmetric IRs also lag -n
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            int lag;
                            if (side == 0) lag = (n <= L - D - 1) ? n : n - L;
                            else {
                                if (!sym || n == 0 || n >= D) continue;
                                lag = -n;
                            }
                            const int i = lag + D;
                            float w = 1.0f;
                            if (win_mode == DDSP_B200_WINDOW_HANN) {
                                // 0.5 - 0.5 cos(2 pi i / L) = 0.5 + 0.5 cos(2 pi i / L - pi), argument in [-pi, pi)
                                w = fmaf(0.5f, cos_approx(fmaf((float)i, two_pi_over_L, -DDSP_PI_F)), 0.5f);
                            } else if (win_mode == DDSP_B200_WINDOW_DYNAMIC) {
                                float x = (float)lag * hw_inv;
                                x = (x > 1.0f) ? 0.0f : x;                           // core.py:297 (only x>1 is cleared)
                                x -= 2.0f * rintf(0.5f * x);                         // cos(pi x) has period 2: |x| <= 1
                                w = fmaf(0.5f, cos_approx(DDSP_PI_F * x), 0.5f);
                            }
                            // L=1022: 1/4 (even/odd split) * 1/1024 (inverse FFT); L=510: 1/1024 only
                            plane[i] = ir * w * (sym ? (1.0f / 4096.0f) : (1.0f / 1024.0f));
                        }
                    }
                }
                if (lane < 2) plane[L + lane] = 0.0f;
                __syncwarp();
            } else {
                float2* dst = P.spec + ((int64_t)ctx[0] * P.F + ctx[1]) * kLtvSpecFloat2 + lane;
                // L = 510: H[k] of real taps is Hermitian -- only bins 0..512 are stored (ltv_conv510_kernel
                // mirrors the rest); L = 1022: the raw even/odd-packed spectrum needs all 1024 bins
                const bool half = IR_NMAG == 256;
#pragma unroll
                for (int q = 0; q < 32; ++q)
                    if (!half || q < 16 || (q == 16 && lane == 0)) dst[32 * q] = make_float2(DDSP_RE(X, q), DDSP_IM(X, q));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel 1b: the all-pass (group delay, no window) and the noise (exp/128, Hann) impulse responses
// of one frame -- both L = 510 -- from ONE Bluestein pass.  With the full Hermitian extension of
// both spectra, V[k] = Xa[k] + j Xn[k] (k = 0..509) has the real inverse DFT  ir_a[n] + j ir_n[n],
// and 510 + 510 - 1 <= 1024 still fits the circular convolution; the two real tap sets then share ONE 1024-point FFT
// (packed as real / imaginary part and split by conjugate symmetry).  3 FFTs for two filters instead of 6.  Used by CombSub (vocoder.py:540,545-546) and Sins (:415,418-419).
// ---------------------------------------------------------------------------------------------
constexpr int kLtvDualWarpFloats = kPlaneFloats + kLtvCtxInts;           // plane (also holds the packed taps) + ctx
constexpr int kLtvDualSmemBytes = 512 * 16 + kLtvChirpBytes + kLtvWarps * kLtvDualWarpFloats * 4;

struct LtvDualParams {
    const float* gd; const float* nm; int64_t mB, mF;   // group_delay / noise_magnitude control rows (B,F,256)
    float noise_scale;                                   // 1/128
    const float* tw_tables; const float* chirp_c;        // c[m], m < 512 (L = 510)
    const float* chirp_d_dual;                           // FFT_1024 of the wrapped conjugate chirp, m in [-509, 509]
    float2* spec_a; float2* spec_n;                      // (B,F,1024) complex each
    int B, F;
};

__global__ void __launch_bounds__(kLtvThreads, 1) ltv_ir_dual_kernel(const LtvDualParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const float4* tw4 = reinterpret_cast<const float4*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* plane = reinterpret_cast<float*>(smem_raw + 512 * 16 + kLtvChirpBytes) + wid * kLtvDualWarpFloats;
    volatile int* ctx = reinterpret_cast<volatile int*>(plane + kPlaneFloats);
    // chirp tables parked in shared memory: c[m] (512 complex), then the 1024-bin chirp spectrum
    const float2* chirp_c = reinterpret_cast<const float2*>(smem_raw + 512 * 16);
    const float2* chirp_d = chirp_c + 512;
    {
        const float4* src = reinterpret_cast<const float4*>(P.tw_tables);
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        for (int e = threadIdx.x; e < 512; e += kLtvThreads) dst[e] = __ldg(src + e);
        float4* cdst = reinterpret_cast<float4*>(smem_raw + 512 * 16);
        const float4* c1 = reinterpret_cast<const float4*>(P.chirp_c);
        const float4* c2 = reinterpret_cast<const float4*>(P.chirp_d_dual);
        for (int e = threadIdx.x; e < 256; e += kLtvThreads) cdst[e] = __ldg(c1 + e);
        for (int e = threadIdx.x; e < 512; e += kLtvThreads) cdst[256 + e] = __ldg(c2 + e);
        __syncthreads();
    }
    constexpr int L = 510, D = 255, NM = 256;
    const int64_t n_frames = (int64_t)P.B * P.F;

    __nanosleep((unsigned)(wid >> 2) * LTV_STAGGER_NS);
    Pts32 X;
    for (int64_t fr0 = (int64_t)blockIdx.x * kLtvWarps + wid; fr0 < n_frames; fr0 += (int64_t)gridDim.x * kLtvWarps) {
        if (lane == 0) { ctx[0] = (int)(fr0 / P.F); ctx[1] = (int)(fr0 % P.F); }
        __syncwarp();
#pragma unroll 1
        for (int phase = 0; phase < 3; ++phase) {
            if (phase == 0) {
                const int64_t ro = (int64_t)ctx[0] * P.mB + (int64_t)ctx[1] * P.mF;
                const float* rg = P.gd + ro;
                const float* rn = P.nm + ro;
                float cg[8], cn[8];
#pragma unroll
                for (int n1 = 0; n1 < 8; ++n1) { cg[n1] = __ldg(rg + 32 * n1 + lane); cn[n1] = __ldg(rn + 32 * n1 + lane); }
                // lower half k = 32 n1 + lane < 256: all-pass phase by inclusive scan (in turns), noise magnitude;
                // both are parked in `plane` (as float2 (theta, mag)) for the mirrored upper half
                float2* park = reinterpret_cast<float2*>(plane);
                float carry = 0.0f;
#pragma unroll
                for (int n1 = 0; n1 < 8; ++n1) {
                    float g = 0.5f * tanhf(cg[n1]);
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const float t = __shfl_up_sync(kFullMask, g, d);
                        if (lane >= d) g += t;
                    }
                    g += carry;
                    carry = __shfl_sync(kFullMask, g, 31);
                    carry -= rintf(carry);
                    g -= rintf(g);
                    park[32 * n1 + lane] = make_float2(g, ex2_approx(cn[n1] * DDSP_LOG2E_F) * P.noise_scale);
                }
                __syncwarp();
                const float scale = 1.0f / ((float)L * 1024.0f);
                float vr[16], vi[16];
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    const int k = 32 * n1 + lane;
                    vr[n1] = 0.0f; vi[n1] = 0.0f;
                    if (k < L) {
                        const int km = (k < NM) ? k : L - k;              // Hermitian mirror (1..254 for k >= 256)
                        const float2 pm = park[km];
                        float xr = cos_approx(DDSP_TWO_PI_F * pm.x), xi = sin_approx(DDSP_TWO_PI_F * pm.x);
                        if (k >= NM) xi = -xi;                            // conj
                        if (k == 0 || k == NM - 1) xi = 0.0f;             // irfft ignores Im of DC / Nyquist
                        // V = Xa + j Xn  (Xn real)
                        const float ar = xr * scale, ai = (xi + pm.y) * scale;
                        const float2 c = chirp_c[k];
                        vr[n1] = ar * c.x - ai * c.y;
                        vi[n1] = ar * c.y + ai * c.x;
                    }
                }
                __syncwarp();                                            // park (plane) is dead before the FFT reuses it
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    DDSP_RE(X, brev5(n1)) = (n1 < 16) ? vr[n1 & 15] : 0.0f;
                    DDSP_IM(X, brev5(n1)) = (n1 < 16) ? vi[n1 & 15] : 0.0f;
                }
            } else if (phase == 2) {
                const float2* h2 = reinterpret_cast<const float2*>(plane);
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    float2 v = make_float2(0.0f, 0.0f);
                    if (n1 < 16) v = h2[32 * n1 + lane];
                    DDSP_RE(X, brev5(n1)) = v.x;
                    DDSP_IM(X, brev5(n1)) = v.y;
                }
                __syncwarp();
            }

            warp_fft1024(X, plane, tw4, lane);

            if (phase == 0) {
                float pr[32], pi[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const float2 dh = chirp_d[lane + 32 * q];
                    const float ur = DDSP_RE(X, q), ui = DDSP_IM(X, q);
                    pr[q] = ur * dh.x - ui * dh.y;
                    pi[q] = ur * dh.y + ui * dh.x;
                }
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    DDSP_RE(X, brev5(q)) = pi[q];
                    DDSP_IM(X, brev5(q)) = pr[q];
                }
            } else if (phase == 1) {
                // S[n] = c[n] * conv[n] = ir_allpass[n] + j ir_noise[n]  (zero-phase form), n = lane + 32 q < 510
                const float two_pi_over_L = DDSP_TWO_PI_F / (float)L;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int n = lane + 32 * q;
                    if (n < L) {
                        const float2 c = chirp_c[n];
                        const float cr = DDSP_IM(X, q), ci = DDSP_RE(X, q);      // conv = cr + j ci (swapped FFT)
                        const float sa = cr * c.x - ci * c.y, sn = cr * c.y + ci * c.x;
                        const int lag = (n <= L - D - 1) ? n : n - L;
                        const int i = lag + D;
                        const float w = fmaf(0.5f, cos_approx(fmaf((float)i, two_pi_over_L, -DDSP_PI_F)), 0.5f);   // Hann (core.py:262)
                        // packed taps z[i] = h_allpass[i] + j h_noise[i]; 1/1024 (inverse FFT of the convolution)
                        // * 1/2 (split of the two spectra below).  All-pass: no window (core.py:326)
                        reinterpret_cast<float2*>(plane)[i] = make_float2(sa * (1.0f / 2048.0f), sn * w * (1.0f / 2048.0f));
                    }
                }
                if (lane < 2) reinterpret_cast<float2*>(plane)[L + lane] = make_float2(0.0f, 0.0f);
                __syncwarp();
            } else {
                // one FFT carried both real tap sets: H_a[k] = Z[k] + conj Z[N-k],  H_n[k] = (Z[k] - conj Z[N-k]) / j
                const int partner = (32 - lane) & 31;
                const bool lane0 = lane == 0;
                float2* da = P.spec_a + ((int64_t)ctx[0] * P.F + ctx[1]) * kLtvSpecFloat2 + lane;
                float2* dn = P.spec_n + ((int64_t)ctx[0] * P.F + ctx[1]) * kLtvSpecFloat2 + lane;
                // both spectra are Hermitian (real taps): only bins 0..512 are stored (registers 0..15 of every
                // lane, bin 512 = lane 0's register 16); ltv_conv510_kernel mirrors the upper half
#pragma unroll
                for (int q = 0; q < 17; ++q) {
                    float c, d;
                    LTV_PARTNER(X, q, c, d);
                    const float a = DDSP_RE(X, q), b = DDSP_IM(X, q);
                    if (q < 16 || lane0) {
                        da[32 * q] = make_float2(a + c, b - d);
                        dn[32 * q] = make_float2(b + d, c - a);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel 2: framing, FFT convolution with the stored tap spectra, overlap-add.  A warp owns a run
// of consecutive frames of one clip.  AMODE >= 0 fixes the audio source at compile time.
// ---------------------------------------------------------------------------------------------
template <int AMODE>
__global__ void __launch_bounds__(kLtvThreads, 1) ltv_conv_kernel(const LtvParams P) {
    const int amode = AMODE >= 0 ? AMODE : P.audio_mode;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const float4* tw4 = reinterpret_cast<const float4*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + 512 * 16 + wid * kLtvConvWarpBytes;
    float* plane = reinterpret_cast<float*>(wbase);
    float* ring = plane + kPlaneFloats;
    volatile int* ctx = reinterpret_cast<volatile int*>(ring + kLtvRing);
    {
        const float4* src = reinterpret_cast<const float4*>(P.tw_tables);
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        for (int e = threadIdx.x; e < 512; e += kLtvThreads) dst[e] = __ldg(src + e);
        __syncthreads();
    }
    {
        const int64_t slot = (int64_t)wid * gridDim.x + blockIdx.x;     // long runs evenly over the schedulers (slot_to_run)
        if (slot >= (int64_t)P.B * P.runs_per_clip) return;
        int b0, r;
        slot_to_run(slot, P.B, P.runs_per_clip, P.run_rem, b0, r);
        const int mb = run_begin(r, P.run_len, P.run_rem);
        if (lane == 0) {
            ctx[0] = b0; ctx[1] = mb; ctx[2] = mb + P.run_len + (r < P.run_rem ? 1 : 0);
            const uint64_t k64 = noise_key64(P.seed, (uint32_t)b0);
            ctx[3] = (int)((uint32_t)k64 + P.key_offset);
            ctx[5] = (int)(uint32_t)(k64 >> 32);
            bulk_mbar_init(smem_addr(ctx + 6));          // ctx[6..7]: mbarrier of the audio copies into the transpose plane
        }
        __syncwarp();
    }
#define CV_B ctx[0]
#define CV_MBEGIN ctx[1]
#define CV_MEND ctx[2]
#define CV_BAR smem_addr(ctx + 6)                                                  // (recomputed where used: no register
#define CV_SKEW ((int)((reinterpret_cast<uint64_t>(P.audio) & 15ull) >> 2))       //  is held across the transforms)
    const int F = P.F;
    const int64_t T = (int64_t)F * kHop;
    constexpr int D = 511;                              // L/2 (this kernel serves n_mag = 512 only): delay compensation (core.py:177)
    const int partner = (32 - lane) & 31;
    const bool lane0 = lane == 0;

    for (int i = lane; i < kLtvRing; i += 32) ring[i] = 0.0f;
    __syncwarp();
    if (P.run_len >= 4) __nanosleep((unsigned)(wid >> 2) * LTV_STAGGER_NS);   // de-phase the warps of a scheduler (see combsubfast.cuh)
    int rs = 0;                                          // ring start (logical sample 0 of the current frame)
    // The frame's 1024 input samples (amode 0 / 1) arrive in the transpose plane by one bulk copy (common.cuh) that lane 0
    // starts inside the INVERSE transform of the previous frame (before the loop for the run's first frame), where the
    // plane falls idle; layout and alignment handling as in ltv_conv510_kernel.  (The 8 KB tap spectrum does not fit
    // beside it and keeps its software-pipelined loads.)
    const bool stage_audio = amode != 2;
    auto start_audio = [&](int mm) {
        const bool hA = mm >= 1, hB = mm < F;
        const float* src = P.audio + (int64_t)CV_B * T + (int64_t)(hA ? mm - 1 : 0) * kHop;
        const uint32_t bytes = (uint32_t)((hA ? 2048 : 0) + (hB ? 2048 : 0) + (CV_SKEW ? 16 : 0));
        const uint32_t bar = CV_BAR;
        fence_proxy_async_smem();
        bulk_mbar_expect(bar, bytes);
        bulk_copy_g2s(smem_addr(plane) + (hA ? 0 : 2048), reinterpret_cast<const void*>(reinterpret_cast<uint64_t>(src) & ~15ull),
                      bytes, bar);
    };
    if (stage_audio && lane0) start_audio(CV_MBEGIN);

    // The frame counter lives in the per-warp shared-memory context (read where used), like the step counter of
    // combsubfast_kernel: ptxas otherwise spills it and the addresses derived from it to local memory, whose reloads miss
    // the tiny L1 left beside the shared memory.
#define CV_M ctx[4]
    if (lane0) CV_M = CV_MBEGIN;
    __syncwarp();
    Pts32 X;
#pragma unroll 1
    for (;;) {
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == 0) {
                const int m = CV_M;
                const float2* zh = P.spec + ((int64_t)CV_B * F + min(m, F - 1)) * kLtvSpecFloat2;
                {   // pull this frame's tap spectrum (8 KB) into L2 while the audio FFT runs
                    const char* pz = reinterpret_cast<const char*>(zh) + 128 * lane;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pz));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pz + 4096));
                }
                // z[n] = a[2n] + j a[2n+1], n = 32 n1 + lane < 512; a = bartlett * frame (core.py:218-222)
                const bool vA = m >= 1, vB = m < F;
                const float* src = plane + CV_SKEW;                     // the staged frame (see start_audio)
                if (stage_audio) bulk_mbar_wait(CV_BAR, (uint32_t)((m - CV_MBEGIN) & 1));
                const uint32_t key = (uint32_t)ctx[3], key2 = (uint32_t)ctx[5];
                uint32_t stA = noise_seed(key, key2, (uint32_t)(m - 1), (uint32_t)lane);
                uint32_t stB = noise_seed(key, key2, (uint32_t)m, (uint32_t)lane);
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    float v0 = 0.0f, v1 = 0.0f;
                    if (n1 < 16) {
                        const int i = 64 * n1 + 2 * lane;               // frame-relative sample index (even)
                        const bool ok = (n1 < 8) ? vA : vB;
                        if (amode == 2) {
                            uint32_t& st = (n1 < 8) ? stA : stB;
                            st = noise_next(st); v0 = noise_f31(st) * 4.6566128730773926e-10f;
                            st = noise_next(st); v1 = noise_f31(st) * 4.6566128730773926e-10f;
                            v0 = ok ? v0 : 0.0f;
                            v1 = ok ? v1 : 0.0f;
                        } else if (ok) {
                            const float2 x = *reinterpret_cast<const float2*>(src + i);
                            v0 = x.x; v1 = x.y;
                            if (amode == 1) { v0 = fmaf(2.0f, v0, -1.0f); v1 = fmaf(2.0f, v1, -1.0f); }
                        }
                        v0 *= bartlett1024(i);
                        v1 *= bartlett1024(i + 1);
                    }
                    DDSP_RE(X, brev5(n1)) = v0;
                    DDSP_IM(X, brev5(n1)) = v1;
                }
                if (stage_audio) __syncwarp();          // every lane has read its samples before the transform reuses the plane
            }

            warp_fft1024(X, plane, tw4, lane, [&] {
                if (lane0 && phase == 1 && stage_audio && CV_M + 1 < CV_MEND) start_audio(CV_M + 1);
            });

            if (phase == 0) {
                // even/odd-domain product:  Zy = (Ea Eh + W1024^k Oa Oh) + j (Ea Oh + Oa Eh)
                const float2* zh = P.spec + ((int64_t)CV_B * F + min(CV_M, F - 1)) * kLtvSpecFloat2;   // last IR repeated (core.py:228)
                const float2 wl = make_float2(tw4[lane].y, tw4[lane].w);        // W1024^lane = (cos, -sin)
                float zr[32], zi[32];
                // Ey and Oy are spectra of REAL sequences (the even / odd output samples), so
                //   Zy[1024-k] = conj(Ey[k]) + j conj(Oy[k]):
                // only bins k <= 512 are multiplied out (registers q = 0..15 of every lane hold k = lane + 32 q
                // < 512; bin 512 is lane 0's register 16), the upper half arrives by one conjugate-pair exchange
                // (bin 1024-k lives in lane 32-l, register 31-q; lane 0: its own register 32-q).
                float pr[17], pi[17];                     // (Re, Im) of Zy[1024-k] for this lane's k = lane + 32 q
                // tap-spectrum loads run kLook bins ahead of their use (software pipeline over the unrolled loop)
                constexpr int kLook = 4;
                float2 zkq[kLook], zpq[kLook];
#pragma unroll
                for (int q = 0; q < kLook; ++q) {
                    zkq[q] = __ldg(zh + lane + 32 * q);
                    zpq[q] = __ldg(zh + ((1024 - (lane + 32 * q)) & 1023));
                }
#pragma unroll
                for (int q = 0; q < 17; ++q) {            // q = 16: bin 512, meaningful on lane 0 only (others: in-bounds dummies)
                    const float2 zk = zkq[q % kLook], zp = zpq[q % kLook];
                    if (q + kLook < 17) {
                        zkq[q % kLook] = __ldg(zh + lane + 32 * (q + kLook));
                        zpq[q % kLook] = __ldg(zh + ((1024 - (lane + 32 * (q + kLook))) & 1023));
                    }
                    float cr, ci;
                    LTV_PARTNER(X, q, cr, ci);
                    const float ar = DDSP_RE(X, q), ai = DDSP_IM(X, q);
                    const float Ear = ar + cr, Eai = ai - ci, Oar = ai + ci, Oai = cr - ar;
                    const float Ehr = zk.x + zp.x, Ehi = zk.y - zp.y, Ohr = zk.y + zp.y, Ohi = zp.x - zk.x;
                    // W1024^k = W1024^lane * W32^q
                    const float cq = (q < 16) ? cos32(q & 15) : -cos32(q & 15);
                    const float sq = (q < 16) ? -sin32(q & 15) : sin32(q & 15);      // W32^q = cq + j sq
                    const float wr = wl.x * cq - wl.y * sq, wi = wl.x * sq + wl.y * cq;
                    const float oor = Oar * Ohr - Oai * Ohi, ooi = Oar * Ohi + Oai * Ohr;
                    const float eer = Ear * Ehr - Eai * Ehi, eei = Ear * Ehi + Eai * Ehr;
                    const float eyr = eer + (wr * oor - wi * ooi), eyi = eei + (wr * ooi + wi * oor);
                    const float oyr = (Ear * Ohr - Eai * Ohi) + (Oar * Ehr - Oai * Ehi);
                    const float oyi = (Ear * Ohi + Eai * Ohr) + (Oar * Ehi + Oai * Ehr);
                    zr[q] = eyr - oyi;          // Ey + j Oy
                    zi[q] = eyi + oyr;
                    pr[q] = eyr + oyi;          // conj(Ey) + j conj(Oy)
                    pi[q] = oyr - eyi;
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float tr = __shfl_sync(kFullMask, pr[q], partner);
                    const float ti = __shfl_sync(kFullMask, pi[q], partner);
                    // lane 0: register 31-q holds bin 32 (31-q) = 1024 - 32 (q+1), the partner of its own bin q+1;
                    // register 16 (q = 15) is bin 512 itself
                    const float r0 = (q < 15) ? pr[q + 1] : zr[16];
                    const float i0 = (q < 15) ? pi[q + 1] : zi[16];
                    if (q < 15) { zr[31 - q] = lane0 ? r0 : tr; zi[31 - q] = lane0 ? i0 : ti; }
                    else { zr[16] = lane0 ? r0 : tr; zi[16] = lane0 ? i0 : ti; }
                }
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    DDSP_RE(X, brev5(q)) = zi[q];      // swapped -> inverse
                    DDSP_IM(X, brev5(q)) = zr[q];
                }
            } else {
                // y[2n] = Re z'[n] = X.im, y[2n+1] = Im z'[n] = X.re, n = lane + 32 q; overlap-add (core.py:233-235)
                // (rs is a multiple of 512, so the ring index wraps only between 512-sample blocks)
#pragma unroll
                for (int blk = 0; blk < 4; ++blk) {
                    float2* base = reinterpret_cast<float2*>(ring + ((rs + blk * kHop) & (kLtvRing - 1)) + 2 * lane);
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) {
                        const int q = 8 * blk + qq;
                        float2 v = make_float2(DDSP_IM(X, q), DDSP_RE(X, q));
                        if (blk < 3) { const float2 o = base[32 * qq]; v.x += o.x; v.y += o.y; }
                        base[32 * qq] = v;
                    }
                }
                __syncwarp();
                // retire the first 512 samples of the window: output index t = 512(m-1) - D + j  (core.py:238,177-182)
                const int m = CV_M, m_begin = CV_MBEGIN;
                const bool complete = (m - 3 >= m_begin) || (m_begin == 0);
                float* ob = P.out + (int64_t)CV_B * T;
                const int64_t tb = (int64_t)(m - 1) * kHop - D;       // frame m starts at input sample 512 (m - 1)
                const float* blk0 = ring + rs + lane;
                if (complete && tb >= 0 && tb + kHop <= T) {       // the common case: a whole finished hop inside the clip
                    float* dst = ob + tb + lane;
#pragma unroll
                    for (int r = 0; r < 16; ++r) dst[32 * r] = blk0[32 * r];
                } else {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const int j = lane + 32 * r;
                        const int64_t t = tb + j;
                        if (t >= 0 && t < T) {
                            const float v = blk0[32 * r];
                            if (complete) ob[t] = v; else atomicAdd(ob + t, v);
                        }
                    }
                }
                __syncwarp();
                rs = (rs + kHop) & (kLtvRing - 1);
            }
        }
        const int m_next = CV_M + 1;
        if (m_next >= CV_MEND) break;
        __syncwarp();
        if (lane0) CV_M = m_next;
        __syncwarp();
    }
    // flush the three remaining hops of the ring
    {
        const int m_begin = CV_MBEGIN, m_end = CV_MEND;
        const int m_last = m_end - 1;
        float* ob = P.out + (int64_t)CV_B * T;
        const int64_t tb = (int64_t)(m_last - 1) * kHop - D + kHop;       // rs already advanced past the retired hop
        for (int c = 0; c < 3; ++c) {
            const bool complete = (m_end == F + 1) && ((m_last - 2 + c >= m_begin) || (m_begin == 0));
            for (int r = 0; r < 16; ++r) {
                const int j = lane + 32 * r + kHop * c;
                const int64_t t = tb + j;
                if (t >= 0 && t < T) {
                    const float v = ring[(rs + j) & (kLtvRing - 1)];
                    if (complete) ob[t] = v; else atomicAdd(ob + t, v);
                }
            }
        }
    }
}

// Zero the output samples that receive atomic adds from two neighbouring runs: the first `hops` hops a run
// (other than the first of its clip) retires, i.e. samples [512 (m_begin - 1) - D, + 512 hops) with
// m_begin = run_begin(r).  Everything else is written with plain stores, so the rest of `out` needs no clearing.
// One CTA of 128 threads per (seam, hop).
__global__ void __launch_bounds__(128) ltv_zero_seams_kernel(float* __restrict__ out, int F, int run_len, int run_rem,
                                                             int runs_per_clip, int D, int hops, int n_seams) {
    const int seam = blockIdx.x / hops, c = blockIdx.x % hops;
    if (seam >= n_seams) return;
    const int b = seam / (runs_per_clip - 1), r = seam % (runs_per_clip - 1) + 1;
    const int64_t T = (int64_t)F * kHop;
    const int64_t t0 = (int64_t)(run_begin(r, run_len, run_rem) - 1 + c) * kHop - D;
    float* ob = out + (int64_t)b * T;
#pragma unroll
    for (int i = 0; i < kHop / 128; ++i) {
        const int64_t t = t0 + threadIdx.x + 128 * i;
        if (t >= 0 && t < T) ob[t] = 0.0f;
    }
}

// sum_out = add_in + out on the seam hops (same geometry as ltv_zero_seams_kernel), after the convolution
// kernel: there `out` is complete only once both neighbouring runs have added their part.
__global__ void __launch_bounds__(128) ltv_sum_seams_kernel(const float* __restrict__ out, const float* __restrict__ add_in,
                                                            float* __restrict__ sum_out, int F, int run_len, int run_rem,
                                                            int runs_per_clip, int D, int hops, int n_seams) {
    const int seam = blockIdx.x / hops, c = blockIdx.x % hops;
    if (seam >= n_seams) return;
    const int b = seam / (runs_per_clip - 1), r = seam % (runs_per_clip - 1) + 1;
    const int64_t T = (int64_t)F * kHop;
    const int64_t t0 = (int64_t)(run_begin(r, run_len, run_rem) - 1 + c) * kHop - D;
#pragma unroll
    for (int i = 0; i < kHop / 128; ++i) {
        const int64_t t = t0 + threadIdx.x + 128 * i;
        if (t >= 0 && t < T) sum_out[(int64_t)b * T + t] = __fadd_rn(add_in[(int64_t)b * T + t], out[(int64_t)b * T + t]);
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel 2b: the convolution for L = 510 filters (n_mag = 256: all-pass and noise filters).
// The Bartlett-windowed 1024-sample frame is the sum of its rising half (samples of hop m-1) and
// its falling half (hop m, placed 512 later); each half convolved with 510 taps spans 1021 <= 1024
// samples, and both halves meet the SAME real impulse response, so they travel as real and
// imaginary part of one complex FFT-1024: ifft(FFT(up + j*down) . H) = up*h + j (down*h) with no
// conjugate-pair exchange and no even/odd twiddles -- the spectral product is 32 plain complex
// multiplies per lane.  The frame's output spans 1536 samples (3 hops) of the ring.
// ---------------------------------------------------------------------------------------------
template <int AMODE>
__global__ void __launch_bounds__(kLtvThreads, 1) ltv_conv510_kernel(const LtvParams P) {
    const int amode = AMODE >= 0 ? AMODE : P.audio_mode;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const float4* tw4 = reinterpret_cast<const float4*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + 512 * 16 + wid * kLtvConvWarpBytes;
    float* plane = reinterpret_cast<float*>(wbase);
    float* ring = plane + kPlaneFloats;
    volatile int* ctx = reinterpret_cast<volatile int*>(ring + kLtvRing);
    {
        const float4* src = reinterpret_cast<const float4*>(P.tw_tables);
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        for (int e = threadIdx.x; e < 512; e += kLtvThreads) dst[e] = __ldg(src + e);
        __syncthreads();
    }
    {
        const int64_t slot = (int64_t)wid * gridDim.x + blockIdx.x;     // long runs evenly over the schedulers (slot_to_run)
        if (slot >= (int64_t)P.B * P.runs_per_clip) return;
        int b0, r;
        slot_to_run(slot, P.B, P.runs_per_clip, P.run_rem, b0, r);
        const int mb = run_begin(r, P.run_len, P.run_rem);
        if (lane == 0) {
            ctx[0] = b0; ctx[1] = mb; ctx[2] = mb + P.run_len + (r < P.run_rem ? 1 : 0);
            const uint64_t k64 = noise_key64(P.seed, (uint32_t)b0);
            ctx[3] = (int)((uint32_t)k64 + P.key_offset);
            ctx[5] = (int)(uint32_t)(k64 >> 32);
            bulk_mbar_init(smem_addr(ctx + 6));          // ctx[6..7]: mbarrier of the bulk copies into the transpose plane
        }
        __syncwarp();
    }
    const int F = P.F;
    const int64_t T = (int64_t)F * kHop;
    constexpr int D = 255;                               // L/2: delay compensation (core.py:177)
    // The transpose plane doubles as the landing zone of two bulk copies per frame (common.cuh), each issued by lane 0
    // where the plane falls idle -- between the two passes of a transform:
    //   audio  (amode 0 / 1): the frame's 1024 input samples, started inside the INVERSE transform of the previous frame
    //          (before the loop for the run's first frame), consumed when the frame is built;
    //   taps   : bins 0..512 of the frame's tap spectrum (4104 -> 4112 B), started inside the forward transform,
    //          consumed by the spectral product.
    // Barrier phase parity: with audio copies the uses alternate audio (0), taps (1); without, taps alone (frame & 1).
    const bool stage_audio = amode != 2;
    // frame m covers samples [512 (m-1), 512 (m+1)) of the clip; hop m-1 is absent for m = 0, hop m for m = F.  The
    // window is laid out frame-relative (sample i of the frame at float `skew + i`), `skew` = the 16-byte misalignment
    // of the audio pointer in floats (0 or 2: the ABI asks for 8-byte alignment).
    auto start_audio = [&](int mm) {
        const bool hA = mm >= 1, hB = mm < F;
        const float* src = P.audio + (int64_t)CV_B * T + (int64_t)(hA ? mm - 1 : 0) * kHop;
        const uint32_t bytes = (uint32_t)((hA ? 2048 : 0) + (hB ? 2048 : 0) + (CV_SKEW ? 16 : 0));
        const uint32_t bar = CV_BAR;
        fence_proxy_async_smem();
        bulk_mbar_expect(bar, bytes);
        bulk_copy_g2s(smem_addr(plane) + (hA ? 0 : 2048), reinterpret_cast<const void*>(reinterpret_cast<uint64_t>(src) & ~15ull),
                      bytes, bar);
    };

    for (int i = lane; i < kLtvRing; i += 32) ring[i] = 0.0f;
    __syncwarp();
    if (P.run_len >= 4) __nanosleep((unsigned)(wid >> 2) * LTV_STAGGER_NS);   // de-phase the warps of a scheduler (see combsubfast.cuh)
    int rs = 0;
    if (stage_audio && lane == 0) start_audio(CV_MBEGIN);

    Pts32 X;
    for (int m = CV_MBEGIN; m < CV_MEND; ++m) {
        const int64_t t0 = (int64_t)(m - 1) * kHop;
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == 0) {
                // z[n] = up[n] + j down[n], n = 32 n1 + lane < 512:
                //   up[n] = x[t0 + n] * n/512, down[n] = x[t0 + 512 + n] * (512 - n)/512   (core.py:218-222)
                const bool vA = m >= 1, vB = m < F;
                const float* src = plane + CV_SKEW + lane;              // the staged frame (see start_audio)
                if (stage_audio) bulk_mbar_wait(CV_BAR, 0u);
                const uint32_t key = (uint32_t)ctx[3], key2 = (uint32_t)ctx[5];
                uint32_t stA = noise_seed(key, key2, (uint32_t)(m - 1), (uint32_t)lane);
                uint32_t stB = noise_seed(key, key2, (uint32_t)m, (uint32_t)lane);
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    float v0 = 0.0f, v1 = 0.0f;
                    if (n1 < 16) {
                        const int n = 32 * n1 + lane;
                        if (amode == 2) {
                            // lane l draws the samples 32 i + l of a hop, i = 0..15, from its (hop, lane) stream
                            stA = noise_next(stA); v0 = noise_f31(stA) * 4.6566128730773926e-10f;
                            stB = noise_next(stB); v1 = noise_f31(stB) * 4.6566128730773926e-10f;
                            v0 = vA ? v0 : 0.0f;
                            v1 = vB ? v1 : 0.0f;
                        } else {
                            if (vA) v0 = src[32 * n1];
                            if (vB) v1 = src[kHop + 32 * n1];
                            if (amode == 1) { v0 = vA ? fmaf(2.0f, v0, -1.0f) : 0.0f; v1 = vB ? fmaf(2.0f, v1, -1.0f) : 0.0f; }
                        }
                        v0 *= (float)n * (1.0f / 512.0f);
                        v1 *= (float)(512 - n) * (1.0f / 512.0f);
                    }
                    DDSP_RE(X, brev5(n1)) = v0;
                    DDSP_IM(X, brev5(n1)) = v1;
                }
                if (stage_audio) __syncwarp();          // every lane has read its samples before the transform reuses the plane
            }

            warp_fft1024(X, plane, tw4, lane, [&] {
                if (lane != 0) return;
                if (phase == 0) {                       // last IR repeated (core.py:228)
                    const uint32_t bar = CV_BAR;
                    fence_proxy_async_smem();
                    bulk_mbar_expect(bar, 4112u);
                    bulk_copy_g2s(smem_addr(plane), P.spec + ((int64_t)CV_B * F + min(m, F - 1)) * kLtvSpecFloat2, 4112u, bar);
                } else if (stage_audio && m + 1 < CV_MEND) {
                    start_audio(m + 1);
                }
            });

            if (phase == 0) {
                // last IR repeated (core.py:228).  Only bins 0..512 of the Hermitian tap spectrum are stored:
                // bin k = lane + 32 q >= 512 is read as conj(H[1024 - k]) (for k = 512 that is H[512] itself, real)
                const float2* zlo = reinterpret_cast<const float2*>(plane) + lane;
                const float2* zhi = reinterpret_cast<const float2*>(plane) + (1024 - lane);
                bulk_mbar_wait(CV_BAR, stage_audio ? 1u : (uint32_t)((m - CV_MBEGIN) & 1));
                float yr[32], yi[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const float2 h = (q < 16) ? zlo[32 * q] : zhi[-32 * q];
                    const float ar = DDSP_RE(X, q), ai = DDSP_IM(X, q);
                    if (q < 16) {
                        yr[q] = ar * h.x - ai * h.y;
                        yi[q] = ar * h.y + ai * h.x;
                    } else {                   // conj(h)
                        yr[q] = ar * h.x + ai * h.y;
                        yi[q] = ai * h.x - ar * h.y;
                    }
                }
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    DDSP_RE(X, brev5(q)) = yi[q];      // swapped -> inverse
                    DDSP_IM(X, brev5(q)) = yr[q];
                }
            } else {
                // after the swapped FFT: X.im = up*h, X.re = down*h, sample n = lane + 32 q.
                // ring[rs + n] += up*h;  ring[rs + 512 + n] (+)= down*h -- its last 512 samples open a fresh hop
                // (rs is a multiple of 512, so the ring index wraps only between 512-sample blocks)
                float* blk0 = ring + rs + lane;
                float* blk1 = ring + ((rs + kHop) & (kLtvRing - 1)) + lane;
                float* blk2 = ring + ((rs + 2 * kHop) & (kLtvRing - 1)) + lane;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    blk0[32 * q] += DDSP_IM(X, q);                                           // up*h, first 512
                    blk1[32 * q] = (blk1[32 * q] + DDSP_IM(X, q + 16)) + DDSP_RE(X, q);      // up*h second, down*h first
                    blk2[32 * q] = DDSP_RE(X, q + 16);                                       // down*h second: fresh hop
                }
                __syncwarp();
                // retire the first 512 samples of the window: output index t = 512(m-1) - D + j  (core.py:238,177-182)
                const int m_begin = CV_MBEGIN;
                const bool complete = (m - 2 >= m_begin) || (m_begin == 0);
                float* ob = P.out + (int64_t)CV_B * T;
                const int64_t tb = t0 - D;
                // finished samples also go out as sum_out = add_in + out when the caller asked for it (the models'
                // signal = harmonic + noise); seam hops are summed by ltv_sum_seams_kernel after both runs added theirs
                const float* ai = P.sum_out ? P.add_in + (int64_t)CV_B * T : nullptr;
                float* so = P.sum_out ? P.sum_out + (int64_t)CV_B * T : nullptr;
                if (complete && tb >= 0 && tb + kHop <= T) {       // the common case: a whole finished hop inside the clip
                    float* dst = ob + tb + lane;
                    if (so) {
                        float hv[16];
#pragma unroll
                        for (int r = 0; r < 16; ++r) hv[r] = __ldg(ai + tb + lane + 32 * r);
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            const float v = blk0[32 * r];
                            dst[32 * r] = v;
                            so[tb + lane + 32 * r] = __fadd_rn(hv[r], v);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < 16; ++r) dst[32 * r] = blk0[32 * r];
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const int j = lane + 32 * r;
                        const int64_t t = tb + j;
                        if (t >= 0 && t < T) {
                            const float v = blk0[32 * r];
                            if (complete) { ob[t] = v; if (so) so[t] = __fadd_rn(__ldg(ai + t), v); }
                            else atomicAdd(ob + t, v);
                        }
                    }
                }
                __syncwarp();
                rs = (rs + kHop) & (kLtvRing - 1);
            }
        }
    }
    // flush the two remaining hops of the ring
    {
        const int m_begin = CV_MBEGIN, m_end = CV_MEND;
        const int m_last = m_end - 1;
        float* ob = P.out + (int64_t)CV_B * T;
        const int64_t tb = (int64_t)(m_last - 1) * kHop - D + kHop;       // rs already advanced past the retired hop
        for (int c = 0; c < 2; ++c) {
            const bool complete = (m_end == F + 1) && ((m_last - 1 + c >= m_begin) || (m_begin == 0));
            for (int r = 0; r < 16; ++r) {
                const int j = lane + 32 * r + kHop * c;
                const int64_t t = tb + j;
                if (t >= 0 && t < T) {
                    const float v = ring[(rs + j) & (kLtvRing - 1)];
                    if (complete) {
                        ob[t] = v;
                        if (P.sum_out) P.sum_out[(int64_t)CV_B * T + t] = __fadd_rn(__ldg(P.add_in + (int64_t)CV_B * T + t), v);
                    } else atomicAdd(ob + t, v);
                }
            }
        }
    }
}

// Chirp tables for one L: c[m] = exp(i pi m^2 / L) (m < 512) and the 1024-point spectrum of the
// wrapped conjugate chirp d[m] = exp(-i pi m^2 / L), m in [-(K-1), n_out-1].  Evaluated in double
// with exact integer phase reduction (one-time setup; O(N^2) direct DFT on purpose).
__global__ void chirp_tables_kernel(float2* __restrict__ c_out, float2* __restrict__ d_out, int L, int K, int n_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 1024) return;
    if (t < 512) {
        const long long r = ((long long)t * t) % (2LL * L);
        double s, c;
        sincospi((double)r / (double)L, &s, &c);
        c_out[t] = make_float2((float)c, (float)s);
    }
    double accr = 0.0, acci = 0.0;
    for (int m = -(K - 1); m <= n_out - 1; ++m) {
        const long long r = ((long long)m * m) % (2LL * L);
        double ds, dc;
        sincospi((double)r / (double)L, &ds, &dc);          // d = dc - i ds
        const int idx = (m + 1024) & 1023;
        const int e = (int)(((long long)idx * t) & 1023);   // exp(-2 pi i idx t / 1024)
        double es, ec;
        sincospi((double)e / 512.0, &es, &ec);              // = ec - i es
        accr += dc * ec - ds * es;                          // (dc - i ds)(ec - i es)
        acci += -(dc * es + ds * ec);
    }
    d_out[t] = make_float2((float)accr, (float)acci);
}

}  // namespace ddsp
