// frontend.cuh -- the steps immediately downstream of the synthesizer:
//   mel_spectrogram_kernel   nsf_hifigan/nvSTFT.py:65-116 (get_mel, keyshift 0): reflect padding, Hann-windowed STFT
//                            (n_fft = win = 2048), magnitude, mel basis, log clamp -- one kernel, one warp per frame
//   sinc_resample_kernel     torchaudio Resample (enhancer.py:47,69; gui.py:398-401): polyphase windowed-sinc FIR
//   interp_frames_kernel     enhancer.py:57-63: np.interp of f0 onto the enhancer's frame grid, in double
//   sola_splice_kernel       gui.py:408-426: normalised cross-correlation search, crossfade, new tail
#pragma once
#include "ltvfir.cuh"

namespace ddsp {

constexpr int kMelFft = 2048;                       // n_fft = win_size of the 44.1 kHz NSF-HiFiGAN front-end
constexpr int kMelBins = kMelFft / 2 + 1;
constexpr int kMelWarps = 8;                         // 256 threads: room for 64 data registers + the precise sincospi
constexpr int kMelWarpFloats = kPlaneFloats + 1028;  // transpose plane + magnitudes (1025, padded)
constexpr int kMelSmemBytes = 512 * 16 + kMelWarps * kMelWarpFloats * 4;

struct MelParams {
    const float* audio;          // (B, T)
    int B, T, n_frames, hop, pad_left, reflect;
    const float* mel_basis;      // (n_mels, 1025)
    const int* band_start;       // first / one-past-last non-zero bin of every mel filter
    const int* band_end;
    int n_mels;
    float clip_val;
    const float* tw_tables;
    float* out;                  // (B, n_mels, n_frames)
};

__device__ __forceinline__ float mel_sample(const MelParams& P, const float* row, int i) {
    // torch.nn.functional.pad(mode='reflect') (nvSTFT.py:97-103); 'constant' when the clip is shorter than the padding
    if (i < 0) i = P.reflect ? -i : -1;
    else if (i >= P.T) i = P.reflect ? 2 * (P.T - 1) - i : -1;
    return (i >= 0 && i < P.T) ? __ldg(row + i) : 0.0f;
}

__global__ void __launch_bounds__(kMelWarps * 32, 1) mel_spectrogram_kernel(const MelParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const float4* tw4 = reinterpret_cast<const float4*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* plane = reinterpret_cast<float*>(smem_raw + 512 * 16) + wid * kMelWarpFloats;
    float* mag = plane + kPlaneFloats;
    {
        const float4* src = reinterpret_cast<const float4*>(P.tw_tables);
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        for (int e = threadIdx.x; e < 512; e += kMelWarps * 32) dst[e] = __ldg(src + e);
        __syncthreads();
    }
    const int partner = (32 - lane) & 31;
    const bool lane0 = lane == 0;
    const int64_t n_total = (int64_t)P.B * P.n_frames;
    Pts32 X;
    for (int64_t fr = (int64_t)blockIdx.x * kMelWarps + wid; fr < n_total; fr += (int64_t)gridDim.x * kMelWarps) {
        const int b = (int)(fr / P.n_frames), t = (int)(fr % P.n_frames);
        const float* row = P.audio + (int64_t)b * P.T;
        const int base = t * P.hop - P.pad_left;
        // real FFT-2048 as the complex FFT-1024 of z[n] = a[2n] + j a[2n+1], a = hann * frame (nvSTFT.py:105-106)
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int n = 32 * n1 + lane;
            const float w0 = 0.5f - 0.5f * cospif((float)(2 * n) * (1.0f / 1024.0f));          // torch.hann_window(2048), periodic
            const float w1 = 0.5f - 0.5f * cospif((float)(2 * n + 1) * (1.0f / 1024.0f));
            DDSP_RE(X, brev5(n1)) = w0 * mel_sample(P, row, base + 2 * n);
            DDSP_IM(X, brev5(n1)) = w1 * mel_sample(P, row, base + 2 * n + 1);
        }
        warp_fft1024(X, plane, tw4, lane);
        // X[k] = E[k] + W2048^k O[k], E = (Z[k] + conj Z[1024-k]) / 2, O = (Z[k] - conj Z[1024-k]) / (2j); k = lane + 32 q
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            float cr, ci;
            LTV_PARTNER(X, q, cr, ci);
            const float ar = DDSP_RE(X, q), ai = DDSP_IM(X, q);
            const float er = 0.5f * (ar + cr), ei = 0.5f * (ai - ci), orr = 0.5f * (ai + ci), oi = 0.5f * (cr - ar);
            float s, c;
            sincospif((float)(lane + 32 * q) * (1.0f / 1024.0f), &s, &c);                        // W = c - j s
            const float xr = er + (c * orr + s * oi), xi = ei + (c * oi - s * orr);
            mag[lane + 32 * q] = sqrtf(fmaf(xr, xr, fmaf(xi, xi, 1e-9f)));                       // nvSTFT.py:108
            if (q == 0 && lane0) {                                                               // bin 1024 = E[0] - O[0] (real)
                const float x1024 = er - orr;
                mag[1024] = sqrtf(fmaf(x1024, x1024, 1e-9f));
            }
        }
        __syncwarp();
        // mel projection over the non-zero range of every (triangular) filter, then log clamp (nvSTFT.py:117-119)
        for (int m = lane; m < P.n_mels; m += 32) {
            const int s0 = __ldg(P.band_start + m), e0 = __ldg(P.band_end + m);
            const float* wrow = P.mel_basis + (int64_t)m * kMelBins;
            float acc = 0.0f;
            for (int k = s0; k < e0; ++k) acc = fmaf(__ldg(wrow + k), mag[k], acc);
            P.out[((int64_t)b * P.n_mels + m) * P.n_frames + t] = logf(fmaxf(acc, P.clip_val));
        }
        __syncwarp();
    }
}

// y[b][i * new + j] = sum_k kernel[j][k] * xpad[b][i * orig + k], xpad = x shifted by `width` zeros
// (torchaudio _apply_sinc_resample_kernel).  `kernel_t` is the table transposed to (K, new) so that the threads of a
// warp (consecutive output phases j of one input block i) read consecutive words.  A CTA stages the input span of
// kResTileI consecutive input blocks in shared memory.
constexpr int kResTileI = 8;

__global__ void __launch_bounds__(256) sinc_resample_kernel(const float* __restrict__ x, int T, const float* __restrict__ kernel_t,
                                                            int orig, int nw, int width, int K, float* __restrict__ y, int T_out) {
    extern __shared__ float xs[];
    const int b = blockIdx.y, i0 = blockIdx.x * kResTileI;
    const int span = (kResTileI - 1) * orig + K;
    const float* xb = x + (int64_t)b * T;
    for (int e = threadIdx.x; e < span; e += 256) {
        const int src = i0 * orig + e - width;
        xs[e] = (src >= 0 && src < T) ? __ldg(xb + src) : 0.0f;
    }
    __syncthreads();
    float* yb = y + (int64_t)b * T_out;
    for (int o = threadIdx.x; o < kResTileI * nw; o += 256) {
        const int ii = o / nw, j = o - ii * nw;
        const int64_t idx = (int64_t)(i0 + ii) * nw + j;
        if (idx >= T_out) continue;
        const float* xw = xs + ii * orig;
        float acc = 0.0f;
        for (int k = 0; k < K; ++k) acc = fmaf(__ldg(kernel_t + (int64_t)k * nw + j), xw[k], acc);
        yb[idx] = acc;
    }
}

// np.interp(time_frame, time_org, f0 * scale, left = first, right = last) in double (enhancer.py:57-63):
//   time_org[k] = (hop_over_sr * k) / real_factor,  time_frame[i] = dt_out * i
__global__ void interp_frames_kernel(const float* __restrict__ f0, int64_t fB, int64_t fN, int B, int n, float scale,
                                     double hop_over_sr, double real_factor, double dt_out, float* __restrict__ out, int n_out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)B * n_out) return;
    const int b = (int)(e / n_out), i = (int)(e % n_out);
    const float* row = f0 + (int64_t)b * fB;
    auto fp = [&](int k) { return (double)(__ldg(row + (int64_t)k * fN) * scale); };     // f0_np *= real_factor in float32
    auto xp = [&](int k) { return (hop_over_sr * (double)k) / real_factor; };
    const double t = dt_out * (double)i;
    double v;
    if (t <= xp(0)) v = fp(0);
    else if (t >= xp(n - 1)) v = fp(n - 1);
    else {
        int k = (int)floor(t * real_factor / hop_over_sr);
        k = min(max(k, 0), n - 2);
        while (k > 0 && xp(k) > t) --k;
        while (k < n - 2 && xp(k + 1) <= t) ++k;
        const double slope = (fp(k + 1) - fp(k)) / (xp(k + 1) - xp(k));
        v = slope * (t - xp(k)) + fp(k);
    }
    out[(int64_t)b * n_out + i] = (float)v;
}

// gui.py:408-426 (without the phase vocoder).  One CTA.
//   cor[d] = sum_i x[d+i] buf[i] / sqrt(sum_i x[d+i]^2 + 1e-8), d = 0..S;  shift = first argmax
//   out[i]  = x[shift+i] (* fade_in[i] + buf[i] * fade_out[i] for i < C), i < block;  buf <- x[shift+block .. +C)
constexpr int kSolaThreads = 512;

__global__ void __launch_bounds__(kSolaThreads) sola_splice_kernel(const float* __restrict__ x, float* __restrict__ sola_buffer,
                                                                   const float* __restrict__ fade_in, const float* __restrict__ fade_out,
                                                                   int block, int C, int S, float* __restrict__ out, int* __restrict__ shift_out) {
    extern __shared__ float sbuf[];                       // [C] saved tail, then the reduction scratch
    float* red_v = sbuf + C;
    int* red_i = reinterpret_cast<int*>(red_v + kSolaThreads);
    for (int i = threadIdx.x; i < C; i += kSolaThreads) sbuf[i] = sola_buffer[i];
    __syncthreads();
    float best = -INFINITY;
    int best_d = 0x7fffffff;
    for (int d = threadIdx.x; d <= S; d += kSolaThreads) {
        float nom = 0.0f, den = 0.0f;
        const float* xd = x + d;
        for (int i = 0; i < C; ++i) {
            const float v = __ldg(xd + i);
            nom = fmaf(v, sbuf[i], nom);
            den = fmaf(v, v, den);
        }
        const float c = nom / sqrtf(den + 1e-8f);
        if (c > best) { best = c; best_d = d; }          // d increases per thread: the first maximum is kept
    }
    red_v[threadIdx.x] = best;
    red_i[threadIdx.x] = best_d;
    __syncthreads();
    for (int s = kSolaThreads / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const float v2 = red_v[threadIdx.x + s];
            const int i2 = red_i[threadIdx.x + s];
            if (v2 > red_v[threadIdx.x] || (v2 == red_v[threadIdx.x] && i2 < red_i[threadIdx.x])) {
                red_v[threadIdx.x] = v2;
                red_i[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    int shift = red_i[0];
    if (shift > S || shift < 0) shift = 0;                // all-NaN input: torch.argmax would return 0 as well
    if (threadIdx.x == 0) *shift_out = shift;
    const float* xs = x + shift;
    for (int i = threadIdx.x; i < block; i += kSolaThreads) {
        float v = xs[i];
        if (i < C) v = __fadd_rn(__fmul_rn(v, fade_in[i]), __fmul_rn(sbuf[i], fade_out[i]));
        out[i] = v;
    }
    for (int i = threadIdx.x; i < C; i += kSolaThreads) sola_buffer[i] = xs[block + i];
}

}  // namespace ddsp
