// combsubfast.cuh -- stage B of CombSubFast.forward (ddsp/vocoder.py:455-492) as ONE kernel:
// excitation (combtooth sinc pulse train + uniform noise) -> sqrt-Hann framing -> 1024-point
// FFT -> per-bin filters exp(hm + j*pi*hp), exp(nm)/128 -> inverse FFT -> window -> overlap-add.
// Nothing but the control rows, f0 and the phase prefix is read from HBM and nothing but the
// finished signal is written.
//
// Work decomposition:
//   * a warp owns a run of consecutive frame PAIRS (2p, 2p+1) of one clip;
//   * frame m covers hops m-1 and m of the zero-padded excitation; each excitation hop is
//     generated once into a 2-slot shared-memory ring;
//   * per frame one complex FFT-1024 transforms comb (real part) and noise (imag part) together;
//     the two real spectra are separated with the conjugate-symmetric partner bin, which lives in
//     lane 32-l and is fetched with warp shuffles;
//   * the three filter rows of a frame (3 x 513 floats) travel global -> shared by bulk async copies that one lane
//     issues between the two passes of the frame's FFT, into the then idle transpose plane and the dead ring slot;
//   * per PAIR one complex inverse FFT-1024 returns both real output frames (V = Y_m + j*Y_{m+1});
//   * the hop shared by the two frames of a pair is finished in registers; the hop shared with
//     the next pair is carried in registers; the hop shared with the neighbouring RUN gets one
//     atomic add from each side onto pre-zeroed memory (order-independent: two addends).
#pragma once
#include "fft32.cuh"
#include "phase.cuh"

namespace ddsp {

#ifndef CSF_WARPS
#define CSF_WARPS 16
#endif
constexpr int kCsfWarps = CSF_WARPS;                // warps per CTA (one CTA per SM)
constexpr int kCsfThreads = kCsfWarps * 32;
constexpr int kRingSlot = kHop + kHop / 32;         // 528 floats: one pad word per 32 samples
constexpr int kStashFloat2 = 17 * 32;               // Y_m stash: 16 bins/lane (+ bin 512 on lane 0)
constexpr int kCsfCtxInts = 16;                     // cold per-warp scalars parked in shared memory (+ staged hop operands)
// Experiment switches; the defaults are the shipped configuration.  Measured and dropped (profiles/r02_csf_variants_*.txt,
// git history): filter bins / hop operands staged by per-lane cp.async (+5 % / +3 % time), L1 prefetch of the first
// filter lines (+1 %), the sinc sign flip as a shift-add instead of shift + xor (no change), control values of
// the first 1..6 filter bins or the next hop's operands loaded BEFORE the FFT and held in registers across it (+3 % .. +19 %
// even without spills: the FFT needs every temporary register it can get, profiles/r02_csf_variants_e_early_loads.txt);
// per-lane LDG of the filter rows, software-pipelined 6 bins ahead behind an L2 prefetch (the round-1 form, 178.4 us vs
// 174.3 us for the bulk copies below; an extra L2 prefetch in front of the bulk copies costs 4..7 us,
// profiles/r02_csf_variants_f_bulk_rows.txt, _g_bulk_prefetch.txt).
#ifndef CSF_RED_OLA
#define CSF_RED_OLA 1        // hop shared with the previous pair finished by RED.ADD instead of load + add + store
#endif
#ifndef CSF_ZERO_FAST
#define CSF_ZERO_FAST 1      // unvoiced zeroing (vocoder.py:460) only on hops that touch a non-positive f0 frame (171.2 -> 169.3 us)
#endif
#ifndef CSF_INT_PHASE
#define CSF_INT_PHASE 1      // intra-lane phase in 32-bit fixed point on top of an fp64 lane base
#endif
#ifndef CSF_DBG_SKIP
#define CSF_DBG_SKIP 0       // timing experiments only (wrong output): 1 no FFT, 2 no excitation, 4 no filter arithmetic, 8 no framing, 16 no output stores
#endif
#ifndef CSF_BALANCE
#define CSF_BALANCE 1        // long runs spread evenly over the schedulers (see slot_to_run)
#endif
constexpr int kCsfWarpBytes = kPlaneFloats * 4 + 2 * kRingSlot * 4 + kStashFloat2 * 8 + kCsfCtxInts * 4;
constexpr int kCsfSmemBytes = kTableBytes + kCsfWarps * kCsfWarpBytes;

struct CsfParams {
    const float* hm; const float* hp; const float* nm;   // (B,F,513) views, strides (cB,cF,1)
    int64_t cB, cF;
    const float* f0_frames; int64_t fB, fF;              // (B,F)
    const double* prefix;                                 // (B,F) from stage A (initial phase included)
    const float* noise_u;                                 // (B,T) or null
    const float* window;                                  // (1024,) module buffer or null
    const float* tables;                                  // device tables (twiddles + exact window)
    float* signal;                                        // (B,T)
    uint64_t seed;
    const uint64_t* seed_device;                          // optional: added to `seed` when the kernel starts (CUDA graphs)
    uint32_t key_offset;                                  // streaming: noise key shift of hop_offset hops (0 otherwise)
    int B, F;
    // run partition: run r of a clip owns pairs [r*run_len + min(r,run_rem), ... + run_len + (r < run_rem));
    // all runs differ by at most one pair and together fill the chip's warp slots (see csf_partition)
    int pairs_per_clip, run_len, run_rem, runs_per_clip;
    double inv_sr; float sr;
};

// Operands of one excitation hop, loaded a step ahead of their use so the DRAM latency is covered.
struct HopIn { float x0, x1; double base; };

__device__ __forceinline__ HopIn csf_load_hop(const CsfParams& P, int b, int h) {
    HopIn in;
    const int hc = min(max(h, 0), P.F - 1);
    const float* row = P.f0_frames + (int64_t)b * P.fB;
    in.x0 = __ldg(row + (int64_t)hc * P.fF);
    in.x1 = __ldg(row + (int64_t)min(hc + 1, P.F - 1) * P.fF);
    in.base = __ldg(P.prefix + (int64_t)b * P.F + hc);
    return in;
}

// Generate excitation hop h of clip b into a ring slot (zeros outside [0,F)).
// Lane l owns samples 16l..16l+15; the slot is padded by one word per 32 samples so that both
// this write pattern and the FFT-order read (stride 32 across registers, lanes consecutive) are
// bank-conflict free.
__device__ __forceinline__ void csf_gen_hop(const CsfParams& P, int h, const HopIn& in, float* __restrict__ slot,
                                            int lane) {
    float* dst = slot + 16 * lane + (lane >> 1);
    if (h < 0 || h >= P.F) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[i] = 0.0f;
        return;
    }
    float2 f2[8], rot2[8];
    // (an opaque copy of the lane index: ptxas otherwise hoists the lane's interpolation weights out of the step loop and
    // spills them to local memory, whose reloads miss the tiny L1 left beside 224 KB of shared memory)
    asm volatile("" : "+r"(lane));
#if CSF_INT_PHASE
    hop_rotation_q32(in.x0, in.x1, in.base, P.inv_sr, lane, f2, rot2);
    const float sr_scale = P.sr * 2.3283064365386963e-10f;      // rot2 holds rot * 2^32
#else
    hop_rotation2(in.x0, in.x1, in.base, P.inv_sr, lane, f2, rot2);
    const float sr_scale = P.sr;
#endif
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // vocoder.py:459  sinc(sr * rot / (f0 + 1e-3)), two samples per packed register pair
        const float2 den = add2(f2[j], bc2(1e-3f));
        const float2 xs = fma2(mul2(bc2(sr_scale), rot2[j]), make_float2(rcp_approx(den.x), rcp_approx(den.y)), bc2(1e-30f));
        const float2 c = sinc2_xs(xs);
#if CSF_ZERO_FAST
        dst[2 * j] = c.x;
        dst[2 * j + 1] = c.y;
#else
        dst[2 * j] = (f2[j].x <= 0.0f) ? 0.0f : c.x;       // vocoder.py:460
        dst[2 * j + 1] = (f2[j].y <= 0.0f) ? 0.0f : c.y;
#endif
    }
#if CSF_ZERO_FAST
    // vocoder.py:460 combtooth[f0 <= 0] = 0.  An interpolated sample can only be non-positive when one of the hop's two
    // frame values is (warp-uniform test): voiced hops skip the per-sample compare + select.
    if (in.x0 <= 0.0f || in.x1 <= 0.0f) {
        const float lam0 = (float)(16 * lane) * (1.0f / kHop);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (lerp_torch(in.x0, in.x1, lam0 + (float)i * (1.0f / kHop)) <= 0.0f) dst[i] = 0.0f;
    }
#endif
}

// ---- the three filter rows of one frame by bulk async copy (helpers in common.cuh) --------------------------------
// A row is 513 floats = 2052 B at a 4-byte-aligned address; cp.async.bulk wants 16-byte-aligned addresses and sizes, so
// the copy covers the 16-byte-aligned window around the row: 2064 B from (address & ~15).  The <= 12 B read in front of /
// behind the row lie in the same 16-byte granule as valid bytes of the row (never on another page).
constexpr int kRowWindowBytes = 2064;
// row -> shared window; the row's first float sits at window + csf_row_skew(row)
__device__ __forceinline__ void csf_bulk_row(uint32_t dst, const float* row, uint32_t bar) {
    bulk_copy_g2s(dst, reinterpret_cast<const void*>(reinterpret_cast<uint64_t>(row) & ~15ull), kRowWindowBytes, bar);
}
__device__ __forceinline__ int csf_row_skew(const float* row) { return (int)((reinterpret_cast<uint64_t>(row) & 15ull) >> 2); }

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Zero the seam hops (first output hop of every run that does not start a clip).
__device__ __forceinline__ int csf_run_begin(int r, int run_len, int run_rem) { return r * run_len + min(r, run_rem); }

__global__ void __launch_bounds__(128) csf_zero_seams_kernel(float* __restrict__ signal, int F, int run_len, int run_rem,
                                                             int runs_per_clip, int n_seams) {
    const int seam = blockIdx.x;
    cudaGridDependencySynchronize();          // launched with programmatic stream serialisation
    if (seam >= n_seams) return;
    const int b = seam / (runs_per_clip - 1), r = seam % (runs_per_clip - 1) + 1;
    const int hop = 2 * csf_run_begin(r, run_len, run_rem) - 1;
    if (hop >= F) return;
    float4* dst = reinterpret_cast<float4*>(signal + ((int64_t)b * F + hop) * kHop);
    dst[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Windowed frame fm of clip b in FFT input order: comb excitation (ring slots of hops fm-1, fm)
// -> real part, noise excitation (injected U or the in-kernel stream) -> imaginary part.
template <bool HAS_U>
__device__ __forceinline__ void csf_load_frame(const CsfParams& P, Pts32& X, const float* __restrict__ ring,
                                               const float* __restrict__ win, int fm, int b, uint32_t key, uint32_t key2,
                                               int lane) {
    const int F = P.F;
    const float* slotA = ring + ((fm - 1) & 1) * kRingSlot + lane;   // hop fm-1
    const float* slotB = ring + (fm & 1) * kRingSlot + lane;         // hop fm
    // hops outside [0,F) are zero padding (vocoder.py:463-464): their noise weight is 0 and
    // the (unused) noise sample is read from the start of the clip to stay in bounds
    const bool vA = (fm >= 1) && (fm - 1 < F), vB = fm < F;
    const int64_t baseA = vA ? (int64_t)(fm - 1) * kHop : 0, baseB = vB ? (int64_t)fm * kHop : 0;
    const float okA = vA ? 1.0f : 0.0f, okB = vB ? 1.0f : 0.0f;
    const float sclA = okA * 4.6566128730773926e-10f, sclB = okB * 4.6566128730773926e-10f;   // 2^-31
    const float* u_b = HAS_U ? P.noise_u + (int64_t)b * F * kHop + lane : nullptr;
    uint32_t stA = 0, stB = 0;
    if (!HAS_U) {
        stA = noise_seed(key, key2, (uint32_t)(fm - 1), (uint32_t)lane);
        stB = noise_seed(key, key2, (uint32_t)fm, (uint32_t)lane);
    }
    // samples 32*n1 + lane for n1 = 2j, 2j+1 land in the two halves of one packed register
    // (register index brev5(2j) and brev5(2j) + 16), so each window pair feeds packed multiplies
    const float2* winA = reinterpret_cast<const float2*>(win);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n1 = 2 * j;                                         // both samples lie in the same hop (n1 < 16: hop fm-1)
        const int m0 = n1 & 15, m1 = m0 + 1;                          // sample in hop = 32*m + lane, one pad word per 32
        const float2 w = winA[j * 32 + lane];
        const float* slot = (n1 < 16) ? slotA : slotB;
        const float2 c = make_float2(slot[33 * m0], slot[33 * m1]);
        float2 wz;
        if (HAS_U) {
            const float* up = u_b + (n1 < 16 ? baseA : baseB);
            const float2 u = make_float2(__ldg(up + 32 * m0), __ldg(up + 32 * m1));
            const float2 wn = mul2(w, bc2(n1 < 16 ? okA : okB));
            wz = fma2(u, add2(wn, wn), neg2(wn));                         // w * (2u - 1)   (vocoder.py:461)
        } else {
            uint32_t& st = (n1 < 16) ? stA : stB;
            st = noise_next(st);
            const float v0 = noise_f31(st);
            st = noise_next(st);
            const float v1 = noise_f31(st);
            // 2u - 1 = v * 2^-31 with v = 2^8 * the centred 24-bit draw; sclA/sclB carry the 2^-31 (or 0)
            wz = mul2(make_float2(v0, v1), mul2(w, bc2(n1 < 16 ? sclA : sclB)));
        }
        X.R[brev5(n1)] = mul2(w, c);
        X.I[brev5(n1)] = wz;
    }
}

// HAS_U: the noise excitation is read from the injected U tensor (parity mode) instead of being drawn
// in-kernel; a compile-time switch so that neither variant carries the other's predicated-off code.
template <bool HAS_U>
__global__ void __launch_bounds__(kCsfThreads, 1) combsubfast_kernel(const CsfParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const float4* tw4 = reinterpret_cast<const float4*>(smem_raw);
    float* win = reinterpret_cast<float*>(smem_raw + 512 * 16);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + kTableBytes + wid * kCsfWarpBytes;
    float* plane = reinterpret_cast<float*>(wbase);
    float* ring = plane + kPlaneFloats;
    float2* stash = reinterpret_cast<float2*>(ring + 2 * kRingSlot);

    {   // CTA-wide tables: twiddles + window (the module's `window` buffer when given)
        const float4* src = reinterpret_cast<const float4*>(P.tables);
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        for (int e = threadIdx.x; e < kTableBytes / 16; e += kCsfThreads) dst[e] = __ldg(src + e);
        __syncthreads();
        cudaGridDependencySynchronize();      // launched with programmatic stream serialisation (see ddsp_b200.cu)
        if (P.window) {
            for (int e = threadIdx.x; e < 1024; e += kCsfThreads) {
                const float w = __ldg(P.window + e);
                win[win_analysis_index(e)] = w;
                win[win_synthesis_index(e)] = w;
            }
            __syncthreads();
        }
    }

    // Per-warp scalars that are needed only a few times per step live in shared memory (volatile
    // reads) instead of registers: the FFT keeps 64 data registers live and anything else held
    // across it would be spilled to local memory, whose reloads miss the (tiny) L1 here.
    //   ctx[0] clip, [1] first pair, [2] end pair, [3] noise key (low word + hop offset), [4] step, [5] noise key (high word),
    //   ctx[6..7] mbarrier of the filter-row copies
    volatile int* ctx = reinterpret_cast<volatile int*>(stash + kStashFloat2);
    {
        // runs are dealt over the warp slots so that every scheduler gets the same mix of long and short runs
        const int64_t slot = (int64_t)wid * gridDim.x + blockIdx.x;
        if (slot >= (int64_t)P.B * P.runs_per_clip) return;
        int b0, r;
#if CSF_BALANCE
        slot_to_run(slot, P.B, P.runs_per_clip, P.run_rem, b0, r);
#else
        b0 = (int)(slot / P.runs_per_clip);
        r = (int)(slot % P.runs_per_clip);
#endif
        const int pb = csf_run_begin(r, P.run_len, P.run_rem);
        if (lane == 0) {
            ctx[0] = b0;
            ctx[1] = pb;                                        // p_begin
            ctx[2] = pb + P.run_len + (r < P.run_rem ? 1 : 0);  // p_end
            const uint64_t seed = P.seed + (P.seed_device ? __ldg(P.seed_device) : 0ull);
            const uint64_t k64 = noise_key64(seed, (uint32_t)b0);
            ctx[3] = (int)((uint32_t)k64 + P.key_offset);
            ctx[5] = (int)(uint32_t)(k64 >> 32);
            bulk_mbar_init(smem_addr(ctx + 6));                    // ctx[6..7]: the warp's mbarrier for the filter-row copies
        }
        __syncwarp();
    }
#define CTX_B ctx[0]
#define CTX_PBEGIN ctx[1]
#define CTX_PEND ctx[2]
#define CTX_KEY ((uint32_t)ctx[3])
#define CTX_KEY2 ((uint32_t)ctx[5])
    // Seams between runs: the hop shared by the last frame of run r-1 and the first frame of run r
    // receives one atomic add from each side onto zeros written by csf_zero_seams_kernel
    // (0 + a + b is the same fp32 number in either order, so the result does not depend on timing).
    const int F = P.F;
    const int partner = (32 - lane) & 31;
    const bool lane0 = lane == 0;


    Pts32 X;
    // De-phase the four warps that share a scheduler (wid, wid+4, wid+8, wid+12) by 1 us each so that
    // their FMA-heavy FFT phases do not start in lockstep (measured on the 12-pairs-per-warp headline launch:
    // 0 us 206.0, 0.5 us 203.3, 1 us 201.9, 2 us 206.2 us); skipped for short runs where it would only add latency.
#ifndef CSF_STAGGER_NS
#define CSF_STAGGER_NS 1000u
#endif
    if (P.run_len >= 4) __nanosleep((unsigned)(wid >> 2) * CSF_STAGGER_NS);

    // steps per pair: s=0 frame 2p, s=1 frame 2p+1, s=2 inverse FFT of the pair + overlap-add.
    // s=-1 (first iteration only) just generates the first hop of the run.
    // (the step counter lives in the per-warp shared-memory context too: ptxas otherwise spills it to local memory)
#define CTX_STEP ctx[4]
    int p = CTX_PBEGIN;
    if (lane == 0) CTX_STEP = -1;
    __syncwarp();
    // @section loop
#pragma unroll 1
    for (;;) {
        const int s = CTX_STEP;
        const int fm = 2 * p + s;                           // frame handled by steps 0 and 1
        if (s < 2) {
            // @section excite
            // ---- excitation hop fm (second half of frame fm; fm = 2p-1 on the priming step) ----
#if CSF_DBG_SKIP & 2
            if (fm == -12345) csf_gen_hop(P, fm, csf_load_hop(P, CTX_B, fm), ring + (fm & 1) * kRingSlot, lane);
#else
            // (f0 / prefix of consecutive hops share cache lines: after the first hop of a run these are L2 hits)
            csf_gen_hop(P, fm, csf_load_hop(P, CTX_B, fm), ring + (fm & 1) * kRingSlot, lane);
#endif
            __syncwarp();
            if (s < 0) {
                CTX_STEP = 0; __syncwarp(); continue;
            }
            // @section frame
#if CSF_DBG_SKIP & 8
            if (fm == -12345)
#endif
            csf_load_frame<HAS_U>(P, X, ring, win, fm, CTX_B, CTX_KEY, CTX_KEY2, lane);
        }
        // @section fft
        // s == 2: re/im already hold the packed spectrum of the pair (swapped for the inverse)

#if !(CSF_DBG_SKIP & 1)
        // Between the two passes of the FFT the transpose plane is free (until the next transform) and so is the ring
        // slot of hop fm-1 (until hop fm+1 is generated): lane 0 starts the copies of this frame's three filter rows into
        // them -- hm | hp into the plane, nm into the slot.  They land while the second pass runs; no register is held.
        warp_fft1024(X, plane, tw4, lane, [&] {
            if (lane0 && CTX_STEP < 2) {
                const uint32_t bar = smem_addr(ctx + 6);
                const int fr = 2 * p + CTX_STEP;
                const int64_t ro = (int64_t)CTX_B * P.cB + (int64_t)min(fr, F - 1) * P.cF;
                fence_proxy_async_smem();                  // the warp's generic accesses before the async writes
                bulk_mbar_expect(bar, 3 * kRowWindowBytes);
                csf_bulk_row(smem_addr(plane), P.hm + ro, bar);
                csf_bulk_row(smem_addr(plane) + kRowWindowBytes, P.hp + ro, bar);
                csf_bulk_row(smem_addr(ring + ((fr - 1) & 1) * kRingSlot), P.nm + ro, bar);
            }
        });
#endif

        // @section filter
        if (CTX_STEP < 2) {
            // ---- split the two real spectra, apply the filters (vocoder.py:472-481) ----------
            // last filter frame repeated (:473,476)
            const int64_t ro = (int64_t)CTX_B * P.cB + (int64_t)min(fm, F - 1) * P.cF + lane;
            const int k16 = 512 - lane;                                       // bin 512 (used by lane 0 only): every lane reads that one word
            float yr[17], yi[17];
            // the rows were copied into shared memory while the FFT ran (parity of the barrier = step: two copies per pair)
            const float* hm_r = plane + csf_row_skew(P.hm + ro - lane) + lane;
            const float* hp_r = plane + kRowWindowBytes / 4 + csf_row_skew(P.hp + ro - lane) + lane;
            const float* nm_r = ring + ((fm - 1) & 1) * kRingSlot + csf_row_skew(P.nm + ro - lane) + lane;
            bulk_mbar_wait(smem_addr(ctx + 6), (uint32_t)(CTX_STEP & 1));
            // control loads run kLook bins ahead of their use (software pipeline over the unrolled loop)
#pragma unroll
            for (int q = 0; q < 17; ++q) {
                float a, bb, c, d;
                const int off = (q < 16) ? 32 * q : k16;                         // bin 512: lane 0 (others: dummy)
                const float vhm = hm_r[off], vhp = hp_r[off], vnm = nm_r[off];
                if (q < 16) {
                    a = DDSP_RE(X, q); bb = DDSP_IM(X, q);
                    c = __shfl_sync(kFullMask, DDSP_RE(X, 31 - q), partner);
                    d = __shfl_sync(kFullMask, DDSP_IM(X, 31 - q), partner);
                    // lane 0 holds bins 32q: its partner 1024-32q sits in its own register 32-q
                    const float c0 = DDSP_RE(X, (32 - q) & 31), d0 = DDSP_IM(X, (32 - q) & 31);
                    c = lane0 ? c0 : c;
                    d = lane0 ? d0 : d;
                } else {        // bin 512: lane 0, register 16, its own partner (other lanes: harmless dummy)
                    a = DDSP_RE(X, 16); bb = DDSP_IM(X, 16); c = a; d = bb;
                }
                // packed fp32x2 with single-register broadcast operands (same roundings as the scalar form):
                //   (Cr, Ni) = (c + a, c - a),  (Nr, Ci) = (bb + d, bb - d)
                const float2 CrNi = fma2(bc2(a), make_float2(1.0f, -1.0f), bc2(c));
                const float2 NrCi = fma2(bc2(d), make_float2(1.0f, -1.0f), bc2(bb));
                // H = exp(hm + j*pi*hp) (vocoder.py:472), N = exp(nm)/128 (:475); the 1/2 of the
                // split and the 1/1024 of irfft are folded in as exact powers of two.
#if CSF_DBG_SKIP & 4
                yr[q] = a + c + vhm; yi[q] = bb + d + vhp + vnm;
                continue;
#endif
                const float g = ex2_approx(fmaf(vhm, DDSP_LOG2E_F, -11.0f));
                const float nf = ex2_approx(fmaf(vnm, DDSP_LOG2E_F, -18.0f));
                const float ang = DDSP_PI_F * vhp;
                const float2 H = mul2(bc2(g), make_float2(cos_approx(ang), sin_approx(ang)));
                yr[q] = fmaf(CrNi.x, H.x, fmaf(-NrCi.y, H.y, NrCi.x * nf));
                yi[q] = fmaf(CrNi.x, H.y, fmaf(NrCi.y, H.x, CrNi.y * nf));
            }
            // irfft ignores the imaginary part of the DC and Nyquist bins
            yi[0] = lane0 ? 0.0f : yi[0];
            yi[16] = 0.0f;
            // @section stash
            if (CTX_STEP == 0) {
#pragma unroll
                for (int q = 0; q < 16; ++q) stash[q * 32 + lane] = make_float2(yr[q], yi[q]);
                if (lane0) stash[16 * 32] = make_float2(yr[16], 0.0f);
                CTX_STEP = 1;
            } else {
                // @section pack
                // ---- V = Y_m + j*Y_{m+1}; feed (Im V, Re V) to the forward FFT = inverse FFT ---
                float xr[16], xi[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float2 y = stash[q * 32 + lane];
                    // V[k] = (y.x - yi) + j (y.y + yr);  V[N-k] = (y.x + yi) + j (yr - y.y)
                    DDSP_RE(X, brev5(q)) = y.y + yr[q];       // swapped: "real" input = Im V
                    DDSP_IM(X, brev5(q)) = y.x - yi[q];
                    xr[q] = y.x + yi[q];              // Re V[N-k]
                    xi[q] = yr[q] - y.y;              // Im V[N-k]
                }
                const float v512r = stash[16 * 32].x;  // (lane 0 only meaningful)
                const float v512i = yr[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    float tr = __shfl_sync(kFullMask, xr[q], partner);
                    float ti = __shfl_sync(kFullMask, xi[q], partner);
                    const float tr0 = (q == 15) ? v512r : xr[(q + 1) & 15];
                    const float ti0 = (q == 15) ? v512i : xi[(q + 1) & 15];
                    tr = lane0 ? tr0 : tr;
                    ti = lane0 ? ti0 : ti;
                    DDSP_RE(X, brev5(31 - q)) = ti;            // swapped
                    DDSP_IM(X, brev5(31 - q)) = tr;
                }
                CTX_STEP = 2;
            }
        } else {
            // @section ola
            // ---- window (vocoder.py:486), overlap-add (:485-487), crop (:490) -----------------
            // after the swapped FFT: DDSP_IM(X, ) = Re v = frame 2p, DDSP_RE(X, ) = Im v = frame 2p+1
            // Hop 2p-1 (shared with the previous pair) was left in the output buffer by that pair as a
            // partial sum and is completed here by one reduction per sample (same thread stored the partial:
            // program order on the same address; stored value + one addend is the same fp32 number as a
            // load-add-store); at a run seam both sides add onto zeros instead.
#if CSF_DBG_SKIP & 16
            if (X.I[3].x == 1.2345f)
#endif
            {
            const int hopA = 2 * p - 1, hopB = 2 * p, hopC = 2 * p + 1;
            const int p_begin = CTX_PBEGIN, p_end = CTX_PEND;
            const bool first = p == p_begin, last = p + 1 >= p_end;
            const bool seam_head = p_begin > 0, seam_tail = p_end < P.pairs_per_clip;
            float* out_b = P.signal + (int64_t)CTX_B * F * kHop + lane;
            float* oA = out_b + (int64_t)hopA * kHop;
            float* oB = out_b + (int64_t)hopB * kHop;
            float* oC = out_b + (int64_t)hopC * kHop;
            // window both frames of the pair in place: I[q] -> (hop A part, hop B part of frame 2p),
            // R[q] -> (hop B part, hop C part of frame 2p+1)
            const float2* winB = reinterpret_cast<const float2*>(win + 1024);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float2 w = winB[q * 32 + lane];
                X.I[q] = mul2(X.I[q], w);
                X.R[q] = mul2(X.R[q], w);
            }
#if CSF_RED_OLA
            if (!first || seam_head) {
#pragma unroll
                for (int q = 0; q < 16; ++q) atomicAdd(oA + 32 * q, X.I[q].x);
            }
#else
            if (first) {
                if (seam_head) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) atomicAdd(oA + 32 * q, X.I[q].x);
                }
            } else {
                float prev[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) prev[q] = oA[32 * q];
#pragma unroll
                for (int q = 0; q < 16; ++q) oA[32 * q] = __fadd_rn(prev[q], X.I[q].x);
            }
#endif
            if (hopB < F) {
#pragma unroll
                for (int q = 0; q < 16; ++q) oB[32 * q] = __fadd_rn(X.I[q].y, X.R[q].x);
            }
            if (hopC < F) {
                if (!last) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) oC[32 * q] = X.R[q].y;
                } else if (seam_tail) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) atomicAdd(oC + 32 * q, X.R[q].y);
                }
            }
            }
            if (++p >= CTX_PEND) break;
            CTX_STEP = 0;
        }
        __syncwarp();
    }
}

}  // namespace ddsp
