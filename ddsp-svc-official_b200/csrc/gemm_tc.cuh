// gemm_tc.cuh -- fp32-faithful Linear layer on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   C[m, n] = sum_k A[m, k] * W[n, k]  (+ bias[n]) (+ residual[m, n])          A: (M, K), W: (N, K) = nn.Linear.weight
//
// Used by the control network that feeds the synthesizer (ddsp/unit2control.py:56-62, ddsp/pcmer.py:41-63,
// :191-251: every nn.Linear / 1x1 Conv1d of PCmer and the final 256 -> 1539 projection) -- SURVEY section 8
// row (f1).  The reference computes these in fp32; a single TF32 pass (10-bit mantissa) would move the
// control rows at the 1e-3 level, so each operand is split into two TF32 terms, x = hi + lo with
// hi = x rounded to TF32 (exactly representable) and lo = x - hi (exact in fp32), and three MMAs accumulate
// hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator ("3xTF32": the dropped lo*lo term and the rounding of lo
// are ~2^-21 relative).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   cp.async.bulk.tensor (128-byte swizzle) of the raw fp32 A (128 x 32) and W (BN x 32)
//                              tiles of one k-block into the next free stage
//   warps 6..9  splitter       in shared memory: tile -> hi (in place) + lo (second buffer), then fence.proxy.async
//   warp 1      MMA issuer     one elected lane: 4 k-steps x 3 tcgen05.mma (kind::tf32, M=128, N=BN, K=8) per k-block,
//                              tcgen05.commit releases the stage / publishes the accumulator
//   warps 2..5  epilogue       tcgen05.ld of the finished accumulator (TMEM lane = output row), bias / residual,
//                              128-bit global stores; overlaps the next tile's main loop (two TMEM accumulators)
// Both operands are K-major (A row-major, W = (out, in) row-major), staged in the canonical 128-byte-swizzled
// UMMA layout that TMA writes.  L2 -> SM traffic per k-block is (128 + BN) * 128 B for 3 * 128 * BN * 32 MACs.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddsp {
namespace tc {

constexpr int kBM = 128;           // rows of A per tile = TMEM lanes
constexpr int kBK = 32;            // fp32 elements per 128-byte swizzle row = one k-block
constexpr int kThreads = 320;      // 10 warps (roles above)
constexpr int kEpiWarp0 = 2, kSplitWarp0 = 6;

template <int BN>
struct Cfg {
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128: multiple of 16 in [16, 256]");
    static constexpr int kABytes = kBM * 128;                       // one operand tile of A (hi or lo)
    static constexpr int kWBytes = BN * 128;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kWBytes;   // A_hi | A_lo | W_hi | W_lo
    static constexpr int kStages = (200 * 1024) / kStageBytes < 2 ? 2 : ((200 * 1024) / kStageBytes > 6 ? 6 : (200 * 1024) / kStageBytes);
    static constexpr int kAccCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;   // column pitch of one accumulator
    static constexpr int kTmemCols = 2 * kAccCols;                  // two accumulators (power of two >= 32)
    static constexpr int kBarBytes = 256;
    static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;   // + slack for the 1024-byte alignment
};

struct LinearParams {
    const float* bias;          // (N) or null
    const float* residual;      // (M, N) with row stride ldr, or null
    float* C;                   // (M, N) with row stride ldc
    int64_t ldr, ldc;
    int M, N, K;
    int tiles_m, tiles_n;
    int store_output;           // 0: microbenchmark mode (the epilogue reads TMEM but does not store)
    int virtual_tiles;          // > 0: microbenchmark mode, every tile maps to tile 0 of A / W (operands stay in L2)
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}

// x rounded to the nearest TF32 number (10 explicit mantissa bits; ties away from zero): adding half a TF32 ulp to
// the bit pattern carries into the exponent correctly, masking the low 13 bits then truncates.  x - tf32_hi(x) is
// exact in fp32 and at most 2^-11 |x|.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

// K-major operand tile in the canonical 128-byte-swizzled layout (8-row groups of 1024 B): the shared-memory
// matrix descriptor of cute::UMMA::SmemDescriptor -- start address >> 4 in bits [0,14), leading byte offset
// (unused for swizzled K-major, set to 1) in [16,30), stride byte offset 1024 >> 4 in [32,46), version 1 in
// [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6), a/b format TF32 = 2 at [7,10) and
// [10,13), both operands K-major (bits 15, 16 clear), N >> 3 at [17,23), M >> 4 at [24,29)
template <int BN>
__host__ __device__ constexpr uint32_t umma_idesc_tf32() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

// ---- the kernel --------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const LinearParams P) {
    using C = Cfg<BN>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* bars = smem + C::kStages * C::kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bars);                 // TMA -> splitter
    uint64_t* split_bar = full_bar + C::kStages;                            // splitter -> MMA
    uint64_t* empty_bar = split_bar + C::kStages;                           // MMA -> TMA
    uint64_t* acc_full = empty_bar + C::kStages;                            // MMA -> epilogue   [2]
    uint64_t* acc_empty = acc_full + 2;                                     // epilogue -> MMA   [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    static_assert((3 * C::kStages + 4) * 8 + 4 <= C::kBarBytes, "barrier block too small");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kb = (P.K + kBK - 1) / kBK;
    const int n_tiles = P.virtual_tiles > 0 ? P.virtual_tiles : P.tiles_m * P.tiles_n;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_w);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < C::kStages; ++s) {
                mbar_init(s32(full_bar + s), 1);
                mbar_init(s32(split_bar + s), 128);
                mbar_init(s32(empty_bar + s), 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(s32(acc_full + a), 1);
                mbar_init(s32(acc_empty + a), 128);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<C::kTmemCols>(s32(tmem_slot));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int t = P.virtual_tiles > 0 ? 0 : tile;
                const int m0 = (t / P.tiles_n) * kBM, n0 = (t % P.tiles_n) * BN;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(s32(empty_bar + stage), phase ^ 1);
                    const uint32_t st = s32(smem + stage * C::kStageBytes);
                    mbar_arrive_expect_tx(s32(full_bar + stage), C::kABytes + C::kWBytes);
                    tma_load_2d(st, &map_a, kb * kBK, m0, s32(full_bar + stage));
                    tma_load_2d(st + 2 * C::kABytes, &map_w, kb * kBK, n0, s32(full_bar + stage));
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = umma_idesc_tf32<BN>();
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(s32(acc_empty + acc), acc_phase ^ 1);          // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d = tmem_base + acc * C::kAccCols;
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(s32(split_bar + stage), phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t st = s32(smem + stage * C::kStageBytes);
                    const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + C::kABytes);
                    const uint64_t w_hi = umma_desc_sw128(st + 2 * C::kABytes), w_lo = umma_desc_sw128(st + 2 * C::kABytes + C::kWBytes);
#pragma unroll
                    for (int kk = 0; kk < kBK / 8; ++kk) {
                        // advance 8 tf32 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (>> 4) address field
                        const uint64_t o = (uint64_t)(2 * kk);
                        umma_tf32(d, a_lo + o, w_hi + o, idesc, (kb | kk) != 0);      // small terms first
                        umma_tf32(d, a_hi + o, w_lo + o, idesc, 1);
                        umma_tf32(d, a_hi + o, w_hi + o, idesc, 1);
                    }
                    umma_commit(s32(empty_bar + stage));                              // stage free once these MMAs have read it
                    if (kb == n_kb - 1) umma_commit(s32(acc_full + acc));             // accumulator complete
                }
                __syncwarp();
                if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= kSplitWarp0) {
        // ===== splitter: x -> hi (in place, exactly TF32) + lo (x - hi, exact) =====
        const int t = threadIdx.x - kSplitWarp0 * 32;          // 0..127
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(s32(full_bar + stage), phase);
                float4* a_hi = reinterpret_cast<float4*>(smem + stage * C::kStageBytes);
                float4* a_lo = a_hi + C::kABytes / 16;
                float4* w_hi = a_lo + C::kABytes / 16;
                float4* w_lo = w_hi + C::kWBytes / 16;
#pragma unroll 4
                for (int i = t; i < C::kABytes / 16; i += 128) {
                    const float4 x = a_hi[i];
                    float4 h;
                    h.x = tf32_hi(x.x);
                    h.y = tf32_hi(x.y);
                    h.z = tf32_hi(x.z);
                    h.w = tf32_hi(x.w);
                    a_hi[i] = h;
                    a_lo[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
                }
#pragma unroll 4
                for (int i = t; i < C::kWBytes / 16; i += 128) {
                    const float4 x = w_hi[i];
                    float4 h;
                    h.x = tf32_hi(x.x);
                    h.y = tf32_hi(x.y);
                    h.z = tf32_hi(x.z);
                    h.w = tf32_hi(x.w);
                    w_hi[i] = h;
                    w_lo[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
                }
                fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core (async proxy)
                mbar_arrive(s32(split_bar + stage));
                if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue (warps 2..5): TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int row_in_tile = q * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int t = P.virtual_tiles > 0 ? 0 : tile;
            const int m0 = (t / P.tiles_n) * kBM, n0 = (t % P.tiles_n) * BN;
            const int row = m0 + row_in_tile;
            mbar_wait(s32(acc_full + acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::kAccCols;
            const bool row_ok = row < P.M && P.store_output;
            float* crow = P.C + (int64_t)row * P.ldc;
            const float* rrow = P.residual ? P.residual + (int64_t)row * P.ldr : nullptr;
            const bool vec_ok = ((P.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0) &&
                                (!P.residual || (((P.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.residual) & 15) == 0)));
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
                const int n = n0 + c0;
                if (row_ok && n < P.N) {
                    if (vec_ok && n + 32 <= P.N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                   __uint_as_float(r[j + 3]));
                            if (P.bias) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + n + j));
                                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                            }
                            if (rrow) {
                                const float4 e = *reinterpret_cast<const float4*>(rrow + n + j);
                                v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
                            }
                            *reinterpret_cast<float4*>(crow + n + j) = v;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (n + j < P.N) {
                                float v = __uint_as_float(r[j]);
                                if (P.bias) v += __ldg(P.bias + n + j);
                                if (rrow) v += rrow[n + j];
                                crow[n + j] = v;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(s32(acc_empty + acc));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<C::kTmemCols>(tmem_base);
}

}  // namespace tc
}  // namespace ddsp
