// gemm_attn.cuh -- the control network's GEMM-shaped work on the 5th-generation tensor cores, second generation of
// csrc/gemm_tc.cuh: one batched 3xTF32 tcgen05 GEMM with fused epilogues, and the FAVOR+ feature map as a GEMM
// epilogue.  Together they replace every nn.Linear / 1x1 Conv1d / LayerNorm / einsum of one PCmer layer
// (ddsp/pcmer.py:20-78, :124-160, :191-251) -- SURVEY section 8 row (f1).
//
//   gemm3x_kernel<BN, EPI>      C[z] = A[z] (M x K) * W[z]^T (N x K), operands K-major, fp32-faithful split
//                               accumulation (hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator)
//       EPI_PLAIN   + bias + residual, rows leave through TMA stores (coalesced, clipped at the tensor edge);
//                   optionally LayerNorm(C) as a second output while the row is still in TMEM (N <= BN)
//       EPI_QKV     merged q|k|v projection: q, k stored head-major (B,H,F,64); v stored transposed (B,H,80,Fp)
//                   next to a row of ones, so that the context GEMM also yields sum_n k'
//       EPI_OUT     attention output: columns 0..63 divided by column 64 (= q' . k_sum) + 1e-8, heads merged
//       EPI_GLU     first pointwise conv of the conformer module with GLU fused (pcmer.py:52-53): the weight rows are
//                   interleaved per column tile (128 value channels | their 128 gate channels), the epilogue writes
//                   (a + b_a) * sigmoid(g + b_g) -- half the columns, no (B,N,1024) tensor in HBM
//   favor_features_kernel<Q>    dash = x * (scale * projection)^T with the projection resident in shared memory,
//                               epilogue = the softmax-kernel feature map (pcmer.py:124-160); q' row-major
//                               (B*H, F, 272), k' transposed (B*H, 272, Fp)
//
// Non-causal linear attention (pcmer.py:69-78) then is two batched GEMMs:
//   ctxT[z] (80 x 272)  = [v^T; 1; 0][z] (80 x Fp) * k'^T[z] (272 x Fp)^T          row 64 = k_sum
//   out[z]  (F x 80)    = q'[z] (F x 272) * ctxT[z] (80 x 272)^T                    column 64 = q' . k_sum
#pragma once
#include "gemm_tc.cuh"

namespace ddsp {
namespace tc {

enum { EPI_PLAIN = 0, EPI_QKV = 1, EPI_OUT = 2, EPI_GLU = 3 };

constexpr int kFeat = 266;          // int(64 * ln 64) random features (pcmer.py:166)
constexpr int kFeatPad = 272;       // padded to a multiple of 16 (UMMA N) / 16 bytes
constexpr int kVtRows = 80;         // 64 rows of v^T + the row of ones + zero padding to a multiple of 16

struct GemmParams {
    int Z, M, N, K;                 // per-batch problem
    int tiles_m, tiles_n;           // per batch
    int w_presplit;                 // W arrives as two tensors (hi, lo) split once on the host side
    int a_presplit;                 // A arrives as two tensors (raw = hi, lo) written by the producing epilogue
    int w_batched;                  // W has a batch coordinate (else the same W for every batch)
    const float* bias;              // (N) or null
    const float* residual;          // (M, N), row stride ldr, or null            (Z == 1)
    int64_t ldr;
    const float* ln_gamma;          // LayerNorm over the N columns as second output (tiles_n == 1), or null
    const float* ln_beta;
    float ln_eps;
    int l2_prefetch;                // > 0: the producer prefetches the DRAM-streamed operand tiles this many k-blocks ahead into L2
    int split_out;                  // EPI_PLAIN: also write the low TF32 term of C through map_c2 (C itself is the high term)
    // EPI_QKV / EPI_OUT
    float* vt;                      // (B, H, 80, Fp)
    float* vt_lo;                   // optional: low TF32 term of v^T, same layout (pre-split operand of the context GEMM)
    float* q; float* k;             // (B, H, F, 64): the same tensors map_c / map_c2 describe (EPI_QKV)
    int frames, frames_pad, heads;
};

// ---- PTX wrappers (beyond gemm_tc.cuh) ------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
// the tile a later tma_load_3d will fetch, pulled into L2 now (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
// The tensor core reads an fp32 word as TF32 by ignoring its low 13 mantissa bits, so the raw tile IS the high term
// (hi = x truncated to TF32) and only the low term lo = x - hi (exact in fp32, |lo| < 2^-10 |x|) has to be written:
// one 128-bit shared load + one 128-bit shared store per four elements.
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void split_tile_lo(uint32_t raw, uint32_t lo, int n_vec, int t) {
#pragma unroll 4
    for (int i = t; i < n_vec; i += 128) {
        const float4 x = ld_shared_v4(raw + 16 * i);
        st_shared_v4(lo + 16 * i, tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
    }
}

__host__ __device__ constexpr uint32_t umma_idesc_tf32_n(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

// 64-byte-swizzled K-major tile (rows of 16 fp32, 8-row groups of 512 B): layout type SWIZZLE_64B = 4, stride byte offset 512
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
template <int BK>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) {
    return BK == 32 ? umma_desc_sw128(smem_addr) : umma_desc_sw64(smem_addr);
}

// One warp's 32 rows x 32 columns (a row per lane) leave through a 128-byte-swizzled staging slice and one TMA
// store: coalesced 128-byte rows in global memory, rows / columns beyond the tensor edge are clipped by the TMA
// unit.  Two slices per warp alternate; `wait_read<1>` makes sure the store that last read this slice is done.
struct RowStore {
    uint32_t base;      // shared address of this warp's two 4-KB slices (1024-byte aligned)
    int count;
    int lane;
    int slices;         // 2, or 1 where a fourth pipeline stage is worth more than overlapped stores (attention output GEMM)
    __device__ __forceinline__ uint32_t begin() {
        const uint32_t buf = base + (slices == 2 ? (count & 1) * 4096 : 0);
        if (lane == 0) {
            if (slices == 2) bulk_wait_read<1>();
            else bulk_wait_read<0>();
        }
        __syncwarp();
        return buf;
    }
    __device__ __forceinline__ void fill(uint32_t buf, const float (&v)[32]) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
            st_shared_v4(buf + lane * 128 + ((c ^ (lane & 7)) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        fence_proxy_async();
        __syncwarp();
        ++count;
    }
    __device__ __forceinline__ void drain() {
        if (lane == 0) bulk_wait_all<0>();
        __syncwarp();
    }
};

// AM = rows of A that exist (and are loaded) per tile.  The MMA always reads 128 rows: with AM < 128 (the 80-row v^T
// operand of the context GEMM) rows AM..127 alias the start of the next buffer of the same stage -- finite numbers that only
// reach accumulator rows nobody stores -- and the bytes saved buy a fourth pipeline stage.
// BK = fp32 elements of one k-block (32: 128-byte swizzle rows; 16: 64-byte rows -- half-sized stages, twice as many of them
// in the same shared memory, for the tiles whose two 96-KB stages leave the ring empty half of the time).
// NSTACK: the products hi*hi and hi*lo(W) as ONE MMA of N = 2 BN over the adjacent W_hi | W_lo tiles of a stage (two
// accumulator column ranges, added in the epilogue), lo(A)*hi as a second MMA of N = BN.  The 3xTF32 GEMMs are bound by the
// shared-memory port (every MMA re-reads its (128 + N) x 32 bytes of operands, profiles/ubench/umma_issue.cu): two MMAs
// instead of three read the A tile twice instead of three times per k-step -- for the narrow attention GEMMs (BN <= 128).
template <int BN, int AM = kBM, int SLICES = 2, int BK = kBK, bool NSTACK = false>
struct GCfg {
    static_assert(!NSTACK || 2 * BN <= 256, "stacked W_hi | W_lo operand: one MMA of N = 2 BN");
    static_assert(BK == 32 || BK == 16, "k-block = one 128-byte or 64-byte swizzle row");
    static constexpr int kRowBytes = BK * 4;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128: multiple of 16 in [16, 256]");
    static_assert(AM % 8 == 0 && AM <= kBM && (kBM - AM) <= AM && (kBM - AM) <= BN,
                  "aliased rows of A_hi stay inside A_lo, those of A_lo inside W_hi (never in a buffer the splitter writes)");
    static constexpr int kABytes = AM * kRowBytes;
    static constexpr int kWBytes = BN * kRowBytes;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kWBytes;   // A_hi | A_lo | W_hi | W_lo
    static constexpr int kStoreBytes = 4 * SLICES * 4096;           // four epilogue warps x staging slices
    static constexpr int kBarBytes = 256;
    static constexpr int kBudget = 227 * 1024 - 1024 - kBarBytes - kStoreBytes;
    static constexpr int kStages = kBudget / kStageBytes < 2 ? 2 : (kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes);
    static constexpr int kAccN = NSTACK ? 2 * BN : BN;               // accumulator columns in use
    static constexpr int kAccCols = kAccN <= 32 ? 32 : kAccN <= 64 ? 64 : kAccN <= 128 ? 128 : 256;
    static constexpr int kTmemCols = 2 * kAccCols;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStoreBytes + kBarBytes + 1024;
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int BN, int EPI, int AM = kBM, int SLICES = 2, int BK = kBK, bool NSTACK = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a_lo,
              const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w_lo,
              const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_c2, const GemmParams P) {
    using C = GCfg<BN, AM, SLICES, BK, NSTACK>;
    static_assert(EPI != EPI_PLAIN || BN % 32 == 0, "the row-store epilogue works in chunks of 32 columns");
    static_assert(EPI != EPI_QKV || BN % 64 == 0, "head-split epilogue: a column tile holds whole heads");
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* store_smem = smem + C::kStages * C::kStageBytes;
    unsigned char* bars = store_smem + C::kStoreBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bars);
    uint64_t* split_bar = full_bar + C::kStages;
    uint64_t* empty_bar = split_bar + C::kStages;
    uint64_t* acc_full = empty_bar + C::kStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    static_assert((3 * C::kStages + 4) * 8 + 4 <= C::kBarBytes, "barrier block too small");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kb = (P.K + BK - 1) / BK;
    const int tiles_per_z = P.tiles_m * P.tiles_n;
    const int n_tiles = P.Z * tiles_per_z;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_w);
        if (P.w_presplit) tma_prefetch_desc(&map_w_lo);
        tma_prefetch_desc(&map_c);
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
            const uint32_t tx = (P.a_presplit ? 2 : 1) * C::kABytes + (P.w_presplit ? 2 : 1) * C::kWBytes;
            // The shared-memory ring holds 2..3 k-blocks: less than a DRAM latency of main-loop time for the narrow batched
            // GEMMs of the attention (K = frames or features, operands streamed once from HBM).  A second cursor runs
            // `l2_prefetch` k-blocks ahead of the loads and pulls the streamed tiles into L2, so that the ring is refilled at
            // L2 latency.  Shared weights (not batched) stay L2-resident by themselves and are not prefetched.
            int ptile = blockIdx.x, pkb = 0;
            auto prefetch_next = [&]() {
                if (ptile >= n_tiles) return;
                const int z = ptile / tiles_per_z, rem = ptile - z * tiles_per_z;
                const int m0 = (rem / P.tiles_n) * kBM, n0 = (rem % P.tiles_n) * BN;
                tma_prefetch_3d(&map_a, pkb * BK, m0, z);
                if (P.a_presplit) tma_prefetch_3d(&map_a_lo, pkb * BK, m0, z);
                if (P.w_batched) {
                    tma_prefetch_3d(&map_w, pkb * BK, n0, z);
                    if (P.w_presplit) tma_prefetch_3d(&map_w_lo, pkb * BK, n0, z);
                }
                if (++pkb == n_kb) { pkb = 0; ptile += gridDim.x; }
            };
            for (int i = 0; i < P.l2_prefetch; ++i) prefetch_next();
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int z = tile / tiles_per_z, rem = tile - z * tiles_per_z;
                const int m0 = (rem / P.tiles_n) * kBM, n0 = (rem % P.tiles_n) * BN;
                const int wz = P.w_batched ? z : 0;
                for (int kb = 0; kb < n_kb; ++kb) {
                    if (P.l2_prefetch > 0) prefetch_next();
                    mbar_wait(s32(empty_bar + stage), phase ^ 1);
                    const uint32_t st = s32(smem + stage * C::kStageBytes);
                    mbar_arrive_expect_tx(s32(full_bar + stage), tx);
                    tma_load_3d(st, &map_a, kb * BK, m0, z, s32(full_bar + stage));
                    if (P.a_presplit) tma_load_3d(st + C::kABytes, &map_a_lo, kb * BK, m0, z, s32(full_bar + stage));
                    tma_load_3d(st + 2 * C::kABytes, &map_w, kb * BK, n0, wz, s32(full_bar + stage));
                    if (P.w_presplit)
                        tma_load_3d(st + 2 * C::kABytes + C::kWBytes, &map_w_lo, kb * BK, n0, wz, s32(full_bar + stage));
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = umma_idesc_tf32_n(BN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(s32(acc_empty + acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + acc * C::kAccCols;
            for (int kb = 0; kb < n_kb; ++kb) {
                const uint32_t st = s32(smem + stage * C::kStageBytes);
                const uint64_t a_hi = umma_desc_kmajor<BK>(st), a_lo = umma_desc_kmajor<BK>(st + C::kABytes);
                const uint64_t w_hi = umma_desc_kmajor<BK>(st + 2 * C::kABytes), w_lo = umma_desc_kmajor<BK>(st + 2 * C::kABytes + C::kWBytes);
                // The raw tile is the high term: hi*hi -- and every product whose low operand arrived pre-split -- can start
                // the moment TMA has landed; only the products that need a low term from the splitter wait for it, so the
                // splitter's latency hides behind the early MMAs.
                mbar_wait(s32(full_bar + stage), phase);
                tc_fence_after();
                if constexpr (NSTACK) {
                    constexpr uint32_t idesc2 = umma_idesc_tf32_n(2 * BN);       // W_hi | W_lo: 2 BN adjacent rows of the stage
                    if (P.w_presplit) {
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < BK / 8; ++kk) umma_tf32(d, a_hi + (uint64_t)(2 * kk), w_hi + (uint64_t)(2 * kk), idesc2, (kb | kk) != 0);
                        }
                        __syncwarp();
                    }
                    mbar_wait(s32(split_bar + stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < BK / 8; ++kk) {
                            const uint64_t o = (uint64_t)(2 * kk);
                            if (!P.w_presplit) umma_tf32(d, a_hi + o, w_hi + o, idesc2, (kb | kk) != 0);
                            umma_tf32(d, a_lo + o, w_hi + o, idesc, 1);
                        }
                        umma_commit(s32(empty_bar + stage));
                        if (kb == n_kb - 1) umma_commit(s32(acc_full + acc));
                    }
                    __syncwarp();
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                    continue;
                }
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < BK / 8; ++kk) {
                        const uint64_t o = (uint64_t)(2 * kk);
                        umma_tf32(d, a_hi + o, w_hi + o, idesc, (kb | kk) != 0);
                        if (P.w_presplit) umma_tf32(d, a_hi + o, w_lo + o, idesc, 1);
                        if (P.a_presplit) umma_tf32(d, a_lo + o, w_hi + o, idesc, 1);
                    }
                }
                __syncwarp();
                mbar_wait(s32(split_bar + stage), phase);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < BK / 8; ++kk) {
                        const uint64_t o = (uint64_t)(2 * kk);
                        if (!P.w_presplit) umma_tf32(d, a_hi + o, w_lo + o, idesc, 1);
                        if (!P.a_presplit) umma_tf32(d, a_lo + o, w_hi + o, idesc, 1);
                    }
                    umma_commit(s32(empty_bar + stage));
                    if (kb == n_kb - 1) umma_commit(s32(acc_full + acc));
                }
                __syncwarp();
                if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= kSplitWarp0) {
        // ===== splitter: x -> hi (in place) + lo =====
        const int t = threadIdx.x - kSplitWarp0 * 32;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(s32(full_bar + stage), phase);
                const uint32_t st = s32(smem + stage * C::kStageBytes);
                if (!P.a_presplit) split_tile_lo(st, st + C::kABytes, C::kABytes / 16, t);
                if (!P.w_presplit) split_tile_lo(st + 2 * C::kABytes, st + 2 * C::kABytes + C::kWBytes, C::kWBytes / 16, t);
                fence_proxy_async();
                mbar_arrive(s32(split_bar + stage));
                if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue (warps 2..5): TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int row_in_tile = q * 32 + lane;
        RowStore rs;
        rs.base = s32(store_smem + (warp - kEpiWarp0) * (SLICES * 4096));
        rs.slices = SLICES;
        rs.count = 0;
        rs.lane = lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int z = tile / tiles_per_z, rem = tile - z * tiles_per_z;
            const int m0 = (rem / P.tiles_n) * kBM, n0 = (rem % P.tiles_n) * BN;
            const int row = m0 + row_in_tile;
            mbar_wait(s32(acc_full + acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::kAccCols;
            uint32_t r[32];
            float v[32];

            if constexpr (EPI == EPI_PLAIN) {
                const bool row_ok = row < P.M;
                const float* rrow = (P.residual && row_ok) ? P.residual + (int64_t)row * P.ldr : nullptr;
                const bool ln = P.ln_gamma != nullptr;
                float sum = 0.0f;
                // the residual of chunk c+1 is loaded while chunk c is staged and stored (its latency was exposed once per chunk)
                float4 e[8];
                auto load_res = [&](int n) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) e[j] = *reinterpret_cast<const float4*>(rrow + n + 4 * j);
                };
                if (rrow) load_res(n0);
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    const int n = n0 + c0;
                    if (n >= P.N) break;
                    tmem_ld32(taddr + c0, r);
                    if constexpr (NSTACK) {
                        uint32_t r2[32];
                        tmem_ld32(taddr + BN + c0, r2);            // the hi*lo(W) column range
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
                    } else {
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    }
                    if (P.bias) {
                        if (n + 32 <= P.N) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + n + j));
                                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += (n + j < P.N) ? __ldg(P.bias + n + j) : 0.0f;
                        }
                    }
                    if (rrow) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[4 * j] += e[j].x; v[4 * j + 1] += e[j].y; v[4 * j + 2] += e[j].z; v[4 * j + 3] += e[j].w;
                        }
                        if (c0 + 32 < BN && n + 32 < P.N) load_res(n + 32);
                    }
                    if (ln) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { sum += v[j]; r[j] = __float_as_uint(v[j]); }
                        tmem_st32(taddr + c0, r);
                    }
                    const uint32_t buf = rs.begin();
                    rs.fill(buf, v);
                    if (lane == 0) { tma_store_3d(&map_c, buf, n, m0 + q * 32, z); bulk_commit(); }
                    if (P.split_out) {
                        // the consumer GEMM reads C as the high term (the tensor core ignores the low 13 mantissa bits);
                        // its low term leaves here, so that the consumer needs no splitter pass over this operand
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = tf32_lo(v[j]);
                        const uint32_t buf2 = rs.begin();
                        rs.fill(buf2, v);
                        if (lane == 0) { tma_store_3d(&map_c2, buf2, n, m0 + q * 32, z); bulk_commit(); }
                    }
                }
                if (ln) {
                    // LayerNorm of the finished row (torch.nn.LayerNorm: biased variance, eps inside the sqrt), two more
                    // passes over the row parked in TMEM: exact mean first, then the centred sum of squares (a one-pass
                    // shifted variance was measured: no faster -- the epilogue hides behind the next tile's main loop --
                    // and 2x the error)
                    tmem_st_wait();
                    const float inv_n = 1.0f / (float)P.N;
                    const float mean = sum * inv_n;
                    float ssq = 0.0f;
#pragma unroll 1
                    for (int c0 = 0; c0 < P.N; c0 += 32) {
                        tmem_ld32(taddr + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) { const float dlt = __uint_as_float(r[j]) - mean; ssq = fmaf(dlt, dlt, ssq); }
                    }
                    const float rstd = rsqrtf(ssq * inv_n + P.ln_eps);
#pragma unroll 1
                    for (int c0 = 0; c0 < P.N; c0 += 32) {
                        tmem_ld32(taddr + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 g = __ldg(reinterpret_cast<const float4*>(P.ln_gamma + c0 + j));
                            const float4 b = __ldg(reinterpret_cast<const float4*>(P.ln_beta + c0 + j));
                            v[j] = fmaf((__uint_as_float(r[j]) - mean) * rstd, g.x, b.x);
                            v[j + 1] = fmaf((__uint_as_float(r[j + 1]) - mean) * rstd, g.y, b.y);
                            v[j + 2] = fmaf((__uint_as_float(r[j + 2]) - mean) * rstd, g.z, b.z);
                            v[j + 3] = fmaf((__uint_as_float(r[j + 3]) - mean) * rstd, g.w, b.w);
                        }
                        const uint32_t buf = rs.begin();
                        rs.fill(buf, v);
                        if (lane == 0) { tma_store_3d(&map_c2, buf, c0, m0 + q * 32, z); bulk_commit(); }
                    }
                }
            } else if constexpr (EPI == EPI_QKV) {
                // N = 3 * heads * 64: column block -> (q | k | v, head); BN is a multiple of 64
                const int F = P.frames, H = P.heads;
                const int inner = H * 64;
                const int which = n0 / inner;                       // 0 q, 1 k, 2 v (tiles never straddle: inner % BN == 0)
                const int r0 = m0 + q * 32;                         // first row of this warp's slice
                const int b0 = r0 / F, f0 = r0 - b0 * F;
                const int bb = row / F, ff = row - bb * F;
                const int n_clips = P.M / F;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    const int n = n0 + c0;
                    const int col = n - which * inner;
                    const int h = col >> 6, e0 = col & 63;
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (P.bias) b = __ldg(reinterpret_cast<const float4*>(P.bias + n + j));
                        v[j] = __uint_as_float(r[j]) + b.x;
                        v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
                        v[j + 2] = __uint_as_float(r[j + 2]) + b.z;
                        v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
                    }
                    if (which < 2) {
                        const uint32_t buf = rs.begin();
                        rs.fill(buf, v);
                        // a slice that runs past the end of its clip continues in the next clip: the TMA store clips the
                        // tail rows, and the few rows of the next clip are stored directly (a TMA store must not start at a
                        // negative coordinate)
                        const bool straddle = f0 + 32 > F;
                        if (lane == 0) {
                            if (b0 < n_clips) {
                                if (which == 0) tma_store_4d(&map_c, buf, e0, f0, h, b0);
                                else tma_store_4d(&map_c2, buf, e0, f0, h, b0);
                            }
                            bulk_commit();
                        }
                        if (straddle && bb > b0 && row < P.M) {
                            float* dst = (which == 0 ? P.q : P.k) + ((((int64_t)bb * H + h) * F + ff) * 64 + e0);
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    } else if (row < P.M) {
                        // v^T: lanes hold consecutive frames -> 128-byte coalesced stores along the frame axis
                        const int64_t off = (((int64_t)bb * H + h) * kVtRows + e0) * P.frames_pad + ff;
                        float* dst = P.vt + off;
#pragma unroll
                        for (int j = 0; j < 32; ++j) dst[(int64_t)j * P.frames_pad] = v[j];
                        if (P.vt_lo) {
                            float* dlo = P.vt_lo + off;
#pragma unroll
                            for (int j = 0; j < 32; ++j) dlo[(int64_t)j * P.frames_pad] = tf32_lo(v[j]);
                        }
                    }
                }
            } else if constexpr (EPI == EPI_GLU) {
                // column tile nt = [value channels 128 nt .. 128 nt + 127 | gate channels of the same range]; bias in the
                // same interleaved order.  Output (M, N / 2): channel 128 nt + c.
                static_assert(BN == 256, "GLU epilogue: 128 value + 128 gate columns per tile");
                uint32_t rg[32];
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    tmem_ld32(taddr + c0, r);
                    tmem_ld32(taddr + 128 + c0, rg);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
                        if (P.bias) {
                            ba = __ldg(reinterpret_cast<const float4*>(P.bias + n0 + c0 + j));
                            bg = __ldg(reinterpret_cast<const float4*>(P.bias + n0 + 128 + c0 + j));
                        }
                        const float a0 = __uint_as_float(r[j]) + ba.x, a1 = __uint_as_float(r[j + 1]) + ba.y;
                        const float a2 = __uint_as_float(r[j + 2]) + ba.z, a3 = __uint_as_float(r[j + 3]) + ba.w;
                        const float g0 = __uint_as_float(rg[j]) + bg.x, g1 = __uint_as_float(rg[j + 1]) + bg.y;
                        const float g2 = __uint_as_float(rg[j + 2]) + bg.z, g3 = __uint_as_float(rg[j + 3]) + bg.w;
                        v[j] = __fdividef(a0, 1.0f + __expf(-g0));
                        v[j + 1] = __fdividef(a1, 1.0f + __expf(-g1));
                        v[j + 2] = __fdividef(a2, 1.0f + __expf(-g2));
                        v[j + 3] = __fdividef(a3, 1.0f + __expf(-g3));
                    }
                    const uint32_t buf = rs.begin();
                    rs.fill(buf, v);
                    if (lane == 0) { tma_store_3d(&map_c, buf, (n0 >> 1) + c0, m0 + q * 32, z); bulk_commit(); }
                }
            } else {   // EPI_OUT
                // batch z = (clip, head); accumulator columns 0..63 = sum_j q'_j ctx_j, column 64 = q' . k_sum
                const int H = P.heads;
                const int b = z / H, h = z - b * H;
                uint32_t r2[32];
                tmem_ld32(taddr + 64, r);
                if constexpr (NSTACK) tmem_ld32(taddr + BN + 64, r2);
                tmem_ld_wait();
                float den = __uint_as_float(r[0]);
                if constexpr (NSTACK) den += __uint_as_float(r2[0]);
                const float d_inv = 1.0f / (den + 1e-8f);                             // pcmer.py:72
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    tmem_ld32(taddr + c0, r);
                    if constexpr (NSTACK) tmem_ld32(taddr + BN + c0, r2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = __uint_as_float(r[j]);
                        if constexpr (NSTACK) x += __uint_as_float(r2[j]);
                        v[j] = x * d_inv;
                    }
                    const uint32_t buf = rs.begin();
                    rs.fill(buf, v);
                    if (lane == 0) { tma_store_3d(&map_c, buf, h * 64 + c0, m0 + q * 32, b); bulk_commit(); }
                }
            }
            tc_fence_before();
            mbar_arrive(s32(acc_empty + acc));
        }
        rs.drain();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<C::kTmemCols>(tmem_base);
}

// ---- FAVOR+ feature map as a GEMM epilogue ---------------------------------------------------------------------------
// x: (Z, F, 64) rows of one head (q or k with its Linear bias already added), W = 64^-0.25 * projection (266 x 64),
// resident in shared memory as hi | lo, both k-blocks.  Per 128-row tile: dash = x W^T in TMEM (272 columns as two halves,
// N = 128 and N = 144 per k-step, through a ring of three accumulator slots), then
//   q: ratio * (exp(dash - diag - max_j dash) + eps)      -> (Z, F, 272) row-major through TMA stores
//   k: ratio * exp(dash - diag + eps)                     -> (Z, 272, Fp) transposed, lanes = consecutive frames
// with diag = |x|^2 / (2 sqrt(64)), ratio = 266^-0.5 (pcmer.py:124-160).  Columns 266..271 of q' are written as
// zeros; rows 266..271 of k'^T are never written (the buffer is zero-initialised once by the host side).
struct FeatParams {
    const float* x;             // (Z, F, 64) (the kernels read it through map_a only)
    float* kt;                  // k: (Z, 272, Fp)
    int Z, F, Fp;
    int tiles_m;                // per batch
    float eps;
};

constexpr int kFeatHalf0 = 128;                                              // features of the first accumulator half (N = 128 | 144)
constexpr int kFeatSlots = 3, kFeatSlotCols = 144;                           // ring of accumulator halves in TMEM (432 of 512 columns)
static_assert(kFeatPad - kFeatHalf0 <= kFeatSlotCols && kFeatSlots * kFeatSlotCols + 16 <= 512, "TMEM budget");
constexpr int kFeatWBytes = kFeatPad * 128;                                  // one k-block of W (hi or lo)
constexpr int kFeatABytes = kBM * 128;
constexpr int kFeatSmemW = 4 * kFeatWBytes;                                  // hi kb0 | hi kb1 | lo kb0 | lo kb1
constexpr int kFeatSmemA = 4 * kFeatABytes;                                  // hi kb0 | hi kb1 | lo kb0 | lo kb1
constexpr int kFeatStoreBytes = 4 * 4096;                                    // one staging slice per epilogue warp
constexpr int kFeatDiagRing = 4;                                             // |x|^2 terms of the tiles in flight between splitter and epilogue
constexpr int kFeatSmemBytes = kFeatSmemW + kFeatSmemA + kFeatStoreBytes + 256 + kFeatDiagRing * kBM * 4 + 1024;
static_assert(kFeatSmemBytes <= 227 * 1024, "shared memory budget");

// Eight splitter warps (6..13): load -> split -> MMA of consecutive tiles is a serial chain through the single A buffer and
// sets the time per tile (the epilogue warps wait about half of the time), so the split is spread over 256 threads and the
// MMA issuer starts the products that need no low term of A (hi*hi, hi*lo(W): two thirds of the work) as soon as TMA has
// landed.  Measured and dropped: a second epilogue warp group for k' (126 -> 128 us).
template <bool IS_Q>
__host__ __device__ constexpr int feat_threads() { return kThreads + 128; }
constexpr int kFeatSplitThreads = 256;

template <bool IS_Q>
__global__ void __launch_bounds__(feat_threads<IS_Q>(), 1)
favor_features_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_c, const FeatParams P) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* w_smem = smem;
    unsigned char* a_smem = smem + kFeatSmemW;
    unsigned char* store_smem = a_smem + kFeatSmemA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(store_smem + kFeatStoreBytes);
    uint64_t* w_bar = bars;            // TMA -> everyone (W landed)
    uint64_t* full_bar = bars + 1;     // TMA -> splitter
    uint64_t* split_bar = bars + 2;    // splitter -> MMA
    uint64_t* empty_bar = bars + 3;    // MMA -> TMA
    uint64_t* acc_full = bars + 4;     // MMA -> epilogue    [kFeatSlots]
    uint64_t* acc_empty = bars + 7;    // epilogue -> MMA    [kFeatSlots]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    uint64_t* lo_empty = bars + 11;    // MMA -> splitter (the low-term buffer of A is free)
    float* diag_ring = reinterpret_cast<float*>(store_smem + kFeatStoreBytes + 256);     // [kFeatDiagRing][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = P.Z * P.tiles_m;

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(s32(w_bar), 1);
            mbar_init(s32(full_bar), 1);
            mbar_init(s32(split_bar), kFeatSplitThreads);
            mbar_init(s32(empty_bar), 1);
            mbar_init(s32(lo_empty), 1);
            for (int i = 0; i < kFeatSlots; ++i) {
                mbar_init(s32(acc_full + i), 1);
                mbar_init(s32(acc_empty + i), 128);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(s32(tmem_slot));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // resident projection: both k-blocks of all 272 rows (rows 266..271 are out of bounds: zero-filled)
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_w);
        mbar_arrive_expect_tx(s32(w_bar), 2 * kFeatWBytes);
        for (int kb = 0; kb < 2; ++kb) {      // box = 136 rows (17 swizzle atoms): two boxes per k-block
            tma_load_3d(s32(w_smem + kb * kFeatWBytes), &map_w, kb * kBK, 0, 0, s32(w_bar));
            tma_load_3d(s32(w_smem + kb * kFeatWBytes + 136 * 128), &map_w, kb * kBK, 136, 0, s32(w_bar));
        }
    }
    mbar_wait(s32(w_bar), 0);
    {
        const uint32_t hi = s32(w_smem), lo = hi + 2 * kFeatWBytes;
        for (int i = threadIdx.x; i < 2 * kFeatWBytes / 16; i += feat_threads<IS_Q>()) {
            const float4 x = ld_shared_v4(hi + 16 * i);
            st_shared_v4(lo + 16 * i, tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
        }
        fence_proxy_async();
    }
    __syncthreads();

    if (warp == 0) {
        // ===== TMA producer: one A tile (two k-blocks) at a time =====
        if (lane == 0) {
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int z = tile / P.tiles_m, m0 = (tile - z * P.tiles_m) * kBM;
                // The single A buffer makes load -> split -> MMA of consecutive tiles a serial chain (it, not the epilogue,
                // sets the time per tile): the next tile is pulled into L2 while this one is in flight, so that its load
                // costs an L2 instead of an HBM latency.
                const int nt = tile + gridDim.x;
                if (nt < n_tiles) {
                    const int nz = nt / P.tiles_m, nm0 = (nt - nz * P.tiles_m) * kBM;
                    tma_prefetch_3d(&map_a, 0, nm0, nz);
                    tma_prefetch_3d(&map_a, kBK, nm0, nz);
                }
                mbar_wait(s32(empty_bar), phase ^ 1);
                mbar_arrive_expect_tx(s32(full_bar), 2 * kFeatABytes);
                tma_load_3d(s32(a_smem), &map_a, 0, m0, z, s32(full_bar));
                tma_load_3d(s32(a_smem + kFeatABytes), &map_a, kBK, m0, z, s32(full_bar));
                phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // A tile's 272 features are issued as two halves (features 0..127 with N = 128, 128..271 with N = 144), each into
        // its own slot of a ring of three 144-column accumulators: while the epilogue drains the halves of tile t, the halves
        // of tile t+1 are already being computed (one 272-column accumulator would fit TMEM only once, and MMA and epilogue
        // of a tile then run back to back).
        constexpr uint32_t idesc_h[2] = {umma_idesc_tf32_n(kFeatHalf0), umma_idesc_tf32_n(kFeatPad - kFeatHalf0)};
        uint32_t phase = 0;
        uint32_t g = 0;                    // running half index: slot = g % 3, use = g / 3
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            // Two thirds of the products need only the raw tile (= high term of A): they are issued for both halves as soon as
            // TMA has landed, and the raw buffer is handed back to the producer behind them -- the next tile's load then
            // overlaps the low-term products of this one instead of waiting for the whole tile.
            mbar_wait(s32(full_bar), phase);
            uint32_t d_h[2], slot_h[2];
#pragma unroll
            for (int half = 0; half < 2; ++half, ++g) {
                const uint32_t slot = g % kFeatSlots, use = g / kFeatSlots;
                slot_h[half] = slot;
                d_h[half] = tmem_base + slot * kFeatSlotCols;
                mbar_wait(s32(acc_empty + slot), (use & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t a_hi = umma_desc_sw128(s32(a_smem + kb * kFeatABytes));
                        const uint32_t wh = s32(w_smem + kb * kFeatWBytes), wl = s32(w_smem + (2 + kb) * kFeatWBytes);
                        const uint64_t w_hi = umma_desc_sw128(wh + half * kFeatHalf0 * 128), w_lo = umma_desc_sw128(wl + half * kFeatHalf0 * 128);
#pragma unroll
                        for (int kk = 0; kk < kBK / 8; ++kk) {
                            const uint64_t o = (uint64_t)(2 * kk);
                            umma_tf32(d_h[half], a_hi + o, w_lo + o, idesc_h[half], (kb | kk) != 0);
                            umma_tf32(d_h[half], a_hi + o, w_hi + o, idesc_h[half], 1);
                        }
                    }
                }
                __syncwarp();
            }
            mbar_wait(s32(split_bar), phase);                          // the low term is in place, and the splitter has read the raw tile
            tc_fence_after();
            if (elect_one()) {
                umma_commit(s32(empty_bar));                           // raw buffer free once the products above have read it
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t a_lo = umma_desc_sw128(s32(a_smem + (2 + kb) * kFeatABytes));
                        const uint64_t w_hi = umma_desc_sw128(s32(w_smem + kb * kFeatWBytes) + half * kFeatHalf0 * 128);
#pragma unroll
                        for (int kk = 0; kk < kBK / 8; ++kk) {
                            const uint64_t o = (uint64_t)(2 * kk);
                            umma_tf32(d_h[half], a_lo + o, w_hi + o, idesc_h[half], 1);
                        }
                    }
                    umma_commit(s32(acc_full + slot_h[half]));
                }
                umma_commit(s32(lo_empty));
            }
            __syncwarp();
            phase ^= 1;
        }
    } else if (warp >= kSplitWarp0) {
        // ===== splitter =====
        const int t = threadIdx.x - kSplitWarp0 * 32;
        // The splitter reads every element of the tile anyway, so it also forms diag = |x|^2 / (2 sqrt(64)) of the 128 rows
        // (the epilogue used to re-read its row from global memory: a load latency per tile in front of the exponentials).
        // Thread t holds float4 column-chunks of rows t/8 + 32 m in both k-blocks; the 8 threads of a row add up by shuffles.
        // The values travel splitter -> split_bar -> MMA -> acc_full -> epilogue; the epilogue runs at most two tiles behind.
        uint32_t phase = 0;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            mbar_wait(s32(full_bar), phase);
            mbar_wait(s32(lo_empty), phase ^ 1);                       // the low-term products of the previous tile are done
            const uint32_t raw = s32(a_smem), lo = raw + 2 * kFeatABytes;
            float ssq[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                float acc = 0.0f;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const int i = t + kFeatSplitThreads * (m + 4 * kb);     // 1024 float4 per k-block: row (i % 1024) / 8
                    const float4 x = ld_shared_v4(raw + 16 * i);
                    st_shared_v4(lo + 16 * i, tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
                    acc = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, acc))));
                }
                ssq[m] = acc;
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                ssq[m] += __shfl_xor_sync(0xffffffffu, ssq[m], 1);
                ssq[m] += __shfl_xor_sync(0xffffffffu, ssq[m], 2);
                ssq[m] += __shfl_xor_sync(0xffffffffu, ssq[m], 4);
            }
            if ((t & 7) == 0) {
                float* dr = diag_ring + (it % kFeatDiagRing) * kBM + (t >> 3);
#pragma unroll
                for (int m = 0; m < 4; ++m) dr[32 * m] = 0.0625f * ssq[m];
            }
            fence_proxy_async();
            mbar_arrive(s32(split_bar));
            phase ^= 1;
        }
    } else {
        // ===== epilogue =====
        const int q = warp & 3;
        const int row_in_tile = q * 32 + lane;
        const uint32_t stage_buf = s32(store_smem + (warp - kEpiWarp0) * 4096);            // q' only
        const float ratio = rsqrtf((float)kFeat);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int z = tile / P.tiles_m, m0 = (tile - z * P.tiles_m) * kBM;
            const int f = m0 + row_in_tile;
            const bool ok = f < P.F;
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
            const uint32_t g0 = 2 * it, g1 = g0 + 1;
            const uint32_t slot_h[2] = {g0 % kFeatSlots, g1 % kFeatSlots};
            const uint32_t par_h[2] = {(g0 / kFeatSlots) & 1, (g1 / kFeatSlots) & 1};
            mbar_wait(s32(acc_full + slot_h[0]), par_h[0]);
            tc_fence_after();
            // diag = |x|^2 * 64^-0.5 / 2, formed by the splitter from the tile in shared memory
            const float diag = diag_ring[(it % kFeatDiagRing) * kBM + row_in_tile];
            // TMEM reads are software-pipelined inside a half: the load of chunk c+1 is in flight while chunk c is processed
            // (tcgen05.wait::ld covers every load issued before it, so the next load is issued right after the wait)
            uint32_t ra[32], rb[32];
            constexpr float kLog2e = 1.4426950408889634f;
            float mx = -3.0e38f;
            if (IS_Q) {
                // pass 1: row maximum over all 266 features (both halves must be complete before anything is written)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    constexpr int kCh0 = kFeatHalf0 / 32;
                    const int n_ch = half ? (kFeatPad - kFeatHalf0 + 31) / 32 : kCh0;
                    if (half) {
                        mbar_wait(s32(acc_full + slot_h[half]), par_h[half]);
                        tc_fence_after();
                    }
                    const uint32_t taddr = lane_base + slot_h[half] * kFeatSlotCols;
                    tmem_ld32(taddr, ra);
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        if (c < n_ch) {
                            uint32_t(&cur)[32] = (c & 1) ? rb : ra;
                            uint32_t(&nxt)[32] = (c & 1) ? ra : rb;
                            tmem_ld_wait();
                            if (c + 1 < n_ch) tmem_ld32(taddr + 32 * (c + 1), nxt);
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (half * kFeatHalf0 + 32 * c + j < kFeat) mx = fmaxf(mx, __uint_as_float(cur[j]));
                        }
                    }
                }
            }
            // exp(dash + shift) = 2^(dash * log2e + shift * log2e): one FFMA + one MUFU.EX2 per element, then ratio * (e + eps);
            // k' has no additive term, so its ratio is folded into the exponent
            const float shift2 = IS_Q ? -(diag + mx) * kLog2e : fmaf(P.eps - diag, kLog2e, -0.5f * log2f((float)kFeat));
            const float add = ratio * P.eps;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (!IS_Q && half) {
                    mbar_wait(s32(acc_full + slot_h[half]), par_h[half]);
                    tc_fence_after();
                }
                constexpr int kCh0 = kFeatHalf0 / 32;
                const int n_ch = half ? (kFeatPad - kFeatHalf0 + 31) / 32 : kCh0;     // 4 | 5 (the last one reads 16 columns past the half: ignored)
                const uint32_t taddr = lane_base + slot_h[half] * kFeatSlotCols;
                tmem_ld32(taddr, ra);
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    if (c < n_ch) {
                        uint32_t(&cur)[32] = (c & 1) ? rb : ra;
                        uint32_t(&nxt)[32] = (c & 1) ? ra : rb;
                        tmem_ld_wait();
                        if (c + 1 < n_ch) tmem_ld32(taddr + 32 * (c + 1), nxt);
                        const int c0 = half * kFeatHalf0 + 32 * c;
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float e = ex2_approx(fmaf(__uint_as_float(cur[j]), kLog2e, shift2));
                            v[j] = !IS_Q ? e : (c0 + j < kFeat) ? fmaf(e, ratio, add) : 0.0f;
                        }
                        if (IS_Q) {
                            if (lane == 0) bulk_wait_read<0>();
                            __syncwarp();
#pragma unroll
                            for (int k4 = 0; k4 < 8; ++k4)
                                st_shared_v4(stage_buf + lane * 128 + ((k4 ^ (lane & 7)) << 4), v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]);
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) { tma_store_3d(&map_c, stage_buf, c0, m0 + q * 32, z); bulk_commit(); }
                        } else if (ok) {
                            float* dst = P.kt + ((int64_t)z * kFeatPad + c0) * P.Fp + f;
                            const uint32_t ld = (uint32_t)P.Fp;                  // 32-bit row offsets: one address instruction per store
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c0 + j < kFeat) dst[(uint32_t)j * ld] = v[j];
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(s32(acc_empty + slot_h[half]));      // this half's slot is free for the MMAs of the next tile
            }
        }
        if (IS_Q) {
            if (lane == 0) bulk_wait_all<0>();
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace tc
}  // namespace ddsp
