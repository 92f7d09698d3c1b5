// Microbenchmark: how many SM cycles does one tcgen05.mma kind::tf32 (M = 128, N, K = 8, both operands in shared memory)
// take as a function of N?  One CTA per SM; one thread issues `reps` MMAs back to back (alternating two accumulators,
// operands: zeroed 128-byte-swizzled tiles), commits, and waits for the mbarrier; clock64 around issue + completion.
// The 3xTF32 GEMMs of the control network issue N = 256 (plain layers), 96 / 80 (attention context / output), 128 + 144
// (feature maps): if narrow MMAs cost more than N / 2 cycles, the attention GEMMs are bound by instruction width.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../ddsp-svc-official_b200/csrc -o umma_issue umma_issue.cu && ./umma_issue
#include <cstdio>
#include <cuda_runtime.h>
#include "gemm_tc.cuh"
using namespace ddsp::tc;

__global__ void __launch_bounds__(128, 1) k(long long* out, int n, int reps) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.0f;
    if (threadIdx.x == 0) { mbar_init(s32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(s32(&tmem_slot));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        const uint64_t a = umma_desc_sw128(s32(smem)), b = umma_desc_sw128(s32(smem + 128 * 128));
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_tf32(tmem_base + (r & 1) * 256, a + 2 * kk, b + 2 * kk, idesc, 1);
        }
        const long long t1 = clock64();
        umma_commit(s32(&bar));
        mbar_wait(s32(&bar), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem_base);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 16);
    const int smem = (128 + 256) * 128 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 2048;
    for (int grid : {1, 148})
        for (int n : {16, 32, 64, 80, 96, 112, 128, 144, 192, 224, 256}) {
            k<<<grid, 128, smem>>>(out, n, reps);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            printf("grid %3d  N = %3d: %7.1f cycles per MMA (issue alone %6.1f), N/2 = %3d\n", grid, n, (double)out[1] / (4.0 * reps),
                   (double)out[0] / (4.0 * reps), n / 2);
        }
    return 0;
}
