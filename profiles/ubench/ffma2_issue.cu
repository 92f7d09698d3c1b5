// Microbenchmark: does a packed fp32x2 instruction (FFMA2) occupy ONE issue slot and two FMA-pipe cycles, or TWO issue
// slots?  Each warp runs a loop of independent FFMA2 chains, optionally interleaved with independent integer ALU work
// (LOP3/IADD3, other pipe) or scalar FFMA.  Reported: SM cycles per loop iteration per scheduler (4 warps/scheduler).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_issue ffma2_issue.cu && ./ffma2_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float2* out, int iters, float s) {
    float2 a[8];
    float f[8];
    unsigned u[8], v[8], w[4];
    float g[8], f2[8];
    float2 b2[8];
    for (int i = 0; i < 4; ++i) w[i] = (unsigned)(s * 1000.f) + i * threadIdx.x;
    float2 q2[8];                                   // opaque register pairs (loaded, so that nothing folds into immediates)
    for (int i = 0; i < 8; ++i) q2[i] = out[(threadIdx.x * 8 + i) & 1023];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = make_float2(threadIdx.x + i, i); f[i] = threadIdx.x * 0.5f + i; u[i] = threadIdx.x + i; v[i] = 3 * threadIdx.x + i; g[i] = f[i] * 0.3f; f2[i] = f[i] * 0.7f; b2[i] = make_float2(1.0f + 1e-6f * i * s, 1.0f - 1e-6f * threadIdx.x * s); }
    const float2 m = make_float2(s, s * 0.5f), c = make_float2(0.25f, 0.125f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0 || MODE == 1 || MODE == 2) a[i] = __ffma2_rn(a[i], m, c);               // 8 FFMA2
                if (MODE == 1) u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e3779b9u;                             // + ALU work (SHF, LOP3, IADD3)
                if (MODE == 2 || MODE == 3) f[i] = fmaf(f[i], s, 0.25f);                               // + 8 scalar FFMA
                if (MODE == 3) f[i] = fmaf(f[i], s, 0.5f);                                             // 16 scalar FFMA, no FFMA2
                if (MODE == 4) u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e3779b9u;                             // ALU only
                if (MODE == 10) { a[i] = __ffma2_rn(a[i], m, c); u[i] = (u[i] & v[i]) ^ w[i & 3]; }           // 8 FFMA2 + 8 LOP3
                if (MODE == 11) { u[i] = (u[i] & v[i]) ^ w[i & 3]; }                                            // 8 LOP3
                if (MODE == 12) { a[i] = __ffma2_rn(a[i], m, c); u[i] = (u[i] & v[i]) ^ w[i & 3]; v[i] = (v[i] | w[(i + 1) & 3]) ^ u[i]; }  // 8 FFMA2 + 16 LOP3
                if (MODE == 13) { u[i] = (u[i] & v[i]) ^ w[i & 3]; v[i] = (v[i] | w[(i + 1) & 3]) ^ u[i]; }    // 16 LOP3
                if (MODE == 14) { f[i] = fmaf(f[i], s, 0.25f); g[i] = fmaf(g[i], s, 0.5f); u[i] = (u[i] & v[i]) ^ w[i & 3]; v[i] = (v[i] | w[(i + 1) & 3]) ^ u[i]; }  // 16 FFMA + 16 LOP3
                if (MODE == 15) { f[i] = fmaf(f[i], g[i], g[(i + 3) & 7]); f2[i] = fmaf(f2[i], g[(i + 1) & 7], g[(i + 5) & 7]); }   // 16 FFMA, three register operands
                if (MODE == 16) { a[i] = __ffma2_rn(a[i], b2[i], b2[(i + 3) & 7]); }                                   // 8 FFMA2, three register-pair operands
                if (MODE == 17) { a[i] = __ffma2_rn(a[i], make_float2(1.0001f, 1.0001f), make_float2(0.25f, 0.25f)); }   // 8 FFMA2, broadcast immediates
                if (MODE == 20) { a[i] = __fadd2_rn(a[i], q2[i]); }                                               // 8 FADD2, two register pairs
                if (MODE == 21) { a[i] = __ffma2_rn(a[i], q2[i], make_float2(0.25f, 0.25f)); }                    // 8 FFMA2, two register pairs + immediate
                if (MODE == 22) { a[i] = __ffma2_rn(a[i], q2[i], q2[(i + 3) & 7]); }                              // 8 FFMA2, three register pairs
                if (MODE == 23) { f[i] = fmaf(f[i], q2[i].x, q2[(i + 3) & 7].y); g[i] = fmaf(g[i], q2[i].y, q2[(i + 5) & 7].x); }   // 16 FFMA, three registers
                if (MODE == 24) { a[i] = __fmul2_rn(a[i], q2[i]); }                                               // 8 FMUL2, two register pairs
                if (MODE == 18) { f[i] = fmaf(f[i], g[i], f2[i]); g[i] = fmaf(g[i], f2[i], f[i]); }          // 16 FFMA, three live register operands each
                if (MODE == 19) { a[i] = __ffma2_rn(a[i], b2[i], a[(i + 1) & 7]); }                               // 8 FFMA2, three live register-pair operands
                if (MODE == 5) { a[i] = __ffma2_rn(a[i], m, c); u[i] = u[i] + w[i & 3]; v[i] = v[i] + w[(i + 1) & 3]; }   // 8 FFMA2 + 16 IADD
                if (MODE == 6) { u[i] = u[i] + w[i & 3]; v[i] = v[i] + w[(i + 1) & 3]; }                        // 16 IADD
                if (MODE == 7) { f[i] = fmaf(f[i], s, 0.25f); g[i] = fmaf(g[i], s, 0.5f); u[i] = u[i] + w[i & 3]; v[i] = v[i] + w[(i + 1) & 3]; }   // 16 FFMA + 16 IADD
                if (MODE == 8) { a[i] = __ffma2_rn(a[i], m, c); u[i] = u[i] + w[i & 3]; }                     // 8 FFMA2 + 8 IADD
                if (MODE == 9) { a[i] = __ffma2_rn(a[i], m, c); f[i] = fmaf(f[i], s, 0.25f); u[i] = u[i] + w[i & 3]; v[i] = v[i] + w[(i + 1) & 3]; }  // 8 FFMA2 + 8 FFMA + 16 IADD
            }
        }
    }
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc.x += a[i].x + f[i]; acc.y += a[i].y + (float)u[i] + (float)v[i] + g[i] + f2[i] + b2[i].x; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int per_iter_note) {
    float2* out;
    cudaMalloc(&out, 148 * 512 * sizeof(float2));
    cudaMemset(out, 0, 148 * 512 * sizeof(float2));
    const int iters = 20000;
    k<MODE><<<148, 512>>>(out, 100, 1.0001f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, 512>>>(out, iters, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // per scheduler: 4 warps x iters x 32 "slots" (4 x 8) per iteration
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-40s %8.3f ms  %6.2f cycles per (warp, group of 8 statements) -> x4 warps = %6.2f cycles/scheduler per group of 8 [%d instr]\n",
           name, ms, cycles / iters / 4 / 4, cycles / iters / 4, per_iter_note);
    cudaFree(out);
}

int main() {
    run<0>("8 FFMA2", 8);
    run<1>("8 FFMA2 + 8x(SHF,LOP3,IADD3)", 32);
    run<2>("8 FFMA2 + 8 FFMA", 16);
    run<3>("16 FFMA", 16);
    run<4>("8x(SHF,LOP3,IADD3)", 24);
    run<10>("8 FFMA2 + 8 LOP3", 16);
    run<11>("8 LOP3", 8);
    run<12>("8 FFMA2 + 16 LOP3", 24);
    run<13>("16 LOP3", 16);
    run<14>("16 FFMA + 16 LOP3", 32);
    run<20>("8 FADD2 (2 register pairs)", 8);
    run<21>("8 FFMA2 (2 register pairs + immediate)", 8);
    run<22>("8 FFMA2 (3 register pairs)", 8);
    run<23>("16 FFMA (3 registers)", 16);
    run<24>("8 FMUL2 (2 register pairs)", 8);
    run<18>("16 FFMA (3 live registers, rotating roles)", 16);
    run<19>("8 FFMA2 (3 live register pairs)", 8);
    run<15>("16 FFMA (3 registers)", 16);
    run<16>("8 FFMA2 (3 register pairs)", 8);
    run<17>("8 FFMA2 (broadcast immediates)", 8);
    run<5>("8 FFMA2 + 16 IADD", 24);
    run<6>("16 IADD", 16);
    run<7>("16 FFMA + 16 IADD", 32);
    run<8>("8 FFMA2 + 8 IADD", 16);
    run<9>("8 FFMA2 + 8 FFMA + 16 IADD", 32);
    return 0;
}
