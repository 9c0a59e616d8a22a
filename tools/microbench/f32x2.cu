// Microbenchmark: does the packed FP32 path of sm_100 (add/mul/fma .f32x2 -> FADD2/FMUL2/FFMA2) deliver the
// same FLOP rate as scalar FFMA with half the issue slots?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x; }

#define NCH 8
template <int MODE, int NALU>
__global__ void __launch_bounds__(256) k(int iters, float seed, float* sink) {
    float a[NCH]; u64 p[NCH]; int ia[4] = {1, 2, 3, 4};
    const float m = 0.999f + seed * 1e-9f, c = 1e-3f + seed;
    const u64 M = pk(m, m), C = pk(c, c);
#pragma unroll
    for (int i = 0; i < NCH; ++i) { a[i] = seed + i; p[i] = pk(seed + i, seed - i); }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) {            // scalar FFMA: 8 per u
#pragma unroll
                for (int i = 0; i < NCH; ++i) a[i] = __fmaf_rn(a[i], m, c);
            } else if (MODE == 1) {     // FFMA2: 4 per u (same flops as 8 scalar)
#pragma unroll
                for (int i = 0; i < NCH / 2; ++i) p[i] = fma2(p[i], M, C);
            } else if (MODE == 2) {     // FMUL2 + FADD2: 2+2 per u (half the flops of MODE 1)
#pragma unroll
                for (int i = 0; i < NCH / 4; ++i) { p[i] = mul2(p[i], M); p[i + 2] = add2(p[i + 2], C); }
            } else if (MODE == 3) {     // scalar FMUL + FADD, 4+4 per u
#pragma unroll
                for (int i = 0; i < NCH / 2; ++i) { a[i] = __fmul_rn(a[i], m); a[i + 4] = __fadd_rn(a[i + 4], c); }
            }
#pragma unroll
            for (int j = 0; j < NALU; ++j) ia[j & 3] = (ia[j & 3] ^ (ia[(j + 1) & 3] + it)) ;   // LOP3/IADD on the ALU pipe
        }
    }
    float s = 0; for (int i = 0; i < NCH; ++i) s += a[i] + lo(p[i]);
    if (s == 123456.789f || ia[0] + ia[1] + ia[2] + ia[3] == 0x7fffffff) sink[0] = s;
}

template <int MODE, int NALU>
void run(const char* name, double flop_per_u) {
    int dev_sms = 148; float* sink; cudaMalloc(&sink, 4);
    const int blocks = dev_sms * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, NALU><<<blocks, threads>>>(iters / 8, 1.f, sink);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k<MODE, NALU><<<blocks, threads>>>(iters, 1.f, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double flop = (double)blocks * threads * iters * 8.0 * flop_per_u;
    double slots = (double)blocks * threads / 32.0 * iters * 8.0;   // warp-iterations of the unrolled body
    printf("%-34s %8.3f ms  %7.2f TFLOP/s   %.2f cycles per u-body per SMSP (1.965 GHz)\n", name, best, flop / (best * 1e-3) / 1e12,
           best * 1e-3 * 1.965e9 / (slots / (148.0 * 4.0)));
}
int main() {
    run<0, 0>("scalar FFMA x8", 16);
    run<1, 0>("FFMA2 x4", 16);
    run<2, 0>("FMUL2 x2 + FADD2 x2", 8);
    run<3, 0>("scalar FMUL x4 + FADD x4", 8);
    run<0, 4>("scalar FFMA x8 + 4 ALU", 16);
    run<1, 4>("FFMA2 x4 + 4 ALU", 16);
    run<0, 8>("scalar FFMA x8 + 8 ALU", 16);
    run<1, 8>("FFMA2 x4 + 8 ALU", 16);
    return 0;
}
