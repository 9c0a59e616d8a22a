// lone_warp.cu -- what does one extra operation per iteration cost a warp that runs ALONE on its SM sub-partition?
// One-warp CTAs, one per SM; each runs a dependent FFMA chain (64 FFMAs per iteration) for N iterations plus, per iteration:
//   mode 0 nothing            mode 1 divergent branch, one lane adds to a register
//   mode 2 one lane atomicAdd on shared memory        mode 3 one lane plain shared store + load
//   mode 4 every lane one 32-byte global store        mode 5 one lane atomicAdd on global memory
//   mode 6 __match_any_sync + __shfl_sync             mode 7 one lane shared atomicCAS (64-bit)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lone_warp lone_warp.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(32) k(int mode, int n, float seed, float* out, unsigned* g, uint4* pool) {
    __shared__ unsigned s[4];
    __shared__ unsigned long long s64;
    const int lane = threadIdx.x;
    if (lane == 0) { s[0] = 0; s64 = 0; }
    __syncwarp();
    float a = seed + lane * 1e-3f, b = 0.999f, c = 1e-3f;
    unsigned acc = 0;
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 64; ++u) a = __fmaf_rn(a, b, c);
        const bool want = a > -1.0f;   // always true, but the compiler cannot know
        if (mode == 1) { if (want && lane == (i & 31)) acc += i; }
        else if (mode == 2) { if (want && lane == (i & 31)) atomicAdd(&s[0], 1u); }
        else if (mode == 3) { if (want && lane == (i & 31)) { ((volatile unsigned*)s)[1] = i; acc += ((volatile unsigned*)s)[1]; } }
        else if (mode == 4) { if (want) { uint4* p = pool + 2ull * ((size_t)blockIdx.x * 65536 + (size_t)(i & 2047) * 32 + lane); p[0] = make_uint4(i, lane, 0, 0); p[1] = make_uint4(0, 0, 0, i); } }
        else if (mode == 5) { if (want && lane == (i & 31)) atomicAdd(g + blockIdx.x, 1u); }
        else if (mode == 6) { if (want) { const unsigned m = __match_any_sync(__activemask(), i & 1); acc += __shfl_sync(m, acc + lane, __ffs(m) - 1); } }
        else if (mode == 7) { if (want && lane == (i & 31)) { unsigned long long o = s64; atomicCAS(&s64, o, o + 1); } }
    }
    if (a == 123.456f || acc == 0xdeadbeef) out[0] = a + acc;
}
int main() {
    float* out; unsigned* g; uint4* pool;
    cudaMalloc(&out, 4); cudaMalloc(&g, 4096); cudaMemset(g, 0, 4096);
    cudaMalloc(&pool, 148ull * 65536 * 32);
    int dev_clock = 0; cudaDeviceGetAttribute(&dev_clock, cudaDevAttrClockRate, 0);
    const int n = 20000;
    for (int blocks : {148, 148 * 16}) {
        for (int mode = 0; mode < 8; ++mode) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k<<<blocks, 32>>>(mode, 100, 1.0f, out, g, pool);
            cudaEventRecord(e0);
            k<<<blocks, 32>>>(mode, n, 1.0f, out, g, pool);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("blocks %5d (%2d warps/SM) mode %d: %8.3f ms  %8.1f cycles per iteration (at %d MHz nominal)\n", blocks, blocks / 148, mode, ms,
                   ms * 1e-3 * dev_clock * 1e3 / n, dev_clock / 1000);
        }
    }
    return 0;
}
