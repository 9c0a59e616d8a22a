// Register-file operand delivery on B200: how many SMSP cycles does a scalar FP32 instruction cost as a function of
// how many distinct register sources it reads?  Each pattern is a loop of 64 independent instructions; the SASS of
// this very binary (cuobjdump -sass) says which registers ptxas picked, so the static bank model in
// tools/sass_bankcost.py can be compared with the measured cycles.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rf_banks rf_banks.cu && ./rf_banks
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void __launch_bounds__(128) k(int iters, const float* prm, float* sink, float cb0, float cb1) {
    float x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = prm[i] + threadIdx.x * 1e-6f; y[i] = prm[8 + i]; z[i] = prm[16 + i]; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = __fmaf_rn(x[i], y[i], z[i]);            // 3 distinct registers
                else if (MODE == 1) x[i] = __fmaf_rn(x[i], y[i], 1.5f);       // 2 registers + immediate
                else if (MODE == 2) x[i] = __fmaf_rn(x[i], cb0, y[i]);        // 2 registers + constant bank
                else if (MODE == 3) x[i] = __fmul_rn(x[i], y[i]);             // FMUL 2 registers
                else if (MODE == 4) x[i] = __fmaf_rn(x[i], x[i], y[i]);       // 2 distinct (one twice)
                else if (MODE == 5) x[i] = __fmaf_rn(x[i], cb0, cb1 > 0 ? 1.25f : 1.25f);  // 1 register
                else if (MODE == 6) x[i] = __fmaf_rn(x[i], y[0], z[0]);       // 3 registers, two shared by every instruction (.reuse)
                else if (MODE == 7) x[i] = __fmaf_rn(x[i], y[i], z[0]);       // 3 registers, one shared
                else if (MODE == 8) x[i] = __fadd_rn(x[i], y[i]);             // FADD 2 registers
                else if (MODE == 9) x[i] = __fmul_rn(x[i], x[(i + 1 + (u % 7)) % 8]);   // FMUL, every register pair: same-bank pairs unavoidable
                else if (MODE == 10) x[i] = __fmaf_rn(x[i], x[(i + 1 + (u % 7)) % 8], 0.75f);   // FFMA imm-form, same
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123456.789f) sink[0] = s;
}

template <int MODE>
void run(const char* name, const float* dprm, int warps_per_smsp) {
    float* sink; cudaMalloc(&sink, 4);
    const int blocks = 148 * warps_per_smsp, threads = 128, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(iters / 8, dprm, sink, 1.0001f, 0.5f);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(iters, dprm, sink, 1.0001f, 0.5f); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double warp_ops = (double)blocks * threads / 32.0 * iters * 64.0;
    printf("mode %d %-52s warps/SMSP %d  %8.3f ms   %.3f cycles per instr per SMSP (at %d MHz nominal)\n", MODE, name, warps_per_smsp, best,
           best * 1e-3 * (clk * 1e3) / (warp_ops / (148.0 * 4.0)), clk / 1000);
}
int main() {
    float h[32]; for (int i = 0; i < 32; ++i) h[i] = 1.0f + 1e-3f * i;
    float* d; cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w : {4, 6}) {
        run<0>("FFMA x=x*y+z  (3 distinct regs)", d, w);
        run<1>("FFMA x=x*y+imm (2 regs)", d, w);
        run<2>("FFMA x=x*c[]+y (2 regs + const bank)", d, w);
        run<3>("FMUL x=x*y (2 regs)", d, w);
        run<4>("FFMA x=x*x+y (2 distinct)", d, w);
        run<5>("FFMA x=x*c[]+imm (1 reg)", d, w);
        run<6>("FFMA x=x*Y+Z (3 regs, Y and Z shared)", d, w);
        run<7>("FFMA x=x*y+Z (3 regs, Z shared)", d, w);
        run<8>("FADD x=x+y (2 regs)", d, w);
        run<9>("FMUL x_i=x_i*x_j, all pairs (forced same-bank pairs)", d, w);
        run<10>("FFMA x_i=x_i*x_j+imm, all pairs", d, w);
    }
    return 0;
}
