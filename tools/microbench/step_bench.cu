// Cost of the render loop's RK4 step in isolation (FMAD contract): every thread carries one ray and takes vacuum steps
// in unchecked bursts, like the burst path of render_kernel.  Reports SMSP cycles per warp-step for several unroll
// factors and residency levels, to separate the cost of the step itself from the loop bookkeeping around it.
//   nvcc -O3 -fmad=true -DRRT_FMAD=1 -gencode arch=compute_100a,code=sm_100a -I../../include -o step_bench step_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include "rrt_device.cuh"

using rrt::Consts;
using rrt::V3;
using rrt::mk;

template <int K, int MINB>
__global__ void __launch_bounds__(32, MINB) k_steps(const __grid_constant__ Consts C, int bursts, float* out) {
    const int t = blockIdx.x * 32 + threadIdx.x;
    V3 p = mk(0.0f + 1e-3f * (t & 255), 10.0f, -240.0f);
    V3 v = mk(1e-4f * (t & 63), 0.17f, 0.98f);
    float r2 = rrt::norm2_loop(p), r = rrt::sqrt_rn_fast(r2);
    float mn = r, mx = r;
#pragma unroll 1
    for (int b = 0; b < bursts; ++b) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float rm = rrt::rk4_step_fast<true>(C, p, v, C.h[0], C.hh[0], C.h6[0], r2, r);
            r2 = rrt::norm2_loop(p);
            r = rrt::sqrt_rn_fast(r2);
            mn = fminf(mn, fminf(rm, r));
            mx = fmaxf(mx, r);
        }
    }
    if (mn + mx + p.x + v.y == 123.456f) out[0] = mn;
}

static Consts consts() {
    Consts C; memset(&C, 0, sizeof(C));
    C.horizon_r = 2.02f; C.acc_rmin = 1.0f; C.radial_k = -3.0f; C.drag_k = 2.0f * 0.99f * 2.0f; C.spin_a = 0.99f; C.event_horizon = 2.0f;
    C.h[0] = 0.3f; C.hh[0] = 0.15f; C.h6[0] = 0.3f / 6.0f;
    return C;
}

template <int K, int MINB>
void run(int warps_per_smsp) {
    float* out; cudaMalloc(&out, 4);
    const Consts C = consts();
    const int blocks = 148 * 4 * warps_per_smsp, steps = 1600;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_steps<K, MINB>);
    int per_sm = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_steps<K, MINB>, 32, 0);
    if (per_sm < 4 * warps_per_smsp) { printf("K=%d minb=%d regs=%d: only %d warps/SM resident, skip %d/SMSP\n", K, MINB, fa.numRegs, per_sm, warps_per_smsp); return; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_steps<K, MINB><<<blocks, 32>>>(C, 16, out);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k_steps<K, MINB><<<blocks, 32>>>(C, steps / K, out); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warp_steps_per_smsp = (double)warps_per_smsp * (steps / K) * K;
    printf("K=%d minb=%2d regs=%3d warps/SMSP=%d  %8.3f ms  %.1f cycles per warp-step per SMSP (nominal %d MHz) -> %.3e steps/s\n", K, MINB,
           fa.numRegs, warps_per_smsp, best, best * 1e-3 * clk * 1e3 / warp_steps_per_smsp, clk / 1000,
           (double)blocks * 32.0 * (steps / K) * K / (best * 1e-3));
}

int main() {
    run<1, 1>(4); run<1, 1>(6); run<1, 1>(8);
    run<2, 1>(4); run<2, 1>(6); run<2, 1>(8);
    run<4, 1>(2); run<4, 1>(3); run<4, 1>(4); run<4, 1>(6); run<4, 1>(8);
    run<8, 1>(4); run<8, 1>(6); run<8, 1>(8);
    run<4, 24>(6); run<4, 32>(8); run<4, 40>(10); run<4, 48>(12);
    run<1, 24>(6); run<1, 32>(8); run<1, 48>(12);
    return 0;
}
