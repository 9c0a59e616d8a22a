// ref_cuda_planes_prelude.h -- instrumentation of the reference's OWN CUDA kernel.  TEST INFRASTRUCTURE ONLY.
//
// oracle/Makefile force-includes this file (nvcc -include) in front of /root/reference/src/raymarcher.cu, which is
// compiled UNMODIFIED from where it lies with exactly the flags of `make ref_cuda` (nvcc defaults, -fmad=true).
// Nothing of the kernel is restated: the reference headers are pulled in first (their include guards make the
// kernel's own #includes no-ops) and then two call sites inside raymarch_kernel are re-pointed by macro at wrappers
// that ALSO write the kernel's internal locals -- the quantities BASELINE.json's north_star grades -- to side planes:
//     final_hdr (src/raymarcher.cu:148-150), vel (:129 normalises it), hit_horizon (:38), transmittance, intensity_*
//     and the number of integrate_rk4 calls (:64).
// The wrappers only read those locals by the names the reference gives them; the arithmetic is the reference's.
// tests/test_gpu_refcuda_planes.py first checks that the uchar4 frame of this build equals the frame of the
// un-instrumented libref_cuda.so byte for byte, i.e. that observing the locals did not change nvcc's FMA fusion.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "config.h"
#include "math_utils.h"
#include "densities.h"
#include "geodesics.h"
#include "integrators.h"
#include "camera_effects/post_processing.h"
#include "raymarcher.h"

#ifndef RRT_PLANES_TAG
#error "define RRT_PLANES_TAG (a000 / a099)"
#endif
#define RRT_CAT2(a, b) a##b
#define RRT_CAT(a, b) RRT_CAT2(a, b)

struct RrtRefPlanes {
    float4* hdr;     // final_hdr.xyz at the store (effects must be off for the graded value), transmittance
    float4* vel;     // vel at loop exit, un-normalised
    float4* pos;     // p at loop exit
    float4* emis;    // intensity_r/g/b
    uint8_t* hit;    // hit_horizon
    int32_t* steps;  // integrate_rk4 calls; the harness zeroes it before the launch
};
__device__ RrtRefPlanes RRT_CAT(g_rrt_planes_, RRT_PLANES_TAG);

static __device__ __forceinline__ void rrt_count_step(int x, int y, int width) {
    int32_t* s = RRT_CAT(g_rrt_planes_, RRT_PLANES_TAG).steps;
    if (s) s[(size_t)y * width + x] += 1;
}
static __device__ __forceinline__ uchar4 rrt_capture(uchar4 px, int x, int y, int width, float3 hdr, float3 vel, float3 p,
                                                     bool hit, float T, float ir, float ig, float ib) {
    const RrtRefPlanes& P = RRT_CAT(g_rrt_planes_, RRT_PLANES_TAG);
    const size_t i = (size_t)y * width + x;
    if (P.hdr) P.hdr[i] = make_float4(hdr.x, hdr.y, hdr.z, T);
    if (P.vel) P.vel[i] = make_float4(vel.x, vel.y, vel.z, 0.f);
    if (P.pos) P.pos[i] = make_float4(p.x, p.y, p.z, 0.f);
    if (P.emis) P.emis[i] = make_float4(ir, ig, ib, 0.f);
    if (P.hit) P.hit[i] = hit ? 1 : 0;
    return px;
}
extern "C" int RRT_CAT(refcudap_set_planes_, RRT_PLANES_TAG)(const RrtRefPlanes* host) {
    return cudaMemcpyToSymbol(RRT_CAT(g_rrt_planes_, RRT_PLANES_TAG), host, sizeof(RrtRefPlanes)) == cudaSuccess ? 0 : -1;
}

// function-level probes of the reference headers as nvcc compiles them (same flags, same SPIN_A)
extern "C" __global__ void RRT_CAT(refcudap_k_probe_, RRT_PLANES_TAG)(int what, int n, const float* a, const float* b, float time, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 q = make_float3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
    if (what == 0) out[i] = getAccretionDensity(q, time);
    else if (what == 1) out[i] = getDustCloudDensity(q, time);
    else if (what == 2) out[i] = calculateRedshiftFactor(q, make_float3(b[3 * i], b[3 * i + 1], b[3 * i + 2]));
    else if (what == 3) out[i] = getDiskTemperature(a[i]);
    else if (what == 4) out[i] = noise3D(q);
    else if (what == 5) out[i] = fbm(q, 5);
}
extern "C" int RRT_CAT(refcudap_probe_, RRT_PLANES_TAG)(int what, int n, const float* d_a, const float* d_b, float time, float* d_out) {
    if (n > 0) RRT_CAT(refcudap_k_probe_, RRT_PLANES_TAG)<<<(n + 127) / 128, 128>>>(what, n, d_a, d_b, time, d_out);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

// ---- the two re-pointed call sites inside raymarch_kernel (names on the right are the kernel's own locals) ----
#define integrate_rk4(P_, V_, H_) (rrt_count_step(x, y, width), integrate_rk4(P_, V_, H_))
#define make_uchar4(R_, G_, B_, A_) \
    rrt_capture(make_uchar4(R_, G_, B_, A_), x, y, width, final_hdr, vel, p, hit_horizon, transmittance, intensity_r, intensity_g, intensity_b)
