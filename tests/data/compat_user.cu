// A "user of the reference headers": includes the drop-in headers by the reference's names and calls the
// functions north_star lists.  Built by tests/test_compat_headers.py.
#include "raymarcher.h"
#include "config.h"
#include "geodesics.h"
#include "integrators.h"
#include "densities.h"

__global__ void k_user(int n, const float3* q, const float3* v, float3* acc, float3* p_rk4, float3* v_rk4, float3* p_eu,
                       float3* v_eu, float* g, float* dens) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    acc[i] = getGeodesicAcc(q[i], v[i]);
    float3 p = q[i], w = v[i];
    integrate_rk4(p, w, STEP_SIZE_M);
    p_rk4[i] = p; v_rk4[i] = w;
    p = q[i]; w = v[i];
    integrate_euler(p, w, STEP_SIZE_M);
    p_eu[i] = p; v_eu[i] = w;
    g[i] = calculateRedshiftFactor(q[i], v[i]);
    dens[i] = getAccretionDensity(q[i], 1.0f) + getDustCloudDensity(q[i], 1.0f) + getDiskTemperature(12.0f);
}

extern "C" int compat_user_run(int n, const float* q, const float* v, float* acc, float* p_rk4, float* v_rk4, float* p_eu,
                               float* v_eu, float* g, float* dens) {
    float3 *dq, *dv, *o[5];
    float *dg, *dd;
    size_t b = (size_t)n * sizeof(float3);
    cudaMalloc(&dq, b); cudaMalloc(&dv, b);
    for (auto& x : o) cudaMalloc(&x, b);
    cudaMalloc(&dg, n * sizeof(float)); cudaMalloc(&dd, n * sizeof(float));
    cudaMemcpy(dq, q, b, cudaMemcpyHostToDevice); cudaMemcpy(dv, v, b, cudaMemcpyHostToDevice);
    k_user<<<(n + 127) / 128, 128>>>(n, dq, dv, o[0], o[1], o[2], o[3], o[4], dg, dd);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    float* outs[5] = {acc, p_rk4, v_rk4, p_eu, v_eu};
    for (int i = 0; i < 5; ++i) cudaMemcpy(outs[i], o[i], b, cudaMemcpyDeviceToHost);
    cudaMemcpy(g, dg, n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(dens, dd, n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(dq); cudaFree(dv); for (auto& x : o) cudaFree(x); cudaFree(dg); cudaFree(dd);
    return 0;
}

// host side of the reference: main.cpp's call site compiles and links against the shim unchanged
extern "C" void compat_user_launch(uchar4* d_out, int w, int h, float t, const float* cam12, unsigned long long tex) {
    CameraState cam;
    cam.pos = make_float3(cam12[0], cam12[1], cam12[2]);
    cam.forward = make_float3(cam12[3], cam12[4], cam12[5]);
    cam.right = make_float3(cam12[6], cam12[7], cam12[8]);
    cam.up = make_float3(cam12[9], cam12[10], cam12[11]);
    CameraEffects fx;  // reference defaults
    launch_raymarch(d_out, w, h, t, cam, (cudaTextureObject_t)tex, fx);
}
