// ref_cuda_planes.cu -- headless driver for the INSTRUMENTED build of the reference's own CUDA kernel
// (see ref_cuda_planes_prelude.h: src/raymarcher.cu unmodified, two call sites re-pointed by macro so that the
// kernel's internal locals are also written to planes).  TEST INFRASTRUCTURE ONLY.
// Float-precision witness for the FMAD contract: final_hdr, vel, hit_horizon and the step count of the reference's
// own GPU arithmetic (SURVEY.md 8c golden item 3).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "raymarcher.h"  // the reference's header: CameraState, CameraEffects

struct RrtRefPlanes { float4* hdr; float4* vel; float4* pos; float4* emis; uint8_t* hit; int32_t* steps; };
void launch_raymarch_p000(uchar4*, int, int, float, CameraState, cudaTextureObject_t, CameraEffects);
void launch_raymarch_p099(uchar4*, int, int, float, CameraState, cudaTextureObject_t, CameraEffects);
extern "C" int refcudap_set_planes_a000(const RrtRefPlanes*);
extern "C" int refcudap_set_planes_a099(const RrtRefPlanes*);
extern "C" int refcudap_probe_a000(int, int, const float*, const float*, float, float*);
extern "C" int refcudap_probe_a099(int, int, const float*, const float*, float, float*);

// host pointers out_*: [h*w*4] float / [h*w] u8 / [h*w] i32, pixel (x, y) at y*w + x (NOT row-flipped); any may be NULL.
extern "C" int refcudap_render(int spin_is_099, int w, int h, float time, const float cam12[12], const int32_t fx_i[4],
                               const float fx_f[5], const uint8_t* sky_rgba, int sky_w, int sky_h, uint8_t* out_rgba,
                               float* out_hdr, float* out_vel, float* out_pos, float* out_emis, uint8_t* out_hit, int32_t* out_steps) {
    CameraState cam;
    memcpy(&cam, cam12, sizeof(cam));
    CameraEffects fx;
    fx.useBloom = fx_i[0] != 0; fx.useVignette = fx_i[1] != 0; fx.useChromaticAberration = fx_i[2] != 0; fx.useLensDistortion = fx_i[3] != 0;
    fx.bloomThreshold = fx_f[0]; fx.bloomIntensity = fx_f[1]; fx.vignetteIntensity = fx_f[2]; fx.caAmount = fx_f[3]; fx.distortionAmount = fx_f[4];

    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
    cudaChannelFormatDesc channelDesc = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);   // src/main.cpp:246-263
    if (cudaMallocArray(&arr, &channelDesc, sky_w, sky_h) != cudaSuccess) return -1;
    cudaMemcpy2DToArray(arr, 0, 0, sky_rgba, (size_t)sky_w * 4, (size_t)sky_w * 4, sky_h, cudaMemcpyHostToDevice);
    cudaResourceDesc resDesc; memset(&resDesc, 0, sizeof(resDesc));
    resDesc.resType = cudaResourceTypeArray; resDesc.res.array.array = arr;
    cudaTextureDesc texDesc; memset(&texDesc, 0, sizeof(texDesc));
    texDesc.addressMode[0] = cudaAddressModeWrap; texDesc.addressMode[1] = cudaAddressModeClamp;
    texDesc.filterMode = cudaFilterModeLinear; texDesc.readMode = cudaReadModeNormalizedFloat; texDesc.normalizedCoords = 1;
    if (cudaCreateTextureObject(&tex, &resDesc, &texDesc, NULL) != cudaSuccess) { cudaFreeArray(arr); return -2; }

    const size_t n = (size_t)w * h;
    uchar4* d_out = nullptr;
    RrtRefPlanes P; memset(&P, 0, sizeof(P));
    int rc = 0;
    if (cudaMalloc(&d_out, n * 4) != cudaSuccess) rc = -3;
    if (!rc && out_hdr && cudaMalloc(&P.hdr, n * 16) != cudaSuccess) rc = -3;
    if (!rc && out_vel && cudaMalloc(&P.vel, n * 16) != cudaSuccess) rc = -3;
    if (!rc && out_pos && cudaMalloc(&P.pos, n * 16) != cudaSuccess) rc = -3;
    if (!rc && out_emis && cudaMalloc(&P.emis, n * 16) != cudaSuccess) rc = -3;
    if (!rc && out_hit && cudaMalloc(&P.hit, n) != cudaSuccess) rc = -3;
    if (!rc && out_steps && (cudaMalloc(&P.steps, n * 4) != cudaSuccess || cudaMemset(P.steps, 0, n * 4) != cudaSuccess)) rc = -3;
    if (!rc) {
        if ((spin_is_099 ? refcudap_set_planes_a099 : refcudap_set_planes_a000)(&P) != 0) rc = -5;
    }
    if (!rc) {
        (spin_is_099 ? launch_raymarch_p099 : launch_raymarch_p000)(d_out, w, h, time, cam, tex, fx);
        if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = -4;
    }
    if (!rc) {
        if (out_rgba) cudaMemcpy(out_rgba, d_out, n * 4, cudaMemcpyDeviceToHost);
        if (out_hdr) cudaMemcpy(out_hdr, P.hdr, n * 16, cudaMemcpyDeviceToHost);
        if (out_vel) cudaMemcpy(out_vel, P.vel, n * 16, cudaMemcpyDeviceToHost);
        if (out_pos) cudaMemcpy(out_pos, P.pos, n * 16, cudaMemcpyDeviceToHost);
        if (out_emis) cudaMemcpy(out_emis, P.emis, n * 16, cudaMemcpyDeviceToHost);
        if (out_hit) cudaMemcpy(out_hit, P.hit, n, cudaMemcpyDeviceToHost);
        if (out_steps) cudaMemcpy(out_steps, P.steps, n * 4, cudaMemcpyDeviceToHost);
    }
    RrtRefPlanes zero; memset(&zero, 0, sizeof(zero));
    (spin_is_099 ? refcudap_set_planes_a099 : refcudap_set_planes_a000)(&zero);
    cudaFree(d_out); cudaFree(P.hdr); cudaFree(P.vel); cudaFree(P.pos); cudaFree(P.emis); cudaFree(P.hit); cudaFree(P.steps);
    cudaDestroyTextureObject(tex); cudaFreeArray(arr);
    return rc;
}

// what: 0 getAccretionDensity(a, time)  1 getDustCloudDensity(a, time)  2 calculateRedshiftFactor(a, b)  3 getDiskTemperature(a[i])
//       4 noise3D(a)  5 fbm(a, 5);  a is [n*3] (or [n] for 3), b [n*3] or NULL; host pointers.
extern "C" int refcudap_probe(int spin_is_099, int what, int n, const float* a, const float* b, float time, float* out) {
    const size_t na = (size_t)n * (what == 3 ? 1 : 3);
    float *d_a = nullptr, *d_b = nullptr, *d_o = nullptr;
    int rc = 0;
    if (cudaMalloc(&d_a, na * 4 + 4) != cudaSuccess || cudaMalloc(&d_o, (size_t)n * 4 + 4) != cudaSuccess) rc = -3;
    if (!rc && b && cudaMalloc(&d_b, (size_t)n * 12 + 4) != cudaSuccess) rc = -3;
    if (!rc) {
        cudaMemcpy(d_a, a, na * 4, cudaMemcpyHostToDevice);
        if (b) cudaMemcpy(d_b, b, (size_t)n * 12, cudaMemcpyHostToDevice);
        rc = (spin_is_099 ? refcudap_probe_a099 : refcudap_probe_a000)(what, n, d_a, d_b, time, d_o);
        if (!rc) cudaMemcpy(out, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_o);
    return rc;
}
