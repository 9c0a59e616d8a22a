// ref_cuda_harness.cu -- headless driver for the reference's OWN CUDA raymarcher, as a reported
// baseline and a second parity witness (uchar4 frames).  TEST INFRASTRUCTURE ONLY (see oracle_abi.h).
//
// oracle/Makefile compiles /root/reference/src/raymarcher.cu, unmodified, twice (SPIN_A = 0.0f and
// 0.99f through a generated config.h found first on the include path; -Dlaunch_raymarch=..._aXXX
// only renames the symbols so both objects can live in one library) and links them with this file.
// The skybox texture is created exactly like src/main.cpp:246-263.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "raymarcher.h"  // the reference's header: CameraState, CameraEffects

void launch_raymarch_a000(uchar4*, int, int, float, CameraState, cudaTextureObject_t, CameraEffects);
void launch_raymarch_a099(uchar4*, int, int, float, CameraState, cudaTextureObject_t, CameraEffects);

extern "C" int refcuda_render(int spin_is_099, int w, int h, float time, const float cam12[12], const int32_t fx_i[4],
                              const float fx_f[5], const uint8_t* sky_rgba, int sky_w, int sky_h, uint8_t* out_rgba,
                              int reps, float* best_ms, float* mean_ms) {
    CameraState cam;
    memcpy(&cam, cam12, sizeof(cam));
    CameraEffects fx;
    fx.useBloom = fx_i[0] != 0; fx.useVignette = fx_i[1] != 0; fx.useChromaticAberration = fx_i[2] != 0; fx.useLensDistortion = fx_i[3] != 0;
    fx.bloomThreshold = fx_f[0]; fx.bloomIntensity = fx_f[1]; fx.vignetteIntensity = fx_f[2]; fx.caAmount = fx_f[3]; fx.distortionAmount = fx_f[4];

    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
    cudaChannelFormatDesc channelDesc = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);
    if (cudaMallocArray(&arr, &channelDesc, sky_w, sky_h) != cudaSuccess) return -1;
    cudaMemcpy2DToArray(arr, 0, 0, sky_rgba, (size_t)sky_w * 4, (size_t)sky_w * 4, sky_h, cudaMemcpyHostToDevice);
    cudaResourceDesc resDesc; memset(&resDesc, 0, sizeof(resDesc));
    resDesc.resType = cudaResourceTypeArray; resDesc.res.array.array = arr;
    cudaTextureDesc texDesc; memset(&texDesc, 0, sizeof(texDesc));
    texDesc.addressMode[0] = cudaAddressModeWrap; texDesc.addressMode[1] = cudaAddressModeClamp;
    texDesc.filterMode = cudaFilterModeLinear; texDesc.readMode = cudaReadModeNormalizedFloat; texDesc.normalizedCoords = 1;
    if (cudaCreateTextureObject(&tex, &resDesc, &texDesc, NULL) != cudaSuccess) { cudaFreeArray(arr); return -2; }

    uchar4* d_out = nullptr;
    if (cudaMalloc(&d_out, (size_t)w * h * 4) != cudaSuccess) return -3;
    auto launch = spin_is_099 ? launch_raymarch_a099 : launch_raymarch_a000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f, sum = 0.f;
    launch(d_out, w, h, time, cam, tex, fx);  // warm-up
    cudaDeviceSynchronize();
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch(d_out, w, h, time, cam, tex, fx);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        sum += ms;
    }
    int rc = cudaGetLastError() == cudaSuccess ? 0 : -4;
    if (out_rgba) cudaMemcpy(out_rgba, d_out, (size_t)w * h * 4, cudaMemcpyDeviceToHost);
    if (best_ms) *best_ms = best;
    if (mean_ms) *mean_ms = reps > 0 ? sum / reps : 0.f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out); cudaDestroyTextureObject(tex); cudaFreeArray(arr);
    return rc;
}
