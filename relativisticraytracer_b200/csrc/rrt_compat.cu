// rrt_compat.cu -- the reference-facing C++ shim: launch_raymarch with the reference's exact signature
// and mangling (_Z15launch_raymarchP6uchar4iif11CameraStatey13CameraEffects; include/raymarcher.h:19,
// definition src/raymarcher.cu:176-180), forwarding to the C ABI (include/rrt.h).
//
// Behaviour kept from the reference launcher: void return, no validation visible to the caller, legacy
// default stream, asynchronous (the caller's cudaGraphicsUnmapResources is the sync point,
// src/main.cpp:469), callee keeps no per-frame state.  A lazily created per-device context holds the
// tile ticket and counters.
#include <cuda_runtime.h>

#include <cstring>
#include <mutex>
#include <string>

#include "../../include/rrt.h"
#include "../../include/compat/config.h"
#include "../../include/compat/raymarcher.h"

namespace {
std::mutex g_mu;
rrt_context* g_ctx[64] = {nullptr};
rrt_params g_params;
bool g_params_set = false;
std::string g_err;
}  // namespace

#define RRT_EXPORT

// The reference bakes include/config.h into its kernel; here the macros of include/compat/config.h (same names, same
// values, SPIN_A overridable with -DRRT_COMPAT_SPIN_A) are the parameter block launch_raymarch uses until
// rrt_compat_set_params replaces it -- so editing that header and rebuilding changes the frame, as in the reference.
static void compat_default_params(rrt_params* o) {
    rrt_default_params(o);   // flags: both media, the rounding contract of the reference's own CUDA build
    o->spin_a = SPIN_A;
    o->event_horizon = EVENT_HORIZON;
    o->isco_radius = ISCO_RADIUS;
    o->disk_out = DISK_OUT_M;
    o->disk_h = DISK_H_M;
    o->disk_luminosity = DISK_LUMINOSITY;
    o->disk_opacity = DISK_OPACITY;
    o->exposure = EXPOSURE;
    o->cloud_h = CLOUD_H_M;
    o->cloud_out = CLOUD_OUT_M;
    o->cloud_opacity = CLOUD_OPACITY;
    o->cloud_luminosity = CLOUD_LUMINOSITY;
    o->step_size = STEP_SIZE_M;
    o->disk_temp_ref = DISK_TEMP_REF;
    o->max_steps = MAX_STEPS;
}

extern "C" RRT_EXPORT void rrt_compat_set_params(const rrt_params* prm) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (prm) { g_params = *prm; g_params_set = true; } else { g_params_set = false; }
}

extern "C" RRT_EXPORT const char* rrt_compat_last_error(void) { return g_err.c_str(); }

RRT_EXPORT void launch_raymarch(uchar4* d_out, int w, int h, float time, CameraState cam, cudaTextureObject_t skyboxTex,
                                CameraEffects effects) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { g_err = "launch_raymarch: no current CUDA device"; return; }
    rrt_context* ctx;
    rrt_params prm;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!g_ctx[dev]) {
            if (rrt_context_create(dev, &g_ctx[dev]) != RRT_OK) { g_err = rrt_last_error(nullptr); g_ctx[dev] = nullptr; return; }
        }
        ctx = g_ctx[dev];
        if (g_params_set) prm = g_params; else compat_default_params(&prm);
    }
    rrt_camera c;
    static_assert(sizeof(c) == sizeof(cam), "camera layouts differ");
    std::memcpy(&c, &cam, sizeof(c));
    rrt_effects fx;
    fx.use_bloom = effects.useBloom; fx.bloom_threshold = effects.bloomThreshold; fx.bloom_intensity = effects.bloomIntensity;
    fx.use_vignette = effects.useVignette; fx.vignette_intensity = effects.vignetteIntensity;
    fx.use_ca = effects.useChromaticAberration; fx.ca_amount = effects.caAmount;
    fx.use_lens = effects.useLensDistortion; fx.distortion_amount = effects.distortionAmount;
    int rc = rrt_render(ctx, &prm, &c, &fx, (uint64_t)skyboxTex, time, w, h, nullptr, d_out, RRT_OUT_FRAME, nullptr, nullptr);
    if (rc != RRT_OK) g_err = rrt_last_error(ctx);
}
