// Drop-in for the reference's include/raymarcher.h: same CameraState (:11-16) and the same
// launch_raymarch prototype (:19, C++ linkage, structs by value), implemented by librrt_b200.so
// (csrc/rrt_compat.cu) on top of the C ABI in include/rrt.h.  src/main.cpp:467 links unchanged.
#ifndef RAYMARCHER_H
#define RAYMARCHER_H

#include <vector_types.h>
#include <texture_types.h>
#include "camera_effects/camera_settings.h"

// Pinhole camera basis handed from the host to the render path (48 bytes, four packed float3).
struct CameraState {
    float3 pos;
    float3 forward;
    float3 right;
    float3 up;
};

static_assert(sizeof(CameraState) == 48, "CameraState must stay four packed float3");

// Renders one w x h frame into the caller-owned device buffer d_out (uchar4, pixel (x,y) stored at
// [(h-1-y)*w + x]) on the legacy default stream, asynchronously, exactly like the reference launcher.
// Tuning: the macros of include/compat/config.h as they were when librrt_b200.so was built (csrc/rrt_compat.cu copies
// them into the run-time parameter block; SPIN_A from -DRRT_COMPAT_SPIN_A), until rrt_compat_set_params replaces
// that block at run time; errors are swallowed to keep `void`.
void launch_raymarch(uchar4* d_out, int w, int h, float time, CameraState cam, cudaTextureObject_t skyboxTex,
                     CameraEffects effects);

// Extensions (not in the reference): override the parameter block used by launch_raymarch, and fetch
// the last error the shim swallowed.  Both are plain C symbols.
extern "C" {
struct rrt_params;
void rrt_compat_set_params(const struct rrt_params* prm);
const char* rrt_compat_last_error(void);
}

#endif
