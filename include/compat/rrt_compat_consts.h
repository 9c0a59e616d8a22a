// Builds the render path's constant block (rrt::Consts, include/rrt_device.cuh) from the config.h macros, for
// the compat device headers.  Every derived value is written with the reference's own association so the
// compiler folds it to the same binary32 constant the reference's expressions fold to.
#ifndef RRT_COMPAT_CONSTS_H
#define RRT_COMPAT_CONSTS_H

#include "config.h"
#include "../rrt_device.cuh"

namespace rrt_compat {
__host__ __device__ inline rrt::Consts consts() {
    rrt::Consts C{};
    C.horizon_r = EVENT_HORIZON * 1.01f;
    C.acc_rmin = EVENT_HORIZON * 0.5f;
    C.radial_k = -1.5f * EVENT_HORIZON;
    C.drag_k = 2.0f * SPIN_A * EVENT_HORIZON;
    C.spin_a = SPIN_A;
    C.event_horizon = EVENT_HORIZON;
    C.disk_zone_y = DISK_H_M * 5.0f;
    C.disk_zone_r = DISK_OUT_M + 5.0f;
    C.dust_zone_y = CLOUD_H_M * 1.5f;
    C.dust_zone_r = CLOUD_OUT_M;
    C.h[0] = STEP_SIZE_M;        C.h[1] = STEP_SIZE_M * 0.1f; C.h[2] = STEP_SIZE_M * 0.3f; C.h[3] = STEP_SIZE_M * 0.5f;
    for (int i = 0; i < 4; ++i) { C.hh[i] = C.h[i] * 0.5f; C.h6[i] = C.h[i] / 6.0f; }
    C.isco = ISCO_RADIUS;
    C.disk_out = DISK_OUT_M;
    C.disk_h = DISK_H_M;
    C.taper_from = DISK_OUT_M * 0.85f;
    C.taper_span = DISK_OUT_M - DISK_OUT_M * 0.85f;
    C.dust_e1 = DISK_OUT_M * 0.8f;
    C.dust_in_e1 = ISCO_RADIUS + 5.0f;
    C.cloud_hh = CLOUD_H_M * 0.5f;
    C.disk_temp_ref = DISK_TEMP_REF;
    C.disk_luminosity = DISK_LUMINOSITY;
    C.disk_opacity = DISK_OPACITY;
    C.cloud_luminosity = CLOUD_LUMINOSITY;
    C.cloud_opacity = CLOUD_OPACITY;
    C.exposure = EXPOSURE;
    C.max_steps = MAX_STEPS;
    C.flags = 3u;
    C.neg_zero = -0.0f;
    return C;
}
__device__ __forceinline__ rrt::V3 v3(float3 a) { return rrt::mk(a.x, a.y, a.z); }
__device__ __forceinline__ float3 f3(rrt::V3 a) { return make_float3(a.x, a.y, a.z); }
constexpr bool kSpin = (SPIN_A != 0.0f);
}  // namespace rrt_compat

#endif
