// Drop-in for the reference's include/integrators.h: integrate_euler / integrate_rk4 with the reference's
// signatures (in-place update of position and velocity by one step h), implemented by include/rrt_device.cuh.
#ifndef INTEGRATORS_H
#define INTEGRATORS_H

#include <cuda_runtime.h>
#include "geodesics.h"

// first-order step (reference integrators.h:12-18); unused by the render loop, kept for the interface
__device__ __forceinline__ void integrate_euler(float3& p, float3& v, float h) {
    rrt::V3 pp = rrt_compat::v3(p), vv = rrt_compat::v3(v);
    rrt::euler_step<rrt_compat::kSpin>(rrt_compat::consts(), pp, vv, h);
    p = rrt_compat::f3(pp);
    v = rrt_compat::f3(vv);
}

// classic fourth-order Runge-Kutta step, no error control (reference integrators.h:23-59)
__device__ __forceinline__ void integrate_rk4(float3& p, float3& v, float h) {
    rrt::V3 pp = rrt_compat::v3(p), vv = rrt_compat::v3(v);
    const float r2 = rrt::dot3(pp, pp);
    rrt::rk4_step<rrt_compat::kSpin>(rrt_compat::consts(), pp, vv, h, h * 0.5f, h / 6.0f, r2, sqrtf(r2));
    p = rrt_compat::f3(pp);
    v = rrt_compat::f3(vv);
}

#endif
