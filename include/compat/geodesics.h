// Drop-in for the reference's include/geodesics.h: the two device functions north_star names, with the
// reference's names and signatures, implemented by the B200 path's own device math (include/rrt_device.cuh).
// These are header-only device functions, so their rounding contract is the INCLUDING translation unit's
// (include/rrt_device.cuh): compile it -fmad=false without RRT_FMAD for the strict contract (bit-identical to the
// reference math evaluated without FMA contraction), or -fmad=true -DRRT_FMAD=1 for the fusion schedule of the
// reference's own CUDA build -- the contract librrt_b200.so's launch_raymarch / rrt_default_params use by default.
#ifndef GEODESICS_H
#define GEODESICS_H

#include <cuda_runtime.h>
#include "rrt_compat_consts.h"

// Doppler x gravitational redshift factor g of the disk gas seen by a ray (reference geodesics.h:11-25).
__device__ __forceinline__ float calculateRedshiftFactor(float3 p_rel, float3 ray_vel) {
    return rrt::redshift(rrt_compat::consts(), rrt_compat::v3(p_rel), rrt_compat::v3(ray_vel));
}

// Acceleration of the ray: Binet pseudo-force -1.5 Rs |p x v|^2 / r^5 p plus the frame-drag term
// (2 a Rs / r^3) (s x p); zero inside r < Rs/2 (reference geodesics.h:30-45).
__device__ __forceinline__ float3 getGeodesicAcc(float3 p_rel, float3 v) {
    return rrt_compat::f3(rrt::geodesic_acc<rrt_compat::kSpin>(rrt_compat::consts(), rrt_compat::v3(p_rel), rrt_compat::v3(v)));
}

#endif
