// Drop-in for the reference's include/densities.h (getDiskTemperature :12-15, getAccretionDensity :20-62,
// getDustCloudDensity :69-132) over include/rrt_device.cuh.
#ifndef DENSITIES_H
#define DENSITIES_H

#include <cuda_runtime.h>
#include "rrt_compat_consts.h"

__device__ __forceinline__ float getDiskTemperature(float r) { return rrt::disk_temperature(rrt_compat::consts(), r); }
__device__ __forceinline__ float getAccretionDensity(float3 p, float time) {
    return rrt::disk_density(rrt_compat::consts(), rrt_compat::v3(p), time);
}
__device__ __forceinline__ float getDustCloudDensity(float3 p, float time) {
    return rrt::dust_density(rrt_compat::consts(), rrt_compat::v3(p), time);
}

#endif
