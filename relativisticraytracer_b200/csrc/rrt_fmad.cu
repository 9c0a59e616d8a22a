// rrt_fmad.cu -- the render kernel and the probe kernels under the rounding contract of the reference's OWN
// CUDA build (nvcc default -fmad=true): selected at run time with RRT_FLAG_FMAD.  See include/rrt_device.cuh
// "Rounding contracts" and csrc/Makefile (this unit is the only one compiled with -fmad=true).
#define RRT_FMAD 1
#include "rrt_kernel.cuh"
