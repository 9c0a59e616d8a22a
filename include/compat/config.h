// Drop-in for the reference's include/config.h: the same macro names with the same values, so code written
// against the reference (src/main.cpp uses WINDOW_WIDTH/WINDOW_HEIGHT/RECORDING_FPS; the device headers use
// the rest) keeps compiling.  In this implementation the render path does NOT bake these in: they are the
// defaults of the run-time parameter block (rrt_params / rrt_default_params in include/rrt.h).  SPIN_A may
// be overridden at compile time with -DRRT_COMPAT_SPIN_A=0.99f for users of the compat device headers.
#ifndef CONFIG_H
#define CONFIG_H

#include <cuda_runtime.h>

// window / recorder (reference config.h:7-9)
#define WINDOW_WIDTH 1000
#define WINDOW_HEIGHT 700
#define RECORDING_FPS 24

// SI constants kept for source compatibility; unused by the path (reference config.h:12-17,26)
#define C_LIGHT 299792458.0f
#define G_CONSTANT 6.67430e-11f
#define SOLAR_MASS 1.98847e30f
#define BH_MASS_SOLAR 4.154e6f
#define M_UNIT (G_CONSTANT * (BH_MASS_SOLAR * SOLAR_MASS) / (C_LIGHT * C_LIGHT))

#define DISK_TEMP_REF 1.5e7f               // reference config.h:18

#ifdef RRT_COMPAT_SPIN_A
#define SPIN_A RRT_COMPAT_SPIN_A
#else
#define SPIN_A 0.0f                        // reference config.h:21 ships Schwarzschild
#endif
#define SPIN_AXIS make_float3(0, 1, 0)     // reference config.h:22

#define EVENT_HORIZON 2.0f                 // reference config.h:29
#define MASS_POS make_float3(0.0f, 0.0f, 0.0f)

#define ISCO_RADIUS 10.0f                  // reference config.h:33-38
#define DISK_OUT_M 25.0f
#define DISK_H_M 0.8f
#define DISK_LUMINOSITY 6.0f
#define DISK_OPACITY 0.4f
#define EXPOSURE 0.8f

#define CLOUD_H_M 0.5f                     // reference config.h:41-44
#define CLOUD_OUT_M 25.0f
#define CLOUD_OPACITY 0.3f
#define CLOUD_LUMINOSITY 0.4f

#define STEP_SIZE_M 0.3f                   // reference config.h:47-48
#define MAX_STEPS 2000

#endif
