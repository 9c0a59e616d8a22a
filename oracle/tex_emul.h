/*
 * tex_emul.h -- host emulation of the texture object the reference samples its skybox with
 * (src/main.cpp:246-263: RGBA8 cudaArray, addressMode[0]=Wrap, [1]=Clamp, filterMode=Linear,
 * readMode=NormalizedFloat, normalizedCoords=1), used by tex2D<float4> at src/raymarcher.cu:139.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_abi.h).
 *
 * Model (CUDA C Programming Guide, "Texture Fetching", linear filtering):
 *   x = N * frac(tx)            (wrap)          y = M * clamp(ty, 0, 1)       (clamp)
 *   xB = x - 0.5, i = floor(xB), alpha = frac(xB) held in 1.8 fixed point
 *   tex = (1-a)(1-b) T[i,j] + a(1-b) T[i+1,j] + (1-a) b T[i,j+1] + a b T[i+1,j+1]
 * with column indices wrapped and row indices clamped.  The fixed-point conversion of the
 * texel-space coordinate is modelled as round-to-nearest on 8 fractional bits; this choice
 * is calibrated against the B200 texture unit by tests/test_gpu_sky.py.
 */
#ifndef RRT_TEX_EMUL_H
#define RRT_TEX_EMUL_H

#include <math.h>
#include <stdint.h>

static inline void tex_emul_fetch(const uint8_t* rgba, int W, int H, float tx, float ty, float out[4]) {
    /* non-finite coordinates (cannot occur for normalised directions) sample texel (0,0) */
    if (!(tx == tx) || !(ty == ty) || fabsf(tx) > 1e30f || fabsf(ty) > 1e30f) {
        for (int c = 0; c < 4; ++c) out[c] = rgba[c] * (1.0f / 255.0f);
        return;
    }
    double fx = (double)tx - floor((double)tx); /* wrap */
    double fy = (double)ty;
    if (fy < 0.0) fy = 0.0;
    if (fy > 1.0) fy = 1.0;
    /* texel-space coordinate, 8 fractional bits, round to nearest */
    long long qx = (long long)floor((fx * (double)W - 0.5) * 256.0 + 0.5);
    long long qy = (long long)floor((fy * (double)H - 0.5) * 256.0 + 0.5);
    long long ix = qx >> 8, iy = qy >> 8; /* arithmetic shift == floor */
    double a = (double)(qx & 255) / 256.0, b = (double)(qy & 255) / 256.0;
    long long ix1 = ix + 1, iy1 = iy + 1;
    ix = ((ix % W) + W) % W;
    ix1 = ((ix1 % W) + W) % W;
    if (iy < 0) iy = 0;
    if (iy > H - 1) iy = H - 1;
    if (iy1 < 0) iy1 = 0;
    if (iy1 > H - 1) iy1 = H - 1;
    const uint8_t* t00 = rgba + 4 * ((size_t)iy * W + ix);
    const uint8_t* t10 = rgba + 4 * ((size_t)iy * W + ix1);
    const uint8_t* t01 = rgba + 4 * ((size_t)iy1 * W + ix);
    const uint8_t* t11 = rgba + 4 * ((size_t)iy1 * W + ix1);
    for (int c = 0; c < 4; ++c) {
        double v = (1.0 - a) * (1.0 - b) * t00[c] + a * (1.0 - b) * t10[c] + (1.0 - a) * b * t01[c] + a * b * t11[c];
        out[c] = (float)(v / 255.0);
    }
}

#endif
