/*
 * rrt_oracle.c -- the PORT checker ("kind": "port"): a plain-C restatement of the reference's
 * per-pixel path (ray setup, zone-stepped RK4 over the Binet pseudo-force + frame-drag term,
 * horizon / escape / exhaustion, volumetric disk + dust transfer, equirect sky, effects, tonemap).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_abi.h).  The product never links or loads this file.
 *
 * Parity pin: tests/test_oracle_vs_ref.py checks every entry point bit-for-bit against
 * oracle/_ref/libref_host.so (the reference's own headers compiled for the host) wherever that
 * library is present, and tests/test_oracle_golden.py checks it against tests/golden/ (npz files), which
 * tests/tools/make_golden.py generated from that same reference build.  The reference ships no tests or
 * golden vectors of its own, so those two are the whole pin.
 *
 * Canonical rounding: compile with -ffp-contract=off; every + - * / sqrt below is one IEEE
 * binary32 operation in exactly the order the reference's expressions associate.  Transcendentals
 * are glibc's (the reference's CUDA build uses libdevice: expect ulp-level differences there).
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_abi.h"
#include "tex_emul.h"

typedef struct { float x, y, z; } v3;

static const float K_PI = 3.1415926535f; /* math_utils.h:7 */

/* ---- rounding contracts ------------------------------------------------------------------------
 * fm = 0 (default): the reference's expressions exactly as written, no contraction (this file is built with
 *   -ffp-contract=off) -- the twin of the reference headers compiled for the host.
 * fm = 1 (ORA_FLAG_FMAD): the twin of the reference's own CUDA build.  nvcc's default -fmad=true fuses a*b+c;
 *   which operations it fuses for the reference's sources (nvcc 12.9, sm_100a) was read off the SASS of
 *   raymarch_kernel and is restated here operation by operation with fmaf(): add(x, y) with x a product ->
 *   fma(x.a, x.b, y), else with y a product -> fma(y.a, y.b, x), plus the two places where the kernel deviates
 *   from that rule (loop-header |p|^2, first RK4 stage's radial + drag sum).  Covers ray setup, the geodesic
 *   integration, the value noise and the final assembly; the density / redshift expressions stay unfused here
 *   (on the device they are left to nvcc), which moves them by ulps only. */
static inline float mad(int fm, float a, float b, float c) { return fm ? fmaf(a, b, c) : a * b + c; }           /* a*b + c */
static inline float msub2(int fm, float a, float b, float c, float d) { return fm ? fmaf(a, b, -(c * d)) : a * b - c * d; }
static inline float mad2(int fm, float a, float b, float c, float d) { return fm ? fmaf(a, b, c * d) : a * b + c * d; }

/* ---- float3 helpers, math_utils.h:11-48 ------------------------------------------------------ */
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline float v_dot(int fm, v3 a, v3 b) {                                                     /* :11-13 */
    return fm ? fmaf(a.z, b.z, fmaf(a.x, b.x, a.y * b.y)) : a.x * b.x + a.y * b.y + a.z * b.z;
}
static inline v3 v_cross(int fm, v3 a, v3 b) {                                                      /* :15-17 */
    return V(msub2(fm, a.y, b.z, a.z, b.y), msub2(fm, a.z, b.x, a.x, b.z), msub2(fm, a.x, b.y, a.y, b.x));
}
static inline float v_len(int fm, v3 a) { return sqrtf(v_dot(fm, a, a)); }                          /* :19-21 */
static inline v3 v_unit(int fm, v3 a) {                                                             /* :23-27 */
    float m = v_len(fm, a);
    if (m < 1e-6f) return V(0, 0, 0);
    return V(a.x / m, a.y / m, a.z / m);
}
static inline v3 v_sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }                   /* :29-31 */
static inline v3 v_add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }                   /* :33-35 */
static inline v3 v_scale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }                    /* :37-39 */
static inline v3 v_axpy(int fm, v3 y, v3 x, float a) { return V(mad(fm, x.x, a, y.x), mad(fm, x.y, a, y.y), mad(fm, x.z, a, y.z)); } /* y + x*a */
static inline float mixf(int fm, float a, float b, float t) { return mad(fm, t, b - a, a); }        /* :41-43 */
static inline float sstep(float e0, float e1, float x) {                                            /* :45-48 */
    float t = fminf(fmaxf((x - e0) / (e1 - e0), 0.0f), 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

/* ---- value noise, math_utils.h:91-121 -------------------------------------------------------- */
static float hash31(int fm, v3 p) {                                                                 /* :91-96 */
    float a = fmodf(p.x * 0.1031f, 1.0f), b = fmodf(p.y * 0.1031f, 1.0f), c = fmodf(p.z * 0.1031f, 1.0f);
    float d = v_dot(fm, V(a, b, c), V(b + 33.33f, c + 33.33f, a + 33.33f));
    a += d; b += d; c += d;
    return fmodf((a + b) * c, 1.0f);
}

static float noise3(int fm, v3 p) {                                                                 /* :98-110 */
    v3 i = V(floorf(p.x), floorf(p.y), floorf(p.z));
    v3 f = V(p.x - i.x, p.y - i.y, p.z - i.z);
    float ux = f.x * f.x * (3.0f - 2.0f * f.x);
    float uy = f.y * f.y * (3.0f - 2.0f * f.y);
    float uz = f.z * f.z * (3.0f - 2.0f * f.z);
    float c[2][2][2];
    for (int dz = 0; dz < 2; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) c[dz][dy][dx] = hash31(fm, v_add(i, V((float)dx, (float)dy, (float)dz)));
    float lo = mixf(fm, mixf(fm, c[0][0][0], c[0][0][1], ux), mixf(fm, c[0][1][0], c[0][1][1], ux), uy);
    float hi = mixf(fm, mixf(fm, c[1][0][0], c[1][0][1], ux), mixf(fm, c[1][1][0], c[1][1][1], ux), uy);
    return mixf(fm, lo, hi, uz);
}

static float fbm(int fm, v3 p, int octaves) {                                                       /* :112-121 */
    float acc = 0.0f, amp = 0.5f;
    for (int k = 0; k < octaves; ++k) {
        acc = mad(fm, amp, noise3(fm, p), acc);
        p = V(mad(fm, p.x, 2.05f, 10.0f), mad(fm, p.y, 2.05f, 10.0f), mad(fm, p.z, 2.05f, 10.0f));
        amp *= 0.5f;
    }
    return acc;
}

/* ---- geodesics.h ----------------------------------------------------------------------------- */
#define FM(P) (((P)->flags & ORA_FLAG_FMAD) != 0)
/* r2 is |q|^2 as the caller's contract computes it; `first` marks the first RK4 stage (see "rounding contracts") */
static v3 geodesic_acc_r2(const ora_params* P, v3 q, v3 v, float r2, int first) {                   /* geodesics.h:30-45 */
    const int fm = FM(P);
    float r = sqrtf(r2);
    if (r < P->event_horizon * 0.5f) return V(0, 0, 0);
    v3 L = v_cross(fm, q, v);
    float L2 = v_dot(fm, L, L);
    float radial = -1.5f * P->event_horizon * L2 / (r2 * r2 * r);
    float drag = (2.0f * P->spin_a * P->event_horizon) / (r2 * r);
    if (!fm) {
        v3 a_rad = v_scale(q, radial);
        v3 axis_x_q = v_cross(0, V(0, 1, 0), q);           /* SPIN_AXIS, config.h:22 */
        return v_add(a_rad, v_scale(axis_x_q, drag));
    }
    /* cross((0,1,0), q) = (q.z, +-0, -q.x) exactly for finite q */
    if (first) return V(fmaf(q.x, radial, q.z * drag), q.y * radial, fmaf(q.z, radial, -q.x * drag));
    return V(fmaf(q.z, drag, q.x * radial), q.y * radial, fmaf(-q.x, drag, q.z * radial));
}
static v3 geodesic_acc(const ora_params* P, v3 q, v3 v) { return geodesic_acc_r2(P, q, v, v_dot(FM(P), q, q), 0); }
/* |p|^2 of the loop header (raymarcher.cu:43-44), shared with the first RK4 stage */
static inline float norm2_loop(int fm, v3 p) { return fm ? fmaf(p.y, p.y, p.x * p.x) + p.z * p.z : v_dot(0, p, p); }

static float redshift_factor(const ora_params* P, v3 q, v3 ray_v) {                                 /* geodesics.h:11-25 */
    const int fm = FM(P);
    float r = v_len(fm, q);
    if (r < P->event_horizon * 1.01f) return 0.0f;
    float g_grav = sqrtf(1.0f - P->event_horizon / r);
    float beta = 1.0f / (powf(r, 1.5f) + P->spin_a);
    v3 gas = v_unit(fm, V(-q.z, 0, q.x));
    float mu = v_dot(fm, ray_v, gas);
    float gamma = 1.0f / sqrtf(1.0f - beta * beta);
    float g_dop = 1.0f / (gamma * (1.0f - beta * mu));
    return g_grav * g_dop;
}

/* ---- integrators.h ---------------------------------------------------------------------------- */
static void step_euler(const ora_params* P, v3* p, v3* v, float h) {                                /* integrators.h:12-18 */
    const int fm = FM(P);
    v3 q = v_sub(*p, V(0, 0, 0));
    v3 a = geodesic_acc(P, q, *v);
    *p = v_axpy(fm, *p, *v, h);
    *v = v_axpy(fm, *v, a, h);
}

static void step_rk4(const ora_params* P, v3* p, v3* v, float h) {                                  /* integrators.h:23-59 */
    const int fm = FM(P);
    const v3 origin = V(0.0f, 0.0f, 0.0f);                 /* MASS_POS, config.h:30 */
    v3 p0 = *p, v0 = *v;
    v3 q1 = v_sub(p0, origin);
    v3 kv1 = geodesic_acc_r2(P, q1, v0, norm2_loop(fm, q1), 1), kp1 = v0;
    v3 v2 = v_axpy(fm, v0, kv1, h * 0.5f);
    v3 kv2 = geodesic_acc(P, v_sub(v_axpy(fm, p0, kp1, h * 0.5f), origin), v2), kp2 = v2;
    v3 v3_ = v_axpy(fm, v0, kv2, h * 0.5f);
    v3 kv3 = geodesic_acc(P, v_sub(v_axpy(fm, p0, kp2, h * 0.5f), origin), v3_), kp3 = v3_;
    v3 v4 = v_axpy(fm, v0, kv3, h);
    v3 kv4 = geodesic_acc(P, v_sub(v_axpy(fm, p0, kp3, h), origin), v4), kp4 = v4;
    /* 2*x is exact, so k + 2*k' rounds the same fused or not */
    v3 kv = v_add(kv1, v_add(v_scale(kv2, 2.0f), v_add(v_scale(kv3, 2.0f), kv4)));
    v3 kp = v_add(kp1, v_add(v_scale(kp2, 2.0f), v_add(v_scale(kp3, 2.0f), kp4)));
    *v = v_axpy(fm, *v, kv, h / 6.0f);
    *p = v_axpy(fm, *p, kp, h / 6.0f);
}

/* ---- densities.h ------------------------------------------------------------------------------ */
static float disk_temperature(const ora_params* P, float r) {                                       /* densities.h:12-15 */
    if (r < P->isco_radius) return 0.0f;
    return P->disk_temp_ref * powf(r / P->isco_radius, -0.75f);
}

static float disk_density(const ora_params* P, v3 p, float time) {                                  /* densities.h:20-62 */
    const int fm = FM(P);
    float r = v_len(0, V(p.x, 0.0f, p.z));   /* (x*x + 0*0) + z*z, unfused in both contracts */
    if (r < P->isco_radius || r > P->disk_out) return 0.0f;
    float taper = 1.0f;
    float taper_from = P->disk_out * 0.85f;
    if (r > taper_from) {
        taper = 1.0f - (r - taper_from) / (P->disk_out - taper_from);
        taper *= taper;
    }
    float hgt = P->disk_h * powf(P->isco_radius / r, 0.5f);
    float vert = expf(-(p.y * p.y) / (2.0f * hgt * hgt + 1e-7f));
    float radial = powf(P->isco_radius / r, 0.4f);
    float envelope = vert * radial * taper;
    float phi = atan2f(p.z, p.x);
    float omega = 3.5f * powf(P->isco_radius / r, 1.5f);
    float ang = phi - time * omega;
    v3 rot = V(r * cosf(ang), p.y * 4.0f, r * sinf(ang));
    float evo = time * 0.35f;
    v3 nc = v_add(v_scale(rot, 0.45f), V(0, evo, 0));
    float n = fbm(fm, nc, 5);
    float streak = fmaxf(0.0f, n - 0.32f);
    streak = powf(streak * 2.8f, 1.6f);
    streak = fminf(6.0f, streak);
    return envelope * (0.02f + 5.0f * streak);
}

static float dust_density(const ora_params* P, v3 p, float time) {                                  /* densities.h:69-132 */
    const int fm = FM(P);
    float r = v_len(0, V(p.x, 0.0f, p.z));   /* (x*x + 0*0) + z*z, unfused in both contracts */
    if (r < P->isco_radius || r > P->disk_out) return 0.0f;
    float outer = sstep(P->disk_out, P->disk_out * 0.8f, r);
    float inner = sstep(P->isco_radius, P->isco_radius + 5.0f, r);
    float hgt = P->cloud_h * 0.5f * powf(P->isco_radius / r, 0.2f);
    float vert = expf(-(p.y * p.y) / (2.0f * hgt * hgt + 1e-7f));
    float base = vert * outer * inner;
    if (base < 0.001f) return 0.0f;
    float phi = atan2f(p.z, p.x);
    float omega = 1.0f * powf(P->isco_radius / r, 1.5f);
    float ang = phi - time * omega;
    v3 c0 = V(r * 0.8f, p.y * 15.0f, ang * 10.0f);
    v3 w1 = V(fbm(fm, v_scale(c0, 0.15f), 2), fbm(fm, v_add(v_scale(c0, 0.15f), V(1, 2, 3)), 2),
              fbm(fm, v_add(v_scale(c0, 0.15f), V(4, 5, 6)), 2));
    v3 c1 = v_add(c0, v_scale(w1, 3.0f));
    v3 w2 = V(fbm(fm, v_scale(c1, 0.4f), 2), fbm(fm, v_add(v_scale(c1, 0.4f), V(2, 1, 0)), 2),
              fbm(fm, v_add(v_scale(c1, 0.4f), V(0, 3, 1)), 2));
    v3 cf = v_add(c0, v_scale(w2, 1.5f));
    float n = 0.0f, amp = 1.0f, freq = 1.0f;
    for (int k = 0; k < 5; ++k) {
        float nv = noise3(fm, v_scale(cf, freq));
        float wisp = 1.0f - fabsf(nv * 2.0f - 1.0f);
        n += wisp * amp;
        amp *= 0.5f;
        freq *= 2.1f;
    }
    float strands = sstep(0.4f, 0.8f, n * 0.55f);
    strands = powf(strands, 4.0f);
    float detail = fbm(fm, v_add(v_scale(cf, 4.0f), V(0, time * 0.5f, 0)), 2);
    strands *= (0.6f + 0.4f * detail);
    return base * strands * 12.0f;
}

/* ---- post_processing.h ------------------------------------------------------------------------ */
static void lens_distort(int fm, float* u, float* v, float k) {                                     /* post_processing.h:19-24 */
    float tu = *u - 0.5f, tv = *v - 0.5f;
    float r2 = mad2(fm, tu, tu, tv, tv);
    float f = mad(fm, r2, k, 1.0f);
    *u = mad(fm, tu, f, 0.5f);
    *v = mad(fm, tv, f, 0.5f);
}

/* ---- one pixel: raymarch_kernel, src/raymarcher.cu:16-173 -------------------------------------- */
typedef struct {
    float hdr[3], T, dir[3], I[3];
    v3 p, v;
    uint8_t cls, rgba[4];
    int steps;
    uint32_t disk_evals, dust_evals, dense;
} pixel_out;

static void trace_pixel(const ora_params* P, int x, int y, int width, int height, float time, const ora_camera* cam,
                        const ora_effects* fx, const uint8_t* sky, int sky_w, int sky_h, pixel_out* o) {
    const int want_disk = (P->flags & ORA_FLAG_DISK) != 0, want_dust = (P->flags & ORA_FLAG_DUST) != 0;
    const int fm = FM(P);
    float uvx = (float)x / width, uvy = (float)y / height;                                          /* :20 */
    if (fx->use_lens) lens_distort(fm, &uvx, &uvy, fx->distortion_amount);                              /* :23-25 */
    float uc = uvx * 2.0f - 1.0f;                                                                   /* :27 (2*x exact: same fused) */
    float vc = uvy * 2.0f - 1.0f;                                                                   /* :28 */
    float aspect = (float)width / height;                                                           /* :29 */
    uc *= aspect;                                                                                   /* :30 */
    v3 F = V(cam->forward[0], cam->forward[1], cam->forward[2]);
    v3 R = V(cam->right[0], cam->right[1], cam->right[2]);
    v3 U = V(cam->up[0], cam->up[1], cam->up[2]);
    v3 p = V(cam->pos[0], cam->pos[1], cam->pos[2]);                                                /* :32 */
    v3 vel = v_unit(fm, V(mad2(fm, R.x, uc, U.x, vc) + F.x, mad2(fm, R.y, uc, U.y, vc) + F.y,
                          mad2(fm, R.z, uc, U.z, vc) + F.z));                                        /* :33-34 */

    float I[3] = {0, 0, 0}, T = 1.0f;                                                               /* :36-37 */
    int captured = 0, touched = 0, steps = 0, it;
    uint32_t n_disk = 0, n_dust = 0, n_dense = 0;
    for (it = 0; it < P->max_steps; ++it) {                                                         /* :41 */
        v3 q = v_sub(p, V(0.0f, 0.0f, 0.0f));                                                       /* :42 */
        float r2 = norm2_loop(fm, q);
        float r = sqrtf(r2);                                                                        /* :44 */
        if (r < P->event_horizon * 1.01f) { captured = 1; T = 0.0f; break; }                        /* :47-51 */
        float h = P->step_size;                                                                     /* :54 */
        int near_bh = r < 18.0f;                                                                    /* :56 */
        int disk_zone = fabsf(q.y) < P->disk_h * 5.0f && r < P->disk_out + 5.0f;                    /* :57 */
        int dust_zone = fabsf(q.y) < P->cloud_h * 1.5f && r < P->cloud_out;                         /* :58 */
        if (near_bh) h *= 0.1f; else if (disk_zone) h *= 0.3f; else if (dust_zone) h *= 0.5f;       /* :60-62 */
        step_rk4(P, &p, &vel, h);                                                                   /* :64 */
        ++steps;
        if (disk_zone || dust_zone) {                                                               /* :67 */
            float dd = 0.0f, dc = 0.0f;
            if (disk_zone && want_disk) { dd = disk_density(P, q, time); ++n_disk; }                /* :68 */
            if (dust_zone && want_dust) { dc = dust_density(P, q, time); ++n_dust; }                /* :69 */
            if (dd > 0.001f || dc > 0.001f) {                                                       /* :71 */
                float e[3] = {0, 0, 0}, kappa = 0;
                touched = 1;
                ++n_dense;
                if (dd > 0.001f) {                                                                  /* :76-88 */
                    float g = redshift_factor(P, q, vel);
                    float Tk = disk_temperature(P, r);
                    float tn = powf(Tk / P->disk_temp_ref, 0.5f);
                    float bol = powf(g, 4.0f) * tn * dd * P->disk_luminosity;
                    float ct = g * powf(Tk / P->disk_temp_ref, 0.4f) * 2.5f;
                    e[0] += 1.0f * bol;
                    e[1] += fminf(0.25f, 0.12f * ct) * bol;
                    e[2] += fmaxf(0.0f, 0.01f * (ct - 2.0f)) * bol;
                    kappa += dd * P->disk_opacity;
                }
                if (dc > 0.001f) {                                                                  /* :91-105 */
                    float g = redshift_factor(P, q, vel);
                    float light = 0.5f + 3.0f * powf(P->isco_radius / fmaxf(r, P->isco_radius), 1.2f);
                    float J = dc * P->cloud_luminosity * light;
                    float sh = sstep(0.7f, 1.3f, g);
                    e[0] += 0.60f * J * mixf(fm, 1.2f, 0.8f, sh);
                    e[1] += 0.65f * J * mixf(fm, 0.8f, 1.1f, sh);
                    e[2] += 0.80f * J * mixf(fm, 0.6f, 1.4f, sh);
                    kappa += dc * P->cloud_opacity;
                }
                float tau = kappa * h;                                                              /* :107 */
                float s = expf(-tau);
                float w = (1.0f - s) * T;
                I[0] = mad(fm, e[0], w, I[0]); I[1] = mad(fm, e[1], w, I[1]); I[2] = mad(fm, e[2], w, I[2]);   /* :111-113 */
                T *= s;                                                                             /* :115 */
            }
        }
        if (r > 250.0f && v_dot(fm, q, vel) > 0) break;                                                 /* :120 */
    }
    const int exhausted = it >= P->max_steps;

    float bg[3] = {0, 0, 0};
    v3 d = V(0, 0, 0);
    if (!captured) {                                                                                /* :128-146 */
        d = v_unit(fm, vel);
        float off = fx->use_ca ? fx->ca_amount : 0.0f;
        float offs[3] = {off, 0.0f, -off};
        for (int c = 0; c < 3; ++c) {
            float phi = atan2f(d.z, d.x) + offs[c];
            float theta = asinf(d.y);
            float tx = 0.5f + phi / (2.0f * K_PI);
            float ty = 0.5f - theta / K_PI;
            float tap[4];
            tex_emul_fetch(sky, sky_w, sky_h, tx, ty, tap);
            bg[c] = tap[c];
        }
    }
    float hdr[3];
    for (int c = 0; c < 3; ++c) hdr[c] = mad(fm, bg[c], T, I[c]);                                          /* :148-150 */

    memcpy(o->hdr, hdr, sizeof hdr);
    o->T = T;
    o->dir[0] = d.x; o->dir[1] = d.y; o->dir[2] = d.z;
    memcpy(o->I, I, sizeof I);
    o->p = p; o->v = vel;
    o->steps = steps;
    o->disk_evals = n_disk; o->dust_evals = n_dust; o->dense = n_dense;
    o->cls = (uint8_t)((captured ? ORA_CLS_CAPTURED : (touched ? ORA_CLS_DISK_HIT : ORA_CLS_ESCAPED)) |
                       (exhausted ? ORA_CLSF_EXHAUSTED : 0u) | (touched ? ORA_CLSF_TOUCHED : 0u));

    if (fx->use_bloom) {                                                                            /* :154-157, post_processing.h:27-31 */
        float lum = v_dot(fm, V(hdr[0], hdr[1], hdr[2]), V(0.2126f, 0.7152f, 0.0722f));
        float b0 = 0, b1 = 0, b2 = 0;
        if (lum > fx->bloom_threshold) { b0 = hdr[0]; b1 = hdr[1]; b2 = hdr[2]; }
        hdr[0] = mad(fm, b0, fx->bloom_intensity, hdr[0]);
        hdr[1] = mad(fm, b1, fx->bloom_intensity, hdr[1]);
        hdr[2] = mad(fm, b2, fx->bloom_intensity, hdr[2]);
    }
    if (fx->use_vignette) {                                                                         /* :159-161, post_processing.h:13-17 */
        float dx = uvx - 0.5f, dy = uvy - 0.5f;
        float dist = sqrtf(mad2(fm, dx, dx, dy, dy) + 0.0f);                                        /* length((dx, dy, 0)) */
        /* smoothstep(0.8, 0.2, dist * intensity): the product is fused into the subtraction of edge0 */
        float tt = fminf(fmaxf((fm ? fmaf(dist, fx->vignette_intensity, -0.8f) : dist * fx->vignette_intensity - 0.8f) / (0.2f - 0.8f), 0.0f), 1.0f);
        float vg = tt * tt * (3.0f - 2.0f * tt);
        hdr[0] *= vg; hdr[1] *= vg; hdr[2] *= vg;
    }
    for (int c = 0; c < 3; ++c) {                                                                   /* :164-172 */
        float t = 1.0f - expf(-hdr[c] * P->exposure);
        o->rgba[c] = (unsigned char)(t * 255);
    }
    o->rgba[3] = 255;
}

/* ---- host camera, src/main.cpp:141-167 and :176-203; paths from src/camera_paths.cpp:31-73 ---- */
static void camera_from(v3 pos, float yaw, float pitch, ora_camera* out) {
    float ry = yaw * 3.14159f / 180.0f;                    /* note: 5-digit pi, main.cpp:142 */
    float rp = pitch * 3.14159f / 180.0f;
    v3 f = V(sinf(ry) * cosf(rp), sinf(rp), cosf(ry) * cosf(rp));
    float m = sqrtf(f.x * f.x + f.y * f.y + f.z * f.z);
    f.x /= m; f.y /= m; f.z /= m;
    v3 wu = V(0.0f, 1.0f, 0.0f);
    v3 rt = V(wu.y * f.z - wu.z * f.y, wu.z * f.x - wu.x * f.z, wu.x * f.y - wu.y * f.x);
    float rm = sqrtf(rt.x * rt.x + rt.y * rt.y + rt.z * rt.z);
    rt.x /= rm; rt.y /= rm; rt.z /= rm;
    v3 up = V(f.y * rt.z - f.z * rt.y, f.z * rt.x - f.x * rt.z, f.x * rt.y - f.y * rt.x);
    out->pos[0] = pos.x; out->pos[1] = pos.y; out->pos[2] = pos.z;
    out->forward[0] = f.x; out->forward[1] = f.y; out->forward[2] = f.z;
    out->right[0] = rt.x; out->right[1] = rt.y; out->right[2] = rt.z;
    out->up[0] = up.x; out->up[1] = up.y; out->up[2] = up.z;
}

typedef struct { float t; v3 pos; float yaw, pitch; } keyframe;

static const keyframe k_path0[] = {                        /* "Gargantua Fly-By", camera_paths.cpp:36-42 */
    {0.0f, {0.0f, 15.0f, -80.0f}, 0.0f, -10.6f},   {6.0f, {15.0f, 3.0f, -30.0f}, -26.6f, -5.1f},
    {12.0f, {35.0f, 0.8f, 10.0f}, -106.0f, -1.2f}, {18.0f, {5.0f, 1.5f, 50.0f}, -174.3f, -1.7f},
    {25.0f, {-20.0f, 12.0f, 70.0f}, -196.0f, -9.3f}};
static const keyframe k_path1[] = {                        /* "Event Horizon Focus", camera_paths.cpp:49-55 */
    {0.0f, {40.0f, 2.0f, 0.0f}, -90.0f, 0.0f},     {8.0f, {0.0f, 5.0f, 40.0f}, -180.0f, -5.0f},
    {16.0f, {-40.0f, 2.0f, 0.0f}, -270.0f, 0.0f},  {24.0f, {0.0f, -5.0f, -40.0f}, -360.0f, 5.0f},
    {32.0f, {40.0f, 2.0f, 0.0f}, -450.0f, 0.0f}};
static const keyframe k_path2[] = {                        /* "Horizon Skimmer", camera_paths.cpp:63-71 */
    {0.0f, {0.0f, 10.0f, -60.0f}, 0.0f, -9.5f},    {8.0f, {15.0f, 2.0f, -15.0f}, -45.0f, -4.7f},
    {14.0f, {4.2f, 0.6f, 4.2f}, -90.0f, -5.7f},    {20.0f, {-20.0f, 8.0f, -20.0f}, -225.0f, -20.0f},
    {26.0f, {-20.0f, 8.0f, -20.0f}, 20.0f, -10.0f}, {29.0f, {-30.0f, 2.0f, -30.0f}, 45.0f, -2.7f}};

static float spline1(float a, float b, float c, float d, float t, float t2, float t3) {             /* camera_paths.cpp:10-15 */
    return 0.5f * ((2.0f * b) + (-a + c) * t + (2.0f * a - 5.0f * b + 4.0f * c - d) * t2 +
                   (-a + 3.0f * b - 3.0f * c + d) * t3);
}
static v3 catmull(v3 a, v3 b, v3 c, v3 d, float t) {                                                /* camera_paths.cpp:6-22 */
    float t2 = t * t, t3 = t2 * t;
    return V(spline1(a.x, b.x, c.x, d.x, t, t2, t3), spline1(a.y, b.y, c.y, d.y, t, t2, t3),
             spline1(a.z, b.z, c.z, d.z, t, t2, t3));
}
static float angle_mix(float a, float b, float t) {                                                 /* camera_paths.cpp:25-29 */
    float diff = fmodf(b - a + 180.0f, 360.0f) - 180.0f;
    if (diff < -180.0f) diff += 360.0f;
    return a + diff * t;
}

/* ================================== exported ABI ================================================ */

void ora_default_params(ora_params* out) {                 /* include/config.h */
    out->spin_a = 0.0f;
    out->event_horizon = 2.0f;
    out->isco_radius = 10.0f;
    out->disk_out = 25.0f;
    out->disk_h = 0.8f;
    out->disk_luminosity = 6.0f;
    out->disk_opacity = 0.4f;
    out->exposure = 0.8f;
    out->cloud_h = 0.5f;
    out->cloud_out = 25.0f;
    out->cloud_opacity = 0.3f;
    out->cloud_luminosity = 0.4f;
    out->step_size = 0.3f;
    out->disk_temp_ref = 1.5e7f;
    out->max_steps = 2000;
    out->flags = ORA_FLAG_DISK | ORA_FLAG_DUST;
}

void ora_default_effects(ora_effects* out) {               /* camera_settings.h:5-16 */
    out->use_bloom = 1; out->bloom_threshold = 0.8f; out->bloom_intensity = 0.5f;
    out->use_vignette = 1; out->vignette_intensity = 0.4f;
    out->use_ca = 0; out->ca_amount = 0.005f;
    out->use_lens = 1; out->distortion_amount = 0.15f;
}

int ora_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* Size of the OpenMP team of the next render (launchers such as torchrun export OMP_NUM_THREADS=1). */
void ora_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void ora_camera_from(const float pos[3], float yaw_deg, float pitch_deg, ora_camera* out) {
    camera_from(V(pos[0], pos[1], pos[2]), yaw_deg, pitch_deg, out);
}

int ora_path_state(int path_index, float t, ora_camera* out, float pyp[5]) {                        /* main.cpp:176-203 */
    const keyframe* k; int n;
    switch (path_index) {
        case 0: k = k_path0; n = 5; break;
        case 1: k = k_path1; n = 5; break;
        case 2: k = k_path2; n = 6; break;
        default: return -1;
    }
    v3 pos = k[0].pos; float yaw = k[0].yaw, pitch = k[0].pitch; int found = 0;
    if (t <= k[0].t) { found = 1; }
    else if (t >= k[n - 1].t) { pos = k[n - 1].pos; yaw = k[n - 1].yaw; pitch = k[n - 1].pitch; found = 1; }
    else {
        for (int i = 0; i < n - 1; ++i) {
            if (t >= k[i].t && t <= k[i + 1].t) {
                float f = (t - k[i].t) / (k[i + 1].t - k[i].t);
                int i0 = i - 1 < 0 ? 0 : i - 1, i3 = i + 2 > n - 1 ? n - 1 : i + 2;
                pos = catmull(k[i0].pos, k[i].pos, k[i + 1].pos, k[i3].pos, f);
                yaw = angle_mix(k[i].yaw, k[i + 1].yaw, f);
                pitch = angle_mix(k[i].pitch, k[i + 1].pitch, f);
                found = 1;
                break;
            }
        }
    }
    if (!found) return -2;
    camera_from(pos, yaw, pitch, out);
    if (pyp) { pyp[0] = pos.x; pyp[1] = pos.y; pyp[2] = pos.z; pyp[3] = yaw; pyp[4] = pitch; }
    return 0;
}

int ora_render(const ora_params* prm, const ora_camera* cam, const ora_effects* fx, const uint8_t* sky_rgba, int sky_w,
               int sky_h, float time, int w, int h, int y0, int y1, uint8_t* out_rgba, const ora_planes* planes,
               ora_counters* counters) {
    if (!prm || !cam || !fx || !sky_rgba || w <= 0 || h <= 0 || y0 < 0 || y1 > h || y0 > y1) return -1;
    ora_planes pl;
    memset(&pl, 0, sizeof pl);
    if (planes) pl = *planes;
    uint64_t c_steps = 0, c_disk = 0, c_dust = 0, c_dense = 0, c_cap = 0, c_esc = 0, c_exh = 0, c_touch = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : c_steps, c_disk, c_dust, c_dense, c_cap, c_esc, c_exh, c_touch)
    for (int y = y0; y < y1; ++y) {
        for (int x = 0; x < w; ++x) {
            pixel_out o;
            trace_pixel(prm, x, y, w, h, time, cam, fx, sky_rgba, sky_w, sky_h, &o);
            size_t idx = (size_t)y * w + x;
            if (pl.hdr) { float* d = pl.hdr + 4 * idx; d[0] = o.hdr[0]; d[1] = o.hdr[1]; d[2] = o.hdr[2]; d[3] = o.T; }
            if (pl.dir) { float* d = pl.dir + 4 * idx; d[0] = o.dir[0]; d[1] = o.dir[1]; d[2] = o.dir[2]; d[3] = 0; }
            if (pl.emis) { float* d = pl.emis + 4 * idx; d[0] = o.I[0]; d[1] = o.I[1]; d[2] = o.I[2]; d[3] = 0; }
            if (pl.pos) { float* d = pl.pos + 4 * idx; d[0] = o.p.x; d[1] = o.p.y; d[2] = o.p.z; d[3] = 0; }
            if (pl.vel) { float* d = pl.vel + 4 * idx; d[0] = o.v.x; d[1] = o.v.y; d[2] = o.v.z; d[3] = 0; }
            if (pl.cls) pl.cls[idx] = o.cls;
            if (pl.steps) pl.steps[idx] = o.steps;
            if (out_rgba) memcpy(out_rgba + 4 * ((size_t)(h - 1 - y) * w + x), o.rgba, 4);          /* :168 row flip */
            c_steps += (uint64_t)o.steps;
            c_disk += o.disk_evals; c_dust += o.dust_evals; c_dense += o.dense;
            int cap = (o.cls & ORA_CLS_MASK) == ORA_CLS_CAPTURED, exh = (o.cls & ORA_CLSF_EXHAUSTED) != 0;
            c_cap += (uint64_t)cap; c_exh += (uint64_t)exh; c_esc += (uint64_t)(!cap && !exh);
            c_touch += (uint64_t)((o.cls & ORA_CLSF_TOUCHED) != 0);
        }
    }
    if (counters) {
        counters->rk4_steps = c_steps; counters->disk_evals = c_disk; counters->dust_evals = c_dust;
        counters->dense_samples = c_dense; counters->n_captured = c_cap; counters->n_escaped = c_esc;
        counters->n_exhausted = c_exh; counters->n_touched = c_touch;
    }
    return 0;
}

static inline v3 ld3(const float* a) { return V(a[0], a[1], a[2]); }
static inline void st3(float* a, v3 v) { a[0] = v.x; a[1] = v.y; a[2] = v.z; }

void ora_geodesic_acc(const ora_params* prm, int n, const float* q, const float* v, float* out) {
    for (int i = 0; i < n; ++i) st3(out + 3 * i, geodesic_acc(prm, ld3(q + 3 * i), ld3(v + 3 * i)));
}
void ora_rk4_step(const ora_params* prm, int n, float* p, float* v, const float* h) {
    for (int i = 0; i < n; ++i) {
        v3 pp = ld3(p + 3 * i), vv = ld3(v + 3 * i);
        step_rk4(prm, &pp, &vv, h[i]);
        st3(p + 3 * i, pp); st3(v + 3 * i, vv);
    }
}
void ora_euler_step(const ora_params* prm, int n, float* p, float* v, const float* h) {
    for (int i = 0; i < n; ++i) {
        v3 pp = ld3(p + 3 * i), vv = ld3(v + 3 * i);
        step_euler(prm, &pp, &vv, h[i]);
        st3(p + 3 * i, pp); st3(v + 3 * i, vv);
    }
}
void ora_redshift(const ora_params* prm, int n, const float* q, const float* v, float* out) {
    for (int i = 0; i < n; ++i) out[i] = redshift_factor(prm, ld3(q + 3 * i), ld3(v + 3 * i));
}
static int g_probe_fm = 0;   /* contract of the parameter-less probes */
void ora_set_probe_contract(int fmad) { g_probe_fm = fmad != 0; }
void ora_hash31(int n, const float* p, float* out) { for (int i = 0; i < n; ++i) out[i] = hash31(g_probe_fm, ld3(p + 3 * i)); }
void ora_noise3d(int n, const float* p, float* out) { for (int i = 0; i < n; ++i) out[i] = noise3(g_probe_fm, ld3(p + 3 * i)); }
void ora_fbm(int n, const float* p, int octaves, float* out) { for (int i = 0; i < n; ++i) out[i] = fbm(g_probe_fm, ld3(p + 3 * i), octaves); }
void ora_disk_temperature(const ora_params* prm, int n, const float* r, float* out) {
    for (int i = 0; i < n; ++i) out[i] = disk_temperature(prm, r[i]);
}
void ora_disk_density(const ora_params* prm, int n, const float* q, float time, float* out) {
    for (int i = 0; i < n; ++i) out[i] = disk_density(prm, ld3(q + 3 * i), time);
}
void ora_dust_density(const ora_params* prm, int n, const float* q, float time, float* out) {
    for (int i = 0; i < n; ++i) out[i] = dust_density(prm, ld3(q + 3 * i), time);
}
void ora_tex2d(const uint8_t* sky_rgba, int sky_w, int sky_h, int n, const float* tx, const float* ty, float* out) {
    for (int i = 0; i < n; ++i) tex_emul_fetch(sky_rgba, sky_w, sky_h, tx[i], ty[i], out + 4 * i);
}
