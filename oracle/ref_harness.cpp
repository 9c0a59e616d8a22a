/*
 * ref_harness.cpp -- the REFERENCE checker ("kind": "reference").
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_abi.h).  Built by oracle/Makefile into
 * oracle/_ref/libref_host.so and only when /root/reference is present.
 *
 * The device math comes from the reference's own, unmodified headers, included from
 * $(REF)/include at build time (g++ sees __device__/__forceinline__ as empty through the
 * toolkit's host_defines.h):
 *     config.h  math_utils.h  geodesics.h  integrators.h  densities.h
 *     camera_effects/post_processing.h  raymarcher.h  camera_paths.h
 * and src/camera_paths.cpp is compiled next to this file from where it lies.
 * Only two things are restated here because they cannot be compiled on a host without
 * CUDA/GLFW: the body of __global__ raymarch_kernel (src/raymarcher.cu:15-174, it calls
 * tex2D) and the two camera helpers that live inside src/main.cpp (:141-167, :176-203).
 *
 * The reference bakes its tuning in as macros (include/config.h).  To run a=0 and a=0.99
 * (and any other parameter set) from one binary, the macros are re-pointed at a global
 * parameter block AFTER config.h has been read and BEFORE the headers whose function
 * bodies expand them.  Constant sub-expressions such as EVENT_HORIZON*1.01f are then
 * evaluated at run time in float, which is what the compilers' constant folding does too.
 *
 * Canonical rounding: build with -ffp-contract=off.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORA_PREFIX ref_
#include "oracle_abi.h"
#include "tex_emul.h"

/* ---- reference headers, verbatim ---- */
#include "config.h"

static ora_params g_prm;
/* defaults captured from the reference's own macros before they are re-pointed */
static const ora_params k_ref_defaults = {SPIN_A,          EVENT_HORIZON,  ISCO_RADIUS,      DISK_OUT_M,
                                          DISK_H_M,        DISK_LUMINOSITY, DISK_OPACITY,    EXPOSURE,
                                          CLOUD_H_M,       CLOUD_OUT_M,    CLOUD_OPACITY,    CLOUD_LUMINOSITY,
                                          STEP_SIZE_M,     DISK_TEMP_REF,  MAX_STEPS,        ORA_FLAG_DISK | ORA_FLAG_DUST};
#undef SPIN_A
#undef EVENT_HORIZON
#undef ISCO_RADIUS
#undef DISK_OUT_M
#undef DISK_H_M
#undef DISK_LUMINOSITY
#undef DISK_OPACITY
#undef EXPOSURE
#undef CLOUD_H_M
#undef CLOUD_OUT_M
#undef CLOUD_OPACITY
#undef CLOUD_LUMINOSITY
#undef STEP_SIZE_M
#undef DISK_TEMP_REF
#undef MAX_STEPS
#define SPIN_A (g_prm.spin_a)
#define EVENT_HORIZON (g_prm.event_horizon)
#define ISCO_RADIUS (g_prm.isco_radius)
#define DISK_OUT_M (g_prm.disk_out)
#define DISK_H_M (g_prm.disk_h)
#define DISK_LUMINOSITY (g_prm.disk_luminosity)
#define DISK_OPACITY (g_prm.disk_opacity)
#define EXPOSURE (g_prm.exposure)
#define CLOUD_H_M (g_prm.cloud_h)
#define CLOUD_OUT_M (g_prm.cloud_out)
#define CLOUD_OPACITY (g_prm.cloud_opacity)
#define CLOUD_LUMINOSITY (g_prm.cloud_luminosity)
#define STEP_SIZE_M (g_prm.step_size)
#define DISK_TEMP_REF (g_prm.disk_temp_ref)
#define MAX_STEPS (g_prm.max_steps)

#include "math_utils.h"
#include "densities.h"
#include "geodesics.h"
#include "integrators.h"
#include "camera_effects/post_processing.h"
#include "raymarcher.h"
#include "camera_paths.h"

static_assert(sizeof(CameraState) == sizeof(ora_camera), "CameraState layout");

namespace {

inline float3 f3(const float* a) { return make_float3(a[0], a[1], a[2]); }
inline void st3(float* a, float3 v) { a[0] = v.x; a[1] = v.y; a[2] = v.z; }
inline void st4(float* a, float x, float y, float z, float w) { a[0] = x; a[1] = y; a[2] = z; a[3] = w; }

struct PixelOut {
    float hdr[3], T, dir[3], I[3];
    float3 p, v;
    uint8_t cls;
    int steps;
    uint32_t disk_evals, dust_evals, dense;
    uint8_t rgba[4];
};

/* One thread of raymarch_kernel (src/raymarcher.cu:16-173), around the reference's own functions. */
void trace_pixel(int x, int y, int width, int height, float time, const CameraState& cam, const CameraEffects& fx,
                 const uint8_t* sky, int sky_w, int sky_h, PixelOut& o) {
    const bool want_disk = (g_prm.flags & ORA_FLAG_DISK) != 0;
    const bool want_dust = (g_prm.flags & ORA_FLAG_DUST) != 0;

    float2 uv = make_float2((float)x / width, (float)y / height);
    if (fx.useLensDistortion) uv = apply_lens_distortion(uv, fx.distortionAmount);

    float u_coord = uv.x * 2.0f - 1.0f;
    float v_coord = uv.y * 2.0f - 1.0f;
    float aspect = (float)width / height;
    u_coord *= aspect;

    float3 p = cam.pos;
    float3 vel = normalize(add(cam.forward, add(mul(cam.right, u_coord), mul(cam.up, v_coord))));

    float ir = 0, ig = 0, ib = 0, transmittance = 1.0f;
    bool hit_horizon = false, touched = false;
    int steps = 0;
    uint32_t n_disk = 0, n_dust = 0, n_dense = 0;
    int i = 0;
    const int max_steps = MAX_STEPS;
    for (; i < max_steps; i++) {
        float3 rel_p = sub(p, MASS_POS);
        float r2 = dot(rel_p, rel_p);
        float r = sqrtf(r2);
        if (r < EVENT_HORIZON * 1.01f) {
            hit_horizon = true;
            transmittance = 0.0f;
            break;
        }
        float current_h = STEP_SIZE_M;
        bool near_bh = (r < 18.0f);
        bool in_disk_zone = (fabsf(rel_p.y) < DISK_H_M * 5.0f && r < DISK_OUT_M + 5.0f);
        bool in_cloud_zone = (fabsf(rel_p.y) < CLOUD_H_M * 1.5f && r < CLOUD_OUT_M);
        if (near_bh) current_h *= 0.1f;
        else if (in_disk_zone) current_h *= 0.3f;
        else if (in_cloud_zone) current_h *= 0.5f;

        integrate_rk4(p, vel, current_h);
        ++steps;

        if (in_disk_zone || in_cloud_zone) {
            float d_disk = 0.0f, d_cloud = 0.0f;
            if (in_disk_zone && want_disk) { d_disk = getAccretionDensity(rel_p, time); ++n_disk; }
            if (in_cloud_zone && want_dust) { d_cloud = getDustCloudDensity(rel_p, time); ++n_dust; }
            if (d_disk > 0.001f || d_cloud > 0.001f) {
                touched = true;
                ++n_dense;
                float3 step_emit = make_float3(0, 0, 0);
                float step_opacity = 0;
                if (d_disk > 0.001f) {
                    float g = calculateRedshiftFactor(rel_p, vel);
                    float T = getDiskTemperature(r);
                    float T_norm = powf(T / DISK_TEMP_REF, 0.5f);
                    float bol_I = powf(g, 4.0f) * T_norm * d_disk * DISK_LUMINOSITY;
                    float color_t = g * powf(T / DISK_TEMP_REF, 0.4f) * 2.5f;
                    step_emit.x += 1.0f * bol_I;
                    step_emit.y += fminf(0.25f, 0.12f * color_t) * bol_I;
                    step_emit.z += fmaxf(0.0f, 0.01f * (color_t - 2.0f)) * bol_I;
                    step_opacity += d_disk * DISK_OPACITY;
                }
                if (d_cloud > 0.001f) {
                    float g = calculateRedshiftFactor(rel_p, vel);
                    float lighting = 0.5f + 3.0f * powf(ISCO_RADIUS / fmaxf(r, ISCO_RADIUS), 1.2f);
                    float cloud_I = d_cloud * CLOUD_LUMINOSITY * lighting;
                    float shift = smoothstep(0.7f, 1.3f, g);
                    float3 base_color = make_float3(0.60f, 0.65f, 0.80f);
                    step_emit.x += base_color.x * cloud_I * lerp(1.2f, 0.8f, shift);
                    step_emit.y += base_color.y * cloud_I * lerp(0.8f, 1.1f, shift);
                    step_emit.z += base_color.z * cloud_I * lerp(0.6f, 1.4f, shift);
                    step_opacity += d_cloud * CLOUD_OPACITY;
                }
                float d_tau = step_opacity * current_h;
                float step_trans = expf(-d_tau);
                float factor = (1.0f - step_trans) * transmittance;
                ir += step_emit.x * factor;
                ig += step_emit.y * factor;
                ib += step_emit.z * factor;
                transmittance *= step_trans;
            }
        }
        if (r > 250.0f && dot(rel_p, vel) > 0) break;
    }
    const bool exhausted = (i >= max_steps);

    float3 bg = make_float3(0, 0, 0);
    float3 d = make_float3(0, 0, 0);
    if (!hit_horizon) {
        d = normalize(vel);
        float offset = fx.useChromaticAberration ? fx.caAmount : 0.0f;
        float offs[3] = {offset, 0.0f, -offset};
        float tap[3][4];
        for (int c = 0; c < 3; ++c) {
            float phi = atan2f(d.z, d.x) + offs[c];
            float theta = asinf(d.y);
            float tx = 0.5f + phi / (2.0f * PI);
            float ty = 0.5f - theta / PI;
            tex_emul_fetch(sky, sky_w, sky_h, tx, ty, tap[c]);
        }
        bg = make_float3(tap[0][0], tap[1][1], tap[2][2]);
    }
    float3 hdr;
    hdr.x = ir + bg.x * transmittance;
    hdr.y = ig + bg.y * transmittance;
    hdr.z = ib + bg.z * transmittance;

    o.hdr[0] = hdr.x; o.hdr[1] = hdr.y; o.hdr[2] = hdr.z;
    o.T = transmittance;
    o.dir[0] = d.x; o.dir[1] = d.y; o.dir[2] = d.z;
    o.I[0] = ir; o.I[1] = ig; o.I[2] = ib;
    o.p = p; o.v = vel;
    o.steps = steps;
    o.disk_evals = n_disk; o.dust_evals = n_dust; o.dense = n_dense;
    uint8_t cls = hit_horizon ? ORA_CLS_CAPTURED : (touched ? ORA_CLS_DISK_HIT : ORA_CLS_ESCAPED);
    if (exhausted) cls |= ORA_CLSF_EXHAUSTED;
    if (touched) cls |= ORA_CLSF_TOUCHED;
    o.cls = cls;

    if (fx.useBloom) {
        float3 bloom = get_bloom_contribution(hdr, fx.bloomThreshold);
        hdr = add(hdr, mul(bloom, fx.bloomIntensity));
    }
    if (fx.useVignette) hdr = apply_vignette(hdr, uv, fx.vignetteIntensity);
    float out_r = 1.0f - expf(-hdr.x * EXPOSURE);
    float out_g = 1.0f - expf(-hdr.y * EXPOSURE);
    float out_b = 1.0f - expf(-hdr.z * EXPOSURE);
    o.rgba[0] = (unsigned char)(out_r * 255);
    o.rgba[1] = (unsigned char)(out_g * 255);
    o.rgba[2] = (unsigned char)(out_b * 255);
    o.rgba[3] = 255;
}

CameraEffects to_ref_fx(const ora_effects* fx) {
    CameraEffects e;
    e.useBloom = fx->use_bloom != 0;
    e.bloomThreshold = fx->bloom_threshold;
    e.bloomIntensity = fx->bloom_intensity;
    e.useVignette = fx->use_vignette != 0;
    e.vignetteIntensity = fx->vignette_intensity;
    e.useChromaticAberration = fx->use_ca != 0;
    e.caAmount = fx->ca_amount;
    e.useLensDistortion = fx->use_lens != 0;
    e.distortionAmount = fx->distortion_amount;
    return e;
}

/* CameraController::getCUDAStateFrom, src/main.cpp:141-167 (lives in main.cpp, which needs GLFW). */
CameraState camera_from(float3 pos, float yaw, float pitch) {
    float radYaw = yaw * 3.14159f / 180.0f;
    float radPitch = pitch * 3.14159f / 180.0f;
    float3 forward;
    forward.x = std::sin(radYaw) * std::cos(radPitch);
    forward.y = std::sin(radPitch);
    forward.z = std::cos(radYaw) * std::cos(radPitch);
    float mag = std::sqrt(forward.x * forward.x + forward.y * forward.y + forward.z * forward.z);
    forward.x /= mag; forward.y /= mag; forward.z /= mag;
    float3 worldUp = {0.0f, 1.0f, 0.0f};
    float3 right;
    right.x = worldUp.y * forward.z - worldUp.z * forward.y;
    right.y = worldUp.z * forward.x - worldUp.x * forward.z;
    right.z = worldUp.x * forward.y - worldUp.y * forward.x;
    float rMag = std::sqrt(right.x * right.x + right.y * right.y + right.z * right.z);
    right.x /= rMag; right.y /= rMag; right.z /= rMag;
    float3 up;
    up.x = forward.y * right.z - forward.z * right.y;
    up.y = forward.z * right.x - forward.x * right.z;
    up.z = forward.x * right.y - forward.y * right.x;
    return {pos, forward, right, up};
}

}  // namespace

extern "C" {

void ref_default_params(ora_params* out) { *out = k_ref_defaults; }

void ref_default_effects(ora_effects* out) {
    CameraEffects e; /* default member initialisers, camera_settings.h:5-16 */
    out->use_bloom = e.useBloom;
    out->bloom_threshold = e.bloomThreshold;
    out->bloom_intensity = e.bloomIntensity;
    out->use_vignette = e.useVignette;
    out->vignette_intensity = e.vignetteIntensity;
    out->use_ca = e.useChromaticAberration;
    out->ca_amount = e.caAmount;
    out->use_lens = e.useLensDistortion;
    out->distortion_amount = e.distortionAmount;
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* Size of the OpenMP team of the next render (launchers such as torchrun export OMP_NUM_THREADS=1). */
void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void ref_camera_from(const float pos[3], float yaw_deg, float pitch_deg, ora_camera* out) {
    CameraState c = camera_from(f3(pos), yaw_deg, pitch_deg);
    std::memcpy(out, &c, sizeof(c));
}

/* PathController::getInterpolatedState, src/main.cpp:176-203, using the reference's own
 * catmull_rom / lerp_angle / initDefaultPaths (src/camera_paths.cpp, compiled as-is). */
int ref_path_state(int path_index, float t, ora_camera* out, float pyp[5]) {
    static bool inited = false;
    if (!inited) { initDefaultPaths(); inited = true; }
    const CameraPath* path = PathManager::instance().getPath(path_index);
    if (!path || path->keyframes.empty()) return -1;
    const auto& keys = path->keyframes;
    float3 pos; float yaw, pitch;
    bool found = false;
    if (t <= keys.front().time) { pos = keys.front().pos; yaw = keys.front().yaw; pitch = keys.front().pitch; found = true; }
    else if (t >= keys.back().time) { pos = keys.back().pos; yaw = keys.back().yaw; pitch = keys.back().pitch; found = true; }
    else {
        for (size_t i = 0; i < keys.size() - 1; ++i) {
            if (t >= keys[i].time && t <= keys[i + 1].time) {
                float factor = (t - keys[i].time) / (keys[i + 1].time - keys[i].time);
                int i0 = std::max(0, (int)i - 1);
                int i1 = (int)i;
                int i2 = (int)i + 1;
                int i3 = std::min((int)keys.size() - 1, (int)i + 2);
                pos = catmull_rom(keys[i0].pos, keys[i1].pos, keys[i2].pos, keys[i3].pos, factor);
                yaw = lerp_angle(keys[i1].yaw, keys[i2].yaw, factor);
                pitch = lerp_angle(keys[i1].pitch, keys[i2].pitch, factor);
                found = true;
                break;
            }
        }
    }
    if (!found) return -2;
    CameraState c = camera_from(pos, yaw, pitch);
    std::memcpy(out, &c, sizeof(c));
    if (pyp) { pyp[0] = pos.x; pyp[1] = pos.y; pyp[2] = pos.z; pyp[3] = yaw; pyp[4] = pitch; }
    return 0;
}

int ref_render(const ora_params* prm, const ora_camera* cam, const ora_effects* fx, const uint8_t* sky_rgba, int sky_w,
               int sky_h, float time, int w, int h, int y0, int y1, uint8_t* out_rgba, const ora_planes* planes,
               ora_counters* counters) {
    if (!prm || !cam || !fx || !sky_rgba || w <= 0 || h <= 0 || y0 < 0 || y1 > h || y0 > y1) return -1;
    g_prm = *prm;
    CameraState cs;
    std::memcpy(&cs, cam, sizeof(cs));
    CameraEffects e = to_ref_fx(fx);
    ora_planes pl;
    std::memset(&pl, 0, sizeof(pl));
    if (planes) pl = *planes;
    uint64_t c_steps = 0, c_disk = 0, c_dust = 0, c_dense = 0, c_cap = 0, c_esc = 0, c_exh = 0, c_touch = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : c_steps, c_disk, c_dust, c_dense, c_cap, c_esc, c_exh, c_touch)
    for (int y = y0; y < y1; ++y) {
        for (int x = 0; x < w; ++x) {
            PixelOut o;
            trace_pixel(x, y, w, h, time, cs, e, sky_rgba, sky_w, sky_h, o);
            size_t idx = (size_t)y * w + x;
            if (pl.hdr) st4(pl.hdr + 4 * idx, o.hdr[0], o.hdr[1], o.hdr[2], o.T);
            if (pl.dir) st4(pl.dir + 4 * idx, o.dir[0], o.dir[1], o.dir[2], 0.0f);
            if (pl.emis) st4(pl.emis + 4 * idx, o.I[0], o.I[1], o.I[2], 0.0f);
            if (pl.pos) st4(pl.pos + 4 * idx, o.p.x, o.p.y, o.p.z, 0.0f);
            if (pl.vel) st4(pl.vel + 4 * idx, o.v.x, o.v.y, o.v.z, 0.0f);
            if (pl.cls) pl.cls[idx] = o.cls;
            if (pl.steps) pl.steps[idx] = o.steps;
            if (out_rgba) std::memcpy(out_rgba + 4 * ((size_t)(h - 1 - y) * w + x), o.rgba, 4);
            c_steps += (uint64_t)o.steps;
            c_disk += o.disk_evals;
            c_dust += o.dust_evals;
            c_dense += o.dense;
            const bool cap = (o.cls & ORA_CLS_MASK) == ORA_CLS_CAPTURED, exh = (o.cls & ORA_CLSF_EXHAUSTED) != 0;
            c_cap += cap;
            c_exh += exh;
            c_esc += (!cap && !exh);
            c_touch += (o.cls & ORA_CLSF_TOUCHED) != 0;
        }
    }
    if (counters) {
        counters->rk4_steps = c_steps; counters->disk_evals = c_disk; counters->dust_evals = c_dust;
        counters->dense_samples = c_dense; counters->n_captured = c_cap; counters->n_escaped = c_esc;
        counters->n_exhausted = c_exh; counters->n_touched = c_touch;
    }
    return 0;
}

void ref_geodesic_acc(const ora_params* prm, int n, const float* q, const float* v, float* out) {
    g_prm = *prm;
    for (int i = 0; i < n; ++i) st3(out + 3 * i, getGeodesicAcc(f3(q + 3 * i), f3(v + 3 * i)));
}
void ref_rk4_step(const ora_params* prm, int n, float* p, float* v, const float* h) {
    g_prm = *prm;
    for (int i = 0; i < n; ++i) {
        float3 pp = f3(p + 3 * i), vv = f3(v + 3 * i);
        integrate_rk4(pp, vv, h[i]);
        st3(p + 3 * i, pp); st3(v + 3 * i, vv);
    }
}
void ref_euler_step(const ora_params* prm, int n, float* p, float* v, const float* h) {
    g_prm = *prm;
    for (int i = 0; i < n; ++i) {
        float3 pp = f3(p + 3 * i), vv = f3(v + 3 * i);
        integrate_euler(pp, vv, h[i]);
        st3(p + 3 * i, pp); st3(v + 3 * i, vv);
    }
}
void ref_redshift(const ora_params* prm, int n, const float* q, const float* v, float* out) {
    g_prm = *prm;
    for (int i = 0; i < n; ++i) out[i] = calculateRedshiftFactor(f3(q + 3 * i), f3(v + 3 * i));
}
void ref_hash31(int n, const float* p, float* out) { for (int i = 0; i < n; ++i) out[i] = hash31(f3(p + 3 * i)); }
void ref_noise3d(int n, const float* p, float* out) { for (int i = 0; i < n; ++i) out[i] = noise3D(f3(p + 3 * i)); }
void ref_fbm(int n, const float* p, int octaves, float* out) { for (int i = 0; i < n; ++i) out[i] = fbm(f3(p + 3 * i), octaves); }
void ref_disk_temperature(const ora_params* prm, int n, const float* r, float* out) {
    g_prm = *prm;
    for (int i = 0; i < n; ++i) out[i] = getDiskTemperature(r[i]);
}
void ref_disk_density(const ora_params* prm, int n, const float* q, float time, float* out) {
    g_prm = *prm;
    for (int i = 0; i < n; ++i) out[i] = getAccretionDensity(f3(q + 3 * i), time);
}
void ref_dust_density(const ora_params* prm, int n, const float* q, float time, float* out) {
    g_prm = *prm;
    for (int i = 0; i < n; ++i) out[i] = getDustCloudDensity(f3(q + 3 * i), time);
}
void ref_tex2d(const uint8_t* sky_rgba, int sky_w, int sky_h, int n, const float* tx, const float* ty, float* out) {
    for (int i = 0; i < n; ++i) tex_emul_fetch(sky_rgba, sky_w, sky_h, tx[i], ty[i], out + 4 * i);
}

}  // extern "C"
