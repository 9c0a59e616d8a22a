/*
 * oracle_abi.h -- C ABI shared by the two CPU checkers in this directory.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load these libraries, and only as the checker.
 *
 * Two libraries export this ABI:
 *   - oracle/_ref/libref_host.so   (prefix ref_)  built by oracle/Makefile from
 *     ref_harness.cpp, which #includes the UNMODIFIED reference headers from
 *     /root/reference/include at build time ("kind": "reference").
 *   - oracle/librrt_oracle.so      (prefix ora_)  built from rrt_oracle.c, a
 *     plain-C restatement of the same algorithm ("kind": "port"), pinned
 *     bit-for-bit against the former (tests/test_oracle_vs_ref.py) and against
 *     the golden vectors committed under tests/golden/.
 *
 * Every function is a pure function of its arguments; arrays are host memory.
 */
#ifndef RRT_ORACLE_ABI_H
#define RRT_ORACLE_ABI_H

#include <stdint.h>

#ifndef ORA_PREFIX
#define ORA_PREFIX ora_
#endif
#define ORA_CAT2(a, b) a##b
#define ORA_CAT(a, b) ORA_CAT2(a, b)
#define ORA_FN(name) ORA_CAT(ORA_PREFIX, name)

#ifdef __cplusplus
extern "C" {
#endif

/* Runtime copy of the reference's compile-time macros (include/config.h). */
typedef struct ora_params {
    float spin_a;           /* SPIN_A            config.h:21 */
    float event_horizon;    /* EVENT_HORIZON     config.h:29 */
    float isco_radius;      /* ISCO_RADIUS       config.h:33 */
    float disk_out;         /* DISK_OUT_M        config.h:34 */
    float disk_h;           /* DISK_H_M          config.h:35 */
    float disk_luminosity;  /* DISK_LUMINOSITY   config.h:36 */
    float disk_opacity;     /* DISK_OPACITY      config.h:37 */
    float exposure;         /* EXPOSURE          config.h:38 */
    float cloud_h;          /* CLOUD_H_M         config.h:41 */
    float cloud_out;        /* CLOUD_OUT_M       config.h:42 */
    float cloud_opacity;    /* CLOUD_OPACITY     config.h:43 */
    float cloud_luminosity; /* CLOUD_LUMINOSITY  config.h:44 */
    float step_size;        /* STEP_SIZE_M       config.h:47 */
    float disk_temp_ref;    /* DISK_TEMP_REF     config.h:18 */
    int32_t max_steps;      /* MAX_STEPS         config.h:48 */
    uint32_t flags;         /* ORA_FLAG_* below */
} ora_params;

#define ORA_FLAG_DISK 1u  /* evaluate getAccretionDensity in the disk zone  (raymarcher.cu:68) */
#define ORA_FLAG_DUST 2u  /* evaluate getDustCloudDensity in the cloud zone (raymarcher.cu:69) */
/* PORT ONLY (rrt_oracle.c): evaluate with the FMA fusion schedule of the reference's own CUDA build (nvcc default
 * -fmad=true, read off the SASS of raymarch_kernel) instead of the unfused host arithmetic -- the CPU twin of
 * the product's RRT_FLAG_FMAD contract.  The reference-header build (ref_harness.cpp) ignores the flag: g++
 * compiles those headers with -ffp-contract=off. */
#define ORA_FLAG_FMAD 4u

/* struct CameraState, include/raymarcher.h:11-16 (four packed float3). */
typedef struct ora_camera {
    float pos[3], forward[3], right[3], up[3];
} ora_camera;

/* struct CameraEffects, include/camera_effects/camera_settings.h:4-17, as plain C. */
typedef struct ora_effects {
    int32_t use_bloom;
    float bloom_threshold, bloom_intensity;
    int32_t use_vignette;
    float vignette_intensity;
    int32_t use_ca;
    float ca_amount;
    int32_t use_lens;
    float distortion_amount;
} ora_effects;

/* Per-pixel planes, indexed [y*w + x] (NOT row-flipped).  Any pointer may be NULL. */
typedef struct ora_planes {
    float* hdr;     /* [n][4] final_hdr.rgb before camera effects (raymarcher.cu:148-150), w = transmittance */
    float* dir;     /* [n][4] normalize(vel) at exit (raymarcher.cu:129); zeros for captured rays */
    float* emis;    /* [n][4] accumulated emission I.rgb (raymarcher.cu:111-113) */
    float* pos;     /* [n][4] final p */
    float* vel;     /* [n][4] final (un-normalised) vel */
    uint8_t* cls;   /* [n] ORA_CLS_* | ORA_CLSF_* */
    int32_t* steps; /* [n] number of integrate_rk4 calls */
} ora_planes;

/* termination class graded by north_star: captured / disk-hit / escaped */
#define ORA_CLS_CAPTURED 0u /* hit_horizon (raymarcher.cu:47-51) */
#define ORA_CLS_DISK_HIT 1u /* not captured, at least one sample passed the >0.001 gate (raymarcher.cu:71) */
#define ORA_CLS_ESCAPED 2u  /* not captured, no medium touched */
#define ORA_CLS_MASK 3u
#define ORA_CLSF_EXHAUSTED 4u /* loop ran MAX_STEPS iterations (raymarcher.cu:41) */
#define ORA_CLSF_TOUCHED 8u   /* a sample passed the gate (also set on captured rays) */

typedef struct ora_counters {
    uint64_t rk4_steps;     /* sum of integrate_rk4 calls */
    uint64_t disk_evals;    /* getAccretionDensity calls (disk zone) */
    uint64_t dust_evals;    /* getDustCloudDensity calls (cloud zone) */
    uint64_t dense_samples; /* samples past the 0.001 gate */
    uint64_t n_captured, n_escaped, n_exhausted, n_touched;
} ora_counters;

void ORA_FN(default_params)(ora_params* out);
void ORA_FN(default_effects)(ora_effects* out);
int ORA_FN(num_threads)(void);
void ORA_FN(set_num_threads)(int n);

/* CameraController::getCUDAStateFrom, src/main.cpp:141-167 (angles in degrees). */
void ORA_FN(camera_from)(const float pos[3], float yaw_deg, float pitch_deg, ora_camera* out);
/* PathController::getInterpolatedState, src/main.cpp:176-203, over initDefaultPaths (camera_paths.cpp:31-73).
 * Returns 0, or -1 for a bad path index. Also returns pos/yaw/pitch when non-NULL. */
int ORA_FN(path_state)(int path_index, float t, ora_camera* out, float pos_yaw_pitch[5]);

/* Whole-frame render of rows [y0,y1): raymarch_kernel, src/raymarcher.cu:15-174.
 * out_rgba (may be NULL) is the full w*h uchar4 frame, row-flipped like the reference store (:168).
 * sky_rgba is an RGBA8 equirect map sampled with an emulation of the reference's texture object
 * (src/main.cpp:246-263: wrap-x, clamp-y, linear, normalised float). */
int ORA_FN(render)(const ora_params* prm, const ora_camera* cam, const ora_effects* fx, const uint8_t* sky_rgba,
                   int sky_w, int sky_h, float time, int w, int h, int y0, int y1, uint8_t* out_rgba,
                   const ora_planes* planes, ora_counters* counters);

/* Function-level probes (arrays of float3 are packed xyz). */
void ORA_FN(geodesic_acc)(const ora_params* prm, int n, const float* q, const float* v, float* out);
void ORA_FN(rk4_step)(const ora_params* prm, int n, float* p, float* v, const float* h);
void ORA_FN(euler_step)(const ora_params* prm, int n, float* p, float* v, const float* h);
void ORA_FN(redshift)(const ora_params* prm, int n, const float* q, const float* v, float* out);
void ORA_FN(hash31)(int n, const float* p, float* out);
void ORA_FN(noise3d)(int n, const float* p, float* out);
void ORA_FN(fbm)(int n, const float* p, int octaves, float* out);
void ORA_FN(disk_temperature)(const ora_params* prm, int n, const float* r, float* out);
void ORA_FN(disk_density)(const ora_params* prm, int n, const float* q, float time, float* out);
void ORA_FN(dust_density)(const ora_params* prm, int n, const float* q, float time, float* out);
/* texture fetch emulation: out[n][4] in [0,1] */
void ORA_FN(tex2d)(const uint8_t* sky_rgba, int sky_w, int sky_h, int n, const float* tx, const float* ty,
                   float* out);

#ifdef __cplusplus
}
#endif
#endif
