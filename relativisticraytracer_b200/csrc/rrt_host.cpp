// rrt_host.cpp -- host-side camera math and keyframe paths behind the C ABI (include/rrt.h).
//
// This is the step immediately before the hot path: it turns (position, yaw, pitch) or a path time
// into the CameraState the render kernel consumes.  It replaces, for headless use, the pieces of the
// reference that live in its GLFW application: CameraController::getCUDAStateFrom (src/main.cpp:141-167),
// PathController::getInterpolatedState (src/main.cpp:176-203) and the spline/angle helpers and keyframe
// tables of src/camera_paths.cpp.  Built with -ffp-contract=off: every operation is one binary32 op in
// the reference's order, so the basis vectors are bit-identical to the reference's on the same libm.
#include <cmath>
#include <cstring>

#include "../../include/rrt.h"

namespace {

struct Vec { float x, y, z; };
struct Key { float t; Vec pos; float yaw, pitch; };
struct Path { const char* name; const Key* keys; int n; };

// src/camera_paths.cpp:33-72
const Key kGargantua[] = {
    {0.0f, {0.0f, 15.0f, -80.0f}, 0.0f, -10.6f},
    {6.0f, {15.0f, 3.0f, -30.0f}, -26.6f, -5.1f},
    {12.0f, {35.0f, 0.8f, 10.0f}, -106.0f, -1.2f},
    {18.0f, {5.0f, 1.5f, 50.0f}, -174.3f, -1.7f},
    {25.0f, {-20.0f, 12.0f, 70.0f}, -196.0f, -9.3f},
};
const Key kHorizonFocus[] = {
    {0.0f, {40.0f, 2.0f, 0.0f}, -90.0f, 0.0f},
    {8.0f, {0.0f, 5.0f, 40.0f}, -180.0f, -5.0f},
    {16.0f, {-40.0f, 2.0f, 0.0f}, -270.0f, 0.0f},
    {24.0f, {0.0f, -5.0f, -40.0f}, -360.0f, 5.0f},
    {32.0f, {40.0f, 2.0f, 0.0f}, -450.0f, 0.0f},
};
const Key kSkimmer[] = {
    {0.0f, {0.0f, 10.0f, -60.0f}, 0.0f, -9.5f},
    {8.0f, {15.0f, 2.0f, -15.0f}, -45.0f, -4.7f},
    {14.0f, {4.2f, 0.6f, 4.2f}, -90.0f, -5.7f},
    {20.0f, {-20.0f, 8.0f, -20.0f}, -225.0f, -20.0f},
    {26.0f, {-20.0f, 8.0f, -20.0f}, 20.0f, -10.0f},
    {29.0f, {-30.0f, 2.0f, -30.0f}, 45.0f, -2.7f},
};
const Path kPaths[] = {
    {"Gargantua Fly-By", kGargantua, 5},
    {"Event Horizon Focus", kHorizonFocus, 5},
    {"Horizon Skimmer", kSkimmer, 6},
};
constexpr int kNumPaths = 3;

// uniform Catmull-Rom, one coordinate (src/camera_paths.cpp:10-15)
float spline(float a, float b, float c, float d, float t, float t2, float t3) {
    return 0.5f * ((2.0f * b) + (-a + c) * t + (2.0f * a - 5.0f * b + 4.0f * c - d) * t2 +
                   (-a + 3.0f * b - 3.0f * c + d) * t3);
}

// shortest-arc interpolation in degrees (src/camera_paths.cpp:25-29)
float mix_angle(float a, float b, float t) {
    float diff = std::fmod(b - a + 180.0f, 360.0f) - 180.0f;
    if (diff < -180.0f) diff += 360.0f;
    return a + diff * t;
}

void basis_from(Vec pos, float yaw, float pitch, rrt_camera* out) {
    const float ry = yaw * 3.14159f / 180.0f;  // the reference's 5-digit pi (src/main.cpp:142-143)
    const float rp = pitch * 3.14159f / 180.0f;
    Vec f = {std::sin(ry) * std::cos(rp), std::sin(rp), std::cos(ry) * std::cos(rp)};
    const float fm = std::sqrt(f.x * f.x + f.y * f.y + f.z * f.z);
    f.x /= fm; f.y /= fm; f.z /= fm;
    const Vec wu = {0.0f, 1.0f, 0.0f};
    Vec r = {wu.y * f.z - wu.z * f.y, wu.z * f.x - wu.x * f.z, wu.x * f.y - wu.y * f.x};
    const float rm = std::sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
    r.x /= rm; r.y /= rm; r.z /= rm;
    const Vec u = {f.y * r.z - f.z * r.y, f.z * r.x - f.x * r.z, f.x * r.y - f.y * r.x};  // not renormalised
    out->pos[0] = pos.x; out->pos[1] = pos.y; out->pos[2] = pos.z;
    out->forward[0] = f.x; out->forward[1] = f.y; out->forward[2] = f.z;
    out->right[0] = r.x; out->right[1] = r.y; out->right[2] = r.z;
    out->up[0] = u.x; out->up[1] = u.y; out->up[2] = u.z;
}

}  // namespace

extern "C" {

void rrt_camera_from(const float pos[3], float yaw_deg, float pitch_deg, rrt_camera* out) {
    if (!pos || !out) return;
    basis_from(Vec{pos[0], pos[1], pos[2]}, yaw_deg, pitch_deg, out);
}

int rrt_path_count(void) { return kNumPaths; }
const char* rrt_path_name(int i) { return (i >= 0 && i < kNumPaths) ? kPaths[i].name : nullptr; }
int rrt_path_num_keys(int i) { return (i >= 0 && i < kNumPaths) ? kPaths[i].n : RRT_ERR_BAD_ARG; }
float rrt_path_duration(int i) { return (i >= 0 && i < kNumPaths) ? kPaths[i].keys[kPaths[i].n - 1].t : -1.0f; }

int rrt_path_state(int path_index, float t, rrt_camera* out, float pyp[5]) {
    if (path_index < 0 || path_index >= kNumPaths || !out) return RRT_ERR_BAD_ARG;
    const Key* k = kPaths[path_index].keys;
    const int n = kPaths[path_index].n;
    Vec pos = k[0].pos;
    float yaw = k[0].yaw, pitch = k[0].pitch;
    if (t <= k[0].t) {
        // clamp to the first key (src/main.cpp:183)
    } else if (t >= k[n - 1].t) {  // clamp to the last key (:184)
        pos = k[n - 1].pos; yaw = k[n - 1].yaw; pitch = k[n - 1].pitch;
    } else {
        int seg = -1;
        for (int i = 0; i + 1 < n; ++i)
            if (t >= k[i].t && t <= k[i + 1].t) { seg = i; break; }  // first matching segment (:186-187)
        if (seg < 0) return RRT_ERR_BAD_ARG;
        const float s = (t - k[seg].t) / (k[seg + 1].t - k[seg].t);
        const int a = seg > 0 ? seg - 1 : 0, d = seg + 2 < n ? seg + 2 : n - 1;  // neighbour clamping (:190-193)
        const float s2 = s * s, s3 = s2 * s;
        pos.x = spline(k[a].pos.x, k[seg].pos.x, k[seg + 1].pos.x, k[d].pos.x, s, s2, s3);
        pos.y = spline(k[a].pos.y, k[seg].pos.y, k[seg + 1].pos.y, k[d].pos.y, s, s2, s3);
        pos.z = spline(k[a].pos.z, k[seg].pos.z, k[seg + 1].pos.z, k[d].pos.z, s, s2, s3);
        yaw = mix_angle(k[seg].yaw, k[seg + 1].yaw, s);
        pitch = mix_angle(k[seg].pitch, k[seg + 1].pitch, s);
    }
    basis_from(pos, yaw, pitch, out);
    if (pyp) { pyp[0] = pos.x; pyp[1] = pos.y; pyp[2] = pos.z; pyp[3] = yaw; pyp[4] = pitch; }
    return RRT_OK;
}

float rrt_path_clock(int frame, float fps) {
    // recording mode advances a float clock by dt = 1.0f / RECORDING_FPS every frame (src/main.cpp:511-516)
    volatile float t = 0.0f;
    const float dt = 1.0f / fps;
    for (int i = 0; i < frame; ++i) t = t + dt;
    return t;
}

}  // extern "C"
