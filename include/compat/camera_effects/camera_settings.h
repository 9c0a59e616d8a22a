// Drop-in for the reference's include/camera_effects/camera_settings.h (struct CameraEffects, :4-17).
// Field names, order, types and default values must match the reference so that src/main.cpp
// (g_Effects, key bindings B/V/L/C at :286-301) compiles and behaves unchanged against this tree.
// Layout: 36 bytes {bool@0, float@4, float@8, bool@12, float@16, bool@20, float@24, bool@28, float@32}.
#ifndef CAMERA_SETTINGS_H
#define CAMERA_SETTINGS_H

struct CameraEffects {
    // bloom gate (post_processing.h:27-31)
    bool useBloom = true;
    float bloomThreshold = 0.8f;
    float bloomIntensity = 0.5f;
    // radial darkening (post_processing.h:13-17)
    bool useVignette = true;
    float vignetteIntensity = 0.4f;
    // per-channel azimuth offset of the sky taps (raymarcher.cu:132-145)
    bool useChromaticAberration = false;
    float caAmount = 0.005f;
    // barrel distortion of the image-plane coordinate (post_processing.h:19-24)
    bool useLensDistortion = true;
    float distortionAmount = 0.15f;
};

static_assert(sizeof(CameraEffects) == 36, "CameraEffects must keep the reference's 36-byte layout");

#endif
