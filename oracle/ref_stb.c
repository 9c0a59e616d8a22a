/* ref_stb.c -- stb_image (v2.30, the copy vendored in the reference tree: include/stb_image.h, which src/main.cpp:240
 * calls as stbi_load(path, &w, &h, &c, 4)) compiled where it lies into oracle/_ref/libref_stb.so: the checker for the
 * product's own decoder (csrc/rrt_image.cpp).  TEST INFRASTRUCTURE ONLY; nothing of stb_image is copied into the repo. */
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"

unsigned char* refstb_load(const char* path, int* w, int* h, int* comp) { return stbi_load(path, w, h, comp, 4); }
void refstb_free(unsigned char* p) { stbi_image_free(p); }
const char* refstb_failure(void) { return stbi_failure_reason(); }
