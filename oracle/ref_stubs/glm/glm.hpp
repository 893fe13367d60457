// Stub of OUR OWN (not reference code): the one glm type the reference's CUDAQuickSurf.h mentions (getMapSize, never called by the
// harness).  glm is a vcpkg dependency of MegaMol that this image does not have.
#pragma once
namespace glm {
struct ivec3 {
    int x, y, z;
    ivec3(int a = 0, int b = 0, int c = 0) : x(a), y(b), z(c) {}
};
} // namespace glm
