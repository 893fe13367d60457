// Stub of OUR OWN (not reference code, not glm): the few glm types and functions the reference headers and translation units on this
// path mention -- glm::ivec3 (CUDAQuickSurf.h), glm::vec2 / vec3 / vec4 with the arithmetic protein_calls::ProteinColor uses
// (MakeWeightedColorTable, InterpolateMultipleColors).  glm is a vcpkg dependency of MegaMol that this image does not have; a real
// MegaMol build uses the real library.  Test infrastructure only (oracle/Makefile.ref).
#pragma once
#include <cmath>
namespace glm {
struct ivec3 {
    int x, y, z;
    ivec3(int a = 0, int b = 0, int c = 0) : x(a), y(b), z(c) {}
};
struct vec2 {
    union { float x, r; };
    union { float y, g; };
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float a, float b) : x(a), y(b) {}
};
struct vec3 {
    union { float x, r; };
    union { float y, g; };
    union { float z, b; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    template<class A, class B, class C> vec3(A a, B b_, C c) : x(static_cast<float>(a)), y(static_cast<float>(b_)), z(static_cast<float>(c)) {}
    vec3& operator+=(const vec3& o) { x += o.x, y += o.y, z += o.z; return *this; }
    vec3& operator-=(const vec3& o) { x -= o.x, y -= o.y, z -= o.z; return *this; }
    vec3& operator*=(float s) { x *= s, y *= s, z *= s; return *this; }
    vec3& operator/=(float s) { x /= s, y /= s, z /= s; return *this; }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4 {
    union { float x, r; };
    union { float y, g; };
    union { float z, b; };
    union { float w, a; };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
};
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 make_vec3(const float* p) { return vec3(p[0], p[1], p[2]); }
inline vec3 make_vec3(const unsigned char* p) { return vec3(p[0], p[1], p[2]); }
template<class T> inline T clamp(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline vec3 mix(const vec3& a, const vec3& b, float t) { return a * (1.0f - t) + b * t; } // glm: x * (1 - a) + y * a
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float length(const vec3& a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
} // namespace glm
