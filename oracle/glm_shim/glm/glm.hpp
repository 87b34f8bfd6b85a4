// Minimal stand-in for the subset of GLM that the vendored Inria rasterizer uses
// (vec3, vec4, column-major mat3 and a handful of free functions).
//
// TEST INFRASTRUCTURE ONLY.  The reference tree ships without third_party/glm
// (SURVEY.md section 0, fact 2), so the comparator build under oracle/_ref needs
// these types from somewhere.  The arithmetic below is written in the same
// expression order GLM's generic (non-SIMD) code path uses, so that nvcc's
// default FMA contraction produces the same rounding as a build against real
// GLM would: mat3*mat3 entries are  a*b + c*d + e*f  evaluated left to right,
// dot(a,b) is  (a.x*b.x + a.y*b.y) + a.z*b.z.
//
// Nothing under ocrfdet_b200/ includes this header.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GLMS_FN __host__ __device__ inline
#else
#define GLMS_FN inline
#endif

namespace glm {

struct vec3 {
  float x, y, z;
  GLMS_FN vec3() : x(0.f), y(0.f), z(0.f) {}
  GLMS_FN vec3(float s) : x(s), y(s), z(s) {}
  GLMS_FN vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  GLMS_FN float& operator[](int i) { return (&x)[i]; }
  GLMS_FN const float& operator[](int i) const { return (&x)[i]; }
  GLMS_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  GLMS_FN vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  GLMS_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
  float x, y, z, w;
  GLMS_FN vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
  GLMS_FN vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  GLMS_FN float& operator[](int i) { return (&x)[i]; }
  GLMS_FN const float& operator[](int i) const { return (&x)[i]; }
};

GLMS_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLMS_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLMS_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLMS_FN vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLMS_FN vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLMS_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLMS_FN vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }

GLMS_FN float dot(const vec3& a, const vec3& b) {
  vec3 t(a * b);
  return t.x + t.y + t.z;
}
GLMS_FN float length(const vec3& a) { return sqrtf(dot(a, a)); }
GLMS_FN vec3 max(const vec3& a, float s) {
  return vec3(a.x < s ? s : a.x, a.y < s ? s : a.y, a.z < s ? s : a.z);
}

// Column-major 3x3: m[c] is column c, m[c][r] is (row r, column c).
struct mat3 {
  vec3 col[3];
  GLMS_FN mat3() {}
  GLMS_FN mat3(float s) {
    col[0] = vec3(s, 0.f, 0.f);
    col[1] = vec3(0.f, s, 0.f);
    col[2] = vec3(0.f, 0.f, s);
  }
  GLMS_FN mat3(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2) {
    col[0] = vec3(x0, y0, z0);
    col[1] = vec3(x1, y1, z1);
    col[2] = vec3(x2, y2, z2);
  }
  GLMS_FN mat3(const vec3& a, const vec3& b, const vec3& c) {
    col[0] = a; col[1] = b; col[2] = c;
  }
  GLMS_FN vec3& operator[](int i) { return col[i]; }
  GLMS_FN const vec3& operator[](int i) const { return col[i]; }
};

GLMS_FN mat3 operator*(const mat3& m1, const mat3& m2) {
  const float A00 = m1[0][0], A01 = m1[0][1], A02 = m1[0][2];
  const float A10 = m1[1][0], A11 = m1[1][1], A12 = m1[1][2];
  const float A20 = m1[2][0], A21 = m1[2][1], A22 = m1[2][2];
  const float B00 = m2[0][0], B01 = m2[0][1], B02 = m2[0][2];
  const float B10 = m2[1][0], B11 = m2[1][1], B12 = m2[1][2];
  const float B20 = m2[2][0], B21 = m2[2][1], B22 = m2[2][2];
  mat3 r;
  r[0][0] = A00 * B00 + A10 * B01 + A20 * B02;
  r[0][1] = A01 * B00 + A11 * B01 + A21 * B02;
  r[0][2] = A02 * B00 + A12 * B01 + A22 * B02;
  r[1][0] = A00 * B10 + A10 * B11 + A20 * B12;
  r[1][1] = A01 * B10 + A11 * B11 + A21 * B12;
  r[1][2] = A02 * B10 + A12 * B11 + A22 * B12;
  r[2][0] = A00 * B20 + A10 * B21 + A20 * B22;
  r[2][1] = A01 * B20 + A11 * B21 + A21 * B22;
  r[2][2] = A02 * B20 + A12 * B21 + A22 * B22;
  return r;
}
GLMS_FN mat3 operator*(float s, const mat3& m) { return mat3(m[0] * s, m[1] * s, m[2] * s); }
GLMS_FN mat3 operator*(const mat3& m, float s) { return mat3(m[0] * s, m[1] * s, m[2] * s); }
GLMS_FN mat3 transpose(const mat3& m) {
  return mat3(m[0][0], m[1][0], m[2][0], m[0][1], m[1][1], m[2][1], m[0][2], m[1][2], m[2][2]);
}

}  // namespace glm
