// TEST INFRASTRUCTURE — CPU oracle for landmass_b200 (see oracle/README.md).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, link or call anything under oracle/.
//
// f32 vector helpers restating the glam 0.30 scalar operations the reference
// uses on its hot path (glam source is NOT in /root/reference; semantics per
// SURVEY.md Appendix C).  Every expression is evaluated left to right exactly
// as written; compile with -ffp-contract=off so no FMA is formed.
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

struct V2 {
  float x, y;
};
struct V3 {
  float x, y, z;
};

inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator*(V2 a, float s) { return {a.x * s, a.y * s}; }
inline V2 operator/(V2 a, float s) { return {a.x / s, a.y / s}; }
inline V2 operator-(V2 a) { return {-a.x, -a.y}; }
inline float dot(V2 a, V2 b) { return (a.x * b.x) + (a.y * b.y); }
inline float perp_dot(V2 a, V2 b) { return (a.x * b.y) - (a.y * b.x); }
inline float length_squared(V2 a) { return dot(a, a); }
inline float length(V2 a) { return std::sqrt(dot(a, a)); }
inline float distance(V2 a, V2 b) { return length(a - b); }
inline V2 perp(V2 a) { return {-a.y, a.x}; }
inline V2 normalize(V2 a) { return a * (1.0f / length(a)); }

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline V2 xy(V3 a) { return {a.x, a.y}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
inline float length_squared(V3 a) { return dot(a, a); }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline float distance(V3 a, V3 b) { return length(a - b); }
inline float distance_squared(V3 a, V3 b) { return length_squared(a - b); }
inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline V3 normalize(V3 a) { return a * (1.0f / length(a)); }
inline V3 midpoint(V3 a, V3 b) { return (a + b) * 0.5f; }
inline V3 lerp(V3 a, V3 b, float s) { return a * (1.0f - s) + b * s; }
// Rust f32::clamp
inline float clampf(float v, float lo, float hi) {
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return v;
}
// glam Vec2::normalize_or_zero
inline V2 normalize_or_zero(V2 a) {
  float rcp = 1.0f / length(a);
  if (std::isfinite(rcp) && rcp > 0.0f) return a * rcp;
  return {0.0f, 0.0f};
}

// Quat::from_rotation_z(angle) * v  (glam SSE2 Quat::mul_vec3a):
//   v*(w*w - b.b) + b*(2*(v.b)) + (b x v)*(2*w),  b = (0,0,s), w = c.
inline V3 rotate_z(float s, float c, V3 v) {
  V3 b{0.0f, 0.0f, s};
  float b2 = dot(b, b);
  float k1 = c * c - b2;
  float k2 = dot(v, b) * 2.0f;
  float k3 = c * 2.0f;
  V3 bxv = cross(b, v);
  return (v * k1 + b * k2) + bxv * k3;
}

// Transform (reference util.rs:315-368)
struct Xform {
  V3 translation;
  float rotation;
  V3 apply(V3 p) const {
    float h = rotation * 0.5f;
    return rotate_z(std::sin(h), std::cos(h), p) + translation;
  }
  V3 apply_inverse(V3 p) const {
    float h = (-rotation) * 0.5f;
    return rotate_z(std::sin(h), std::cos(h), p - translation);
  }
};

}  // namespace orc
