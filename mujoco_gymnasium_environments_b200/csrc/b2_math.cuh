// b2_math.cuh -- small fp32 vector / quaternion / spatial-algebra helpers for the step kernel.
// Spatial vectors are [angular(3); linear(3)] expressed at the tree's com origin.
#pragma once
#ifndef B2_HOST_BUILD
#include <cuda_runtime.h>
#endif

namespace b2 {

struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float norm(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 normalized(V3 a, float* len = nullptr) {
  float n = norm(a);
  if (len) *len = n;
  if (n < 1e-15f) return v3(1.f, 0.f, 0.f);
  float s = 1.0f / n;
  return a * s;
}
// row-major 3x3 times vector
__device__ __forceinline__ V3 mulmat(const float* m, V3 v) {
  return v3(fmaf(m[0], v.x, fmaf(m[1], v.y, m[2] * v.z)), fmaf(m[3], v.x, fmaf(m[4], v.y, m[5] * v.z)),
            fmaf(m[6], v.x, fmaf(m[7], v.y, m[8] * v.z)));
}
// transpose(m) * v
__device__ __forceinline__ V3 mulmatT(const float* m, V3 v) {
  return v3(fmaf(m[0], v.x, fmaf(m[3], v.y, m[6] * v.z)), fmaf(m[1], v.x, fmaf(m[4], v.y, m[7] * v.z)),
            fmaf(m[2], v.x, fmaf(m[5], v.y, m[8] * v.z)));
}
__device__ __forceinline__ V3 matcol(const float* m, int c) { return v3(m[c], m[3 + c], m[6 + c]); }

__device__ __forceinline__ Q4 ldq(const float* p) { Q4 q; q.w = p[0]; q.x = p[1]; q.y = p[2]; q.z = p[3]; return q; }
__device__ __forceinline__ void stq(float* p, Q4 q) { p[0] = q.w; p[1] = q.x; p[2] = q.y; p[3] = q.z; }
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
__device__ __forceinline__ Q4 qnormalize(Q4 q) {
  float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  if (n < 1e-15f) { q.w = 1.f; q.x = q.y = q.z = 0.f; return q; }
  float s = 1.0f / n;
  q.w *= s; q.x *= s; q.y *= s; q.z *= s;
  return q;
}
__device__ __forceinline__ void quat2mat(float* m, Q4 q) {
  float q00 = q.w * q.w, q11 = q.x * q.x, q22 = q.y * q.y, q33 = q.z * q.z;
  float q01 = q.w * q.x, q02 = q.w * q.y, q03 = q.w * q.z, q12 = q.x * q.y, q13 = q.x * q.z, q23 = q.y * q.z;
  m[0] = q00 + q11 - q22 - q33; m[1] = 2.f * (q12 - q03); m[2] = 2.f * (q13 + q02);
  m[3] = 2.f * (q12 + q03); m[4] = q00 - q11 + q22 - q33; m[5] = 2.f * (q23 - q01);
  m[6] = 2.f * (q13 - q02); m[7] = 2.f * (q23 + q01); m[8] = q00 - q11 - q22 + q33;
}
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
  // v + 2 w (u x v) + 2 u x (u x v)
  V3 u = v3(q.x, q.y, q.z);
  V3 t = cross(u, v) * 2.f;
  return v + t * q.w + cross(u, t);
}
__device__ __forceinline__ Q4 axisangle(V3 axis, float angle) {
  float s, c;
  sincosf(0.5f * angle, &s, &c);
  Q4 q; q.w = c; q.x = axis.x * s; q.y = axis.y * s; q.z = axis.z * s;
  return q;
}

// 6-D spatial helpers on raw float[6] = [ang; lin]
struct S6 { V3 a, l; };
__device__ __forceinline__ S6 ld6(const float* p) { S6 s; s.a = ld3(p); s.l = ld3(p + 3); return s; }
__device__ __forceinline__ void st6(float* p, S6 s) { st3(p, s.a); st3(p + 3, s.l); }
__device__ __forceinline__ S6 operator+(S6 x, S6 y) { S6 r; r.a = x.a + y.a; r.l = x.l + y.l; return r; }
__device__ __forceinline__ S6 operator*(S6 x, float s) { S6 r; r.a = x.a * s; r.l = x.l * s; return r; }
__device__ __forceinline__ float dot6(S6 x, S6 y) { return dot(x.a, y.a) + dot(x.l, y.l); }
// motion cross product: vel x v
__device__ __forceinline__ S6 cross_motion(S6 vel, S6 v) {
  S6 r; r.a = cross(vel.a, v.a); r.l = cross(vel.a, v.l) + cross(vel.l, v.a); return r;
}
// force cross product: vel x* f
__device__ __forceinline__ S6 cross_force(S6 vel, S6 f) {
  S6 r; r.a = cross(vel.a, f.a) + cross(vel.l, f.l); r.l = cross(vel.a, f.l); return r;
}
// 10-number spatial inertia [Ixx Iyy Izz Ixy Ixz Iyz | m*c(3) | m] times motion vector
__device__ __forceinline__ S6 mul_inert(const float* i, S6 v) {
  S6 r;
  r.a.x = i[0] * v.a.x + i[3] * v.a.y + i[4] * v.a.z - i[8] * v.l.y + i[7] * v.l.z;
  r.a.y = i[3] * v.a.x + i[1] * v.a.y + i[5] * v.a.z + i[8] * v.l.x - i[6] * v.l.z;
  r.a.z = i[4] * v.a.x + i[5] * v.a.y + i[2] * v.a.z - i[7] * v.l.x + i[6] * v.l.y;
  r.l.x = i[8] * v.a.y - i[7] * v.a.z + i[9] * v.l.x;
  r.l.y = i[6] * v.a.z - i[8] * v.a.x + i[9] * v.l.y;
  r.l.z = i[7] * v.a.x - i[6] * v.a.y + i[9] * v.l.z;
  return r;
}

#ifndef B2_HOST_BUILD
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// several sums in one butterfly: the shuffles of the different values are independent and overlap
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { float x = __shfl_xor_sync(0xffffffffu, a, o), y = __shfl_xor_sync(0xffffffffu, b, o); a += x; b += y; }
}
__device__ __forceinline__ void warp_sum3(float& a, float& b, float& c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float x = __shfl_xor_sync(0xffffffffu, a, o), y = __shfl_xor_sync(0xffffffffu, b, o), z = __shfl_xor_sync(0xffffffffu, c, o);
    a += x; b += y; c += z;
  }
}

#endif
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

}  // namespace b2
