// b2_mpr.cuh -- the convex pair path of the narrow phase (one candidate geom pair per lane).
//
// Stands in for mjc_Convex (MuJoCo engine_collision_convex.c), which the reference reaches through mujoco.mj_step for
// capsule-cylinder, cylinder-cylinder and cylinder-box pairs: martial arts' free cylinder dummies
// (humanoid_martial_arts_env/martial_arts_env.py:193,199) and the arm's screws, posts and links
// (robotic_arm_assembly_env/assets/complete_model.xml:51-82,176-179,211-230).  MuJoCo hands these pairs to libccd's
// Minkowski Portal Refinement (ccdMPRPenetration, libccd src/mpr.c): portal discovery from the centre difference,
// refinement until the portal contains the origin ray, expansion until the support plane is within mpr_tolerance (1e-6)
// of the portal or mpr_iterations (50) ran out, then depth / direction from the portal triangle's closest point to the
// origin and the position from barycentric weights.  Each geom is inflated by margin / 2 in its support function
// (mjccd_support) and dist = margin - depth (mjc_MPRIteration); one contact per pair (multiccd off).
//
// This is the one place where the step kernel computes in fp64.  MPR stops on an absolute 1e-6 test of differences of
// support points and takes its direction from a triangle a fraction of a millimetre across: in fp32 the stopping pass and
// the triangle normal are decided by rounding (measured on the host build: normals off by > 1e-3 in a quarter of shallow
// contacts), in fp64 the same inputs give the same portal sequence as MuJoCo's double-precision libccd.  The pair types
// that come here are a handful per env (two dummies, sixteen cylinders on the arm), B200 issues fp64 at half the fp32 rate, and
// the inputs (geom poses from the fp32 kinematics) and outputs (one raw contact) are fp32.
// libccd's two unbounded loops are capped at 100 passes; a NaN from a collapsed portal is "no contact".
#pragma once
#include "b2_math.cuh"   // included by b2_collide.cuh after raw_put

namespace b2 {

#define B2_CCD_EPS 2.2204460492503131e-16
#define B2_MPR_TOL 1e-6
#define B2_MPR_MAXIT 50
#define B2_MPR_LOOP_CAP 100

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 dcross(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
#ifdef B2_HOST_BUILD
__device__ __forceinline__ D3 dunit(D3 a) { return a * (1.0 / sqrt(ddot(a, a))); }
#else
__device__ __forceinline__ D3 dunit(D3 a) { return a * rsqrt(ddot(a, a)); }      // within an ulp of 1 / sqrt: one instruction sequence instead of two
#endif

struct CcdObj { int type; D3 pos; const float* mat; const float* size; double margin; };
struct CcdSup { D3 v; V3 w1; };      // point of the Minkowski difference (fp64: it decides) and its witness on geom 1 relative to geom 1's
                                     // centre (fp32: only the reported position is formed from it; the witness on geom 2 is w1 - v)

__device__ __forceinline__ bool ccd_zero(double x) { return fabs(x) < B2_CCD_EPS; }
__device__ __forceinline__ bool ccd_eq(double a_, double b_) {
  const double ab = fabs(a_ - b_);
  if (ab < B2_CCD_EPS) return true;
  const double a = fabs(a_), b = fabs(b_);
  return b > a ? ab < B2_CCD_EPS * b : ab < B2_CCD_EPS * a;
}
__device__ __forceinline__ bool ccd_is_origin(D3 v) { return ccd_eq(v.x, 0.0) && ccd_eq(v.y, 0.0) && ccd_eq(v.z, 0.0); }
__device__ __forceinline__ double dsign0(double x) { return x < 0.0 ? -1.0 : (x > 0.0 ? 1.0 : 0.0); }

// mjccd_support: farthest point of the margin-inflated geom along the unit direction, world frame
__device__ __forceinline__ D3 ccd_support1(const CcdObj& o, D3 dir) {
  const float* m = o.mat;
  const D3 ld = d3(m[0] * dir.x + m[3] * dir.y + m[6] * dir.z, m[1] * dir.x + m[4] * dir.y + m[7] * dir.z, m[2] * dir.x + m[5] * dir.y + m[8] * dir.z);
  D3 r = d3(0.0, 0.0, 0.0);
  const double s0 = o.size[0], s1 = o.size[1];
  if (o.type == 2) r = ld * s0;                                                              // sphere
  else if (o.type == 3) r = d3(ld.x * s0, ld.y * s0, ld.z * s0 + dsign0(ld.z) * s1);         // capsule
  else if (o.type == 5) {                                                                    // cylinder
    const double t2 = ld.x * ld.x + ld.y * ld.y;
#ifdef B2_HOST_BUILD
    if (t2 > 1e-30) { const double t = sqrt(t2); r.x = ld.x / t * s0; r.y = ld.y / t * s0; }
#else
    if (t2 > 1e-30) { const double k = rsqrt(t2) * s0; r.x = ld.x * k; r.y = ld.y * k; }
#endif
    r.z = dsign0(ld.z) * s1;
  } else r = d3(dsign0(ld.x) * s0, dsign0(ld.y) * s1, dsign0(ld.z) * (double)o.size[2]);    // box
  return d3(m[0] * r.x + m[1] * r.y + m[2] * r.z + (o.pos.x + dir.x * o.margin), m[3] * r.x + m[4] * r.y + m[5] * r.z + (o.pos.y + dir.y * o.margin),
            m[6] * r.x + m[7] * r.y + m[8] * r.z + (o.pos.z + dir.z * o.margin));
}
__device__ __noinline__ void ccd_support(const CcdObj& o1, const CcdObj& o2, D3 dir, CcdSup& s) {   // __ccdSupport
  const D3 a = ccd_support1(o1, dir); s.v = a - ccd_support1(o2, dir * -1.0);
  s.w1 = v3((float)(a.x - o1.pos.x), (float)(a.y - o1.pos.y), (float)(a.z - o1.pos.z));
}
__device__ __forceinline__ D3 portal_dir(const CcdSup* p) { return dunit(dcross(p[2].v - p[1].v, p[3].v - p[1].v)); }
__device__ __forceinline__ bool portal_reach_tolerance(const CcdSup* p, const CcdSup& v4, D3 dir) {
  const double dv4 = ddot(v4.v, dir);
  double d1 = dv4 - ddot(p[1].v, dir); const double d2 = dv4 - ddot(p[2].v, dir), d3_ = dv4 - ddot(p[3].v, dir);
  if (d2 < d1) d1 = d2;
  if (d3_ < d1) d1 = d3_;
  return ccd_eq(d1, B2_MPR_TOL) || d1 < B2_MPR_TOL;
}
__device__ __forceinline__ void expand_portal(CcdSup* p, const CcdSup& v4) {
  const D3 v4v0 = dcross(v4.v, p[0].v);
  if (ddot(p[1].v, v4v0) > 0.0) { if (ddot(p[2].v, v4v0) > 0.0) p[1] = v4; else p[3] = v4; }
  else { if (ddot(p[3].v, v4v0) > 0.0) p[2] = v4; else p[1] = v4; }
}
// squared distance from the origin to the segment x0-b, closest point in wit (__ccdVec3PointSegmentDist2)
__device__ __forceinline__ double point_seg_dist2(D3 x0, D3 b, D3& wit) {
  const D3 d = b - x0; const double t = -ddot(x0, d) / ddot(d, d);
  if (t < 0.0 || ccd_zero(t)) wit = x0;
  else if (t > 1.0 || ccd_eq(t, 1.0)) wit = b;
  else wit = d * t + x0;
  return ddot(wit, wit);
}
// squared distance from the origin to the triangle x0 B C (ccdVec3PointTriDist2)
__device__ __forceinline__ double point_tri_dist2(D3 x0, D3 B, D3 C, D3& wit) {
  const D3 d1 = B - x0, d2 = C - x0;
  const double v = ddot(d1, d1), w = ddot(d2, d2), p = ddot(x0, d1), q = ddot(x0, d2), r = ddot(d1, d2);
  const double dd = w * v - r * r; double s = -1.0, t = -1.0;
  if (!ccd_zero(dd)) { s = (q * r - w * p) / dd; t = (-s * r - q) / w; }
  if ((ccd_zero(s) || s > 0.0) && (ccd_eq(s, 1.0) || s < 1.0) && (ccd_zero(t) || t > 0.0) && (ccd_eq(t, 1.0) || t < 1.0) &&
      (ccd_eq(t + s, 1.0) || t + s < 1.0)) {
    wit = x0 + d1 * s + d2 * t;
    return ddot(wit, wit);
  }
  D3 w2; double dist = point_seg_dist2(x0, B, wit), dist2 = point_seg_dist2(x0, C, w2);
  if (dist2 < dist) { dist = dist2; wit = w2; }
  dist2 = point_seg_dist2(B, C, w2);
  if (dist2 < dist) { dist = dist2; wit = w2; }
  return dist;
}
// discoverPortal: -1 no intersection, 0 portal found, 1 origin on v1, 2 origin on the v0-v1 segment
__device__ __forceinline__ int discover_portal(const CcdObj& o1, const CcdObj& o2, CcdSup* p) {
  p[0].w1 = v3(0.f, 0.f, 0.f); p[0].v = o1.pos - o2.pos;
  if (ccd_is_origin(p[0].v)) p[0].v.x += B2_CCD_EPS * 10.0;
  D3 dir = dunit(p[0].v * -1.0);
  ccd_support(o1, o2, dir, p[1]);
  double dt = ddot(p[1].v, dir);
  if (ccd_zero(dt) || dt < 0.0) return -1;
  dir = dcross(p[0].v, p[1].v);
  if (ccd_zero(ddot(dir, dir))) return ccd_is_origin(p[1].v) ? 1 : 2;
  dir = dunit(dir);
  ccd_support(o1, o2, dir, p[2]);
  dt = ddot(p[2].v, dir);
  if (ccd_zero(dt) || dt < 0.0) return -1;
  dir = dunit(dcross(p[1].v - p[0].v, p[2].v - p[0].v));
  if (ddot(dir, p[0].v) > 0.0) { const CcdSup t = p[1]; p[1] = p[2]; p[2] = t; dir = dir * -1.0; }
  for (int pass = 0; pass < B2_MPR_LOOP_CAP; pass++) {
    ccd_support(o1, o2, dir, p[3]);
    dt = ddot(p[3].v, dir);
    if (ccd_zero(dt) || dt < 0.0) return -1;
    bool cont = false;
    dt = ddot(dcross(p[1].v, p[3].v), p[0].v);
    if (dt < 0.0 && !ccd_zero(dt)) { p[2] = p[3]; cont = true; }
    if (!cont) {
      dt = ddot(dcross(p[3].v, p[2].v), p[0].v);
      if (dt < 0.0 && !ccd_zero(dt)) { p[1] = p[3]; cont = true; }
    }
    if (!cont) return 0;
    dir = dunit(dcross(p[1].v - p[0].v, p[2].v - p[0].v));
  }
  return -1;
}

// fp32 support point of the un-inflated geom along a unit direction, relative to the geom's centre (pre-cull below)
__device__ __forceinline__ V3 support_f32(int type, const float* mat, const float* size, V3 dir) {
  const V3 ld = mulmatT(mat, dir); V3 r = v3(0.f, 0.f, 0.f);
  if (type == 2) r = ld * size[0];
  else if (type == 3) r = v3(ld.x * size[0], ld.y * size[0], fmaf(ld.z, size[0], (ld.z < 0.f ? -size[1] : size[1])));
  else if (type == 5) {
    const float t = sqrtf(ld.x * ld.x + ld.y * ld.y);
    if (t > 1e-15f) { r.x = ld.x / t * size[0]; r.y = ld.y / t * size[0]; }
    r.z = ld.z < 0.f ? -size[1] : size[1];
  } else r = v3(ld.x < 0.f ? -size[0] : size[0], ld.y < 0.f ? -size[1] : size[1], ld.z < 0.f ? -size[2] : size[2]);
  return mulmat(mat, r);
}
// Pure pruning in front of the fp64 path.  Two convex sets whose supports leave a gap along ANY direction do not intersect, and
// MPR reports "no intersection" for them (from its first support point when the direction is the centre-to-centre one, after a
// few portal steps otherwise).  A handful of cheap fp32 candidates are tried: the centre direction, the axes of the cylinders
// (a wide flat disc -- the dancer's floor and stage -- is inside the bounding sphere of everything above it, and only its own
// axis separates it from a geom hovering over its rim) and the face normals of a box, each signed towards the other geom.  The
// gap must exceed the margin by 1e-5 (fp32 rounding of the supports), so a pair MPR could call touching is never dropped.
__device__ __forceinline__ bool convex_gap_along(int t1, const float* m1, const float* s1, int t2, const float* m2, const float* s2, V3 c, V3 d, float margin) {
  if (dot(c, d) < 0.f) d = d * -1.f;      // from geom 1 towards geom 2
  return dot(support_f32(t1, m1, s1, d), d) - dot(support_f32(t2, m2, s2, d * -1.f), d) + margin + 1e-5f < dot(c, d);
}
__device__ __forceinline__ bool convex_far_apart(int t1, V3 pos1, const float* m1, const float* s1, int t2, V3 pos2, const float* m2, const float* s2, float margin) {
  const V3 c = pos2 - pos1; const float len = norm(c);
  if (len > 1e-6f && convex_gap_along(t1, m1, s1, t2, m2, s2, c, c * (1.0f / len), margin)) return true;
  if (t1 == 5 && convex_gap_along(t1, m1, s1, t2, m2, s2, c, matcol(m1, 2), margin)) return true;
  if (t2 == 5 && convex_gap_along(t1, m1, s1, t2, m2, s2, c, matcol(m2, 2), margin)) return true;
  if (t2 == 6) {
#pragma unroll 1
    for (int k = 0; k < 3; k++) if (convex_gap_along(t1, m1, s1, t2, m2, s2, c, matcol(m2, k), margin)) return true;
  }
  return false;
}

// ccdMPRPenetration + mjc_MPRIteration: writes at most one raw contact, normal from geom 1 to geom 2
__device__ __noinline__ int c_convex_mpr(float* dst, int t1, V3 pos1, const float* m1, const float* s1, int t2, V3 pos2, const float* m2,
                                         const float* s2, float margin) {
#ifdef B2_NO_MPR      /* A/B only: what the convex path costs a task */
  return 0;
#endif
  CcdObj o1, o2;
  o1.type = t1; o1.pos = d3(pos1.x, pos1.y, pos1.z); o1.mat = m1; o1.size = s1; o1.margin = 0.5 * (double)margin;
  o2.type = t2; o2.pos = d3(pos2.x, pos2.y, pos2.z); o2.mat = m2; o2.size = s2; o2.margin = 0.5 * (double)margin;
  CcdSup p[4]; double depth; D3 n, pos;
  const int res = discover_portal(o1, o2, p);
  if (res < 0 || res == 1) return 0;     // touching at v1 has no direction: mjc_MPRIteration discards it
  if (res == 2) {                        // findPenetrSegment
    pos = d3(p[1].w1.x, p[1].w1.y, p[1].w1.z) + o1.pos - p[1].v * 0.5;      // (v1 + v2) / 2 with v2 = v1 - v
    depth = sqrt(ddot(p[1].v, p[1].v)); n = dunit(p[1].v);
  } else {
    bool inside = false; D3 dir; CcdSup v4;
    for (int pass = 0; pass < B2_MPR_LOOP_CAP; pass++) {         // refinePortal
      dir = portal_dir(p);
      double dt = ddot(dir, p[1].v);
      if (ccd_zero(dt) || dt > 0.0) { inside = true; break; }
      ccd_support(o1, o2, dir, v4);
      dt = ddot(v4.v, dir);
      if (!(ccd_zero(dt) || dt > 0.0) || portal_reach_tolerance(p, v4, dir)) return 0;
      expand_portal(p, v4);
    }
    if (!inside) return 0;
    for (int it = 0;; it++) {                                     // findPenetr
      dir = portal_dir(p);
      ccd_support(o1, o2, dir, v4);
      if (portal_reach_tolerance(p, v4, dir) || it > B2_MPR_MAXIT) break;
      expand_portal(p, v4);
    }
    depth = sqrt(point_tri_dist2(p[1].v, p[2].v, p[3].v, n));
    if (ccd_zero(depth)) return 0;
    n = dunit(n);
    // findPos: barycentric weights of the origin in the tetrahedron v0 v1 v2 v3
    dir = portal_dir(p);
    double b0 = ddot(dcross(p[1].v, p[2].v), p[3].v), b1 = ddot(dcross(p[3].v, p[2].v), p[0].v);
    double b2 = ddot(dcross(p[0].v, p[1].v), p[3].v), b3 = ddot(dcross(p[2].v, p[1].v), p[0].v);
    double sum = b0 + b1 + b2 + b3;
    if (ccd_zero(sum) || sum < 0.0) {
      b0 = 0.0; b1 = ddot(dcross(p[2].v, p[3].v), dir); b2 = ddot(dcross(p[3].v, p[1].v), dir); b3 = ddot(dcross(p[1].v, p[2].v), dir);
      sum = b1 + b2 + b3;
    }
    const double inv = 1.0 / sum;
    const double w0 = b0 * inv, w1 = b1 * inv, w2 = b2 * inv, w3 = b3 * inv;      // sum to one: the shift by geom 1's centre passes through
    const D3 a1 = d3(p[0].w1.x * w0 + p[1].w1.x * w1 + p[2].w1.x * w2 + p[3].w1.x * w3, p[0].w1.y * w0 + p[1].w1.y * w1 + p[2].w1.y * w2 + p[3].w1.y * w3,
                     p[0].w1.z * w0 + p[1].w1.z * w1 + p[2].w1.z * w2 + p[3].w1.z * w3);
    const D3 av = p[0].v * w0 + p[1].v * w1 + p[2].v * w2 + p[3].v * w3;
    pos = a1 + o1.pos - av * 0.5;
  }
  if (ccd_is_origin(n)) return 0;
  if (!(depth == depth) || !(n.x == n.x) || !(pos.x == pos.x) || !(pos.y == pos.y) || !(pos.z == pos.z)) return 0;
  const double nl = 1.0 / sqrt(ddot(n, n));                      // mju_normalize3 of the frame normal
  raw_put(dst, (float)((double)margin - depth), v3((float)pos.x, (float)pos.y, (float)pos.z), v3((float)(n.x * nl), (float)(n.y * nl), (float)(n.z * nl)),
          v3(0.f, 0.f, 0.f));
  return 1;
}

}  // namespace b2
