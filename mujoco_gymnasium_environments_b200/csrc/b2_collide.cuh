// b2_collide.cuh -- narrow-phase primitives of the step kernel (fp32, one candidate geom pair per lane).
//
// Stands in for the geom-pair functions behind mujoco.mj_step's collision stage (mjc_PlaneSphere ... mjc_BoxBox and
// the convex path MuJoCo uses for cylinders).  Every function writes raw contacts of 10 floats
// [dist, pos(3), normal(3), tangent hint(3)] into the pair's private slot of the shared-memory arena and returns the
// count; the normal points from geom 1 to geom 2 (geom 1 has the lower MuJoCo geom-type id).  Pair types:
//   plane-{sphere,capsule,box}     closed form (MuJoCo's primitives)
//   sphere-{sphere,capsule,box,cylinder}, capsule-capsule   closed form
//   capsule-box      closest segment point by bisection on the (monotone) derivative of the distance, then sphere-box;
//                    a second sphere-box test at the far end of the stretch lying over the same face
//   box-box          15-axis separating-axis test; face case clips the incident face against the reference face
//                    (the pair's own raw slot doubles as the polygon scratch), edge case = closest points of two edges
//   plane-cylinder   mjc_PlaneCylinder restated (up to four contacts)
//   capsule-cylinder, cylinder-box, cylinder-cylinder   mjc_Convex: libccd's Minkowski Portal Refinement restated
//                    (b2_mpr.cuh), one contact per pair
#pragma once
#include "b2_math.cuh"

namespace b2 {

enum { GT_PLANE = 0, GT_SPHERE = 2, GT_CAPSULE = 3, GT_CYLINDER = 5, GT_BOX = 6 };
#define B2_RAW 10

__device__ __forceinline__ void raw_put(float* dst, float dist, V3 pos, V3 n, V3 t) {
  dst[0] = dist; st3(dst + 1, pos); st3(dst + 4, n); st3(dst + 7, t);
}
__device__ __forceinline__ float sel3(V3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

// ---- plane vs X
__device__ __forceinline__ int c_plane_sphere(float* dst, V3 p1, V3 n, V3 p2, float r, float margin, V3 hint) {
  float cd = dot(p2 - p1, n);
  if (cd > margin + r) return 0;
  float dist = cd - r;
  raw_put(dst, dist, p2 + n * (-dist * 0.5f - r), n, hint);
  return 1;
}
__device__ __forceinline__ int c_plane_capsule(float* dst, V3 p1, V3 n, V3 p2, const float* m2, const float* s2, float margin) {
  V3 ax = matcol(m2, 2);
  int cnt = c_plane_sphere(dst, p1, n, p2 + ax * s2[1], s2[0], margin, ax);
  cnt += c_plane_sphere(dst + B2_RAW * cnt, p1, n, p2 - ax * s2[1], s2[0], margin, ax);
  return cnt;
}
__device__ __forceinline__ int c_plane_box(float* dst, V3 p1, V3 n, V3 p2, const float* m2, const float* s2, float margin, int maxn) {
  float dist = dot(p2 - p1, n); int cnt = 0;
  for (int i = 0; i < 8; i++) {
    V3 v = v3((i & 1) ? s2[0] : -s2[0], (i & 2) ? s2[1] : -s2[1], (i & 4) ? s2[2] : -s2[2]);
    V3 cn = mulmat(m2, v);
    float ld = dot(n, cn);
    if (dist + ld > margin || ld > 0.f) continue;
    float d = dist + ld;
    raw_put(dst + B2_RAW * cnt, d, cn + p2 - n * (d * 0.5f), n, v3(0, 0, 0));
    if (++cnt >= 4 || cnt >= maxn) return cnt;
  }
  return cnt;
}

// mjc_PlaneCylinder restated: the disc-edge point nearest the plane on each cap, plus two side points on the near cap
__device__ __noinline__ int c_plane_cylinder(float* dst, V3 p1, V3 n, V3 p2, const float* m2, const float* s2, float margin, int maxn) {
  V3 ax = matcol(m2, 2);
  float dist0 = dot(p2 - p1, n), prjaxis = dot(n, ax);
  if (prjaxis > 0.f) { ax = ax * -1.f; prjaxis = -prjaxis; }
  V3 vec = ax * prjaxis - n; float len = norm(vec);
  if (len < 1e-12f) vec = matcol(m2, 0) * s2[0]; else vec = vec * (s2[0] / len);
  float prjvec = dot(vec, n);
  ax = ax * s2[1]; prjaxis *= s2[1];
  int cnt = 0;
  if (dist0 + prjaxis + prjvec > margin) return 0;
  { float d = dist0 + prjaxis + prjvec; raw_put(dst, d, p2 + vec + ax - n * (d * 0.5f), n, v3(0, 0, 0)); cnt++; }
  if (cnt < maxn && dist0 - prjaxis + prjvec <= margin) {
    float d = dist0 - prjaxis + prjvec; raw_put(dst + B2_RAW * cnt, d, p2 + vec - ax - n * (d * 0.5f), n, v3(0, 0, 0)); cnt++;
  }
  float prjvec1 = -prjvec * 0.5f;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    V3 vec1 = cross(vec, ax); vec1 = vec1 * (s2[0] * 0.8660254f / norm(vec1));
    float d = dist0 + prjaxis + prjvec1;
    for (int sgi = 0; sgi < 2 && cnt < maxn; sgi++) {
      raw_put(dst + B2_RAW * cnt, d, p2 + vec1 * (sgi ? -1.f : 1.f) + ax - vec * 0.5f - n * (d * 0.5f), n, v3(0, 0, 0)); cnt++;
    }
  }
  return cnt;
}

// ---- sphere vs X
__device__ __forceinline__ int c_sphere_sphere(float* dst, V3 p1, float r1, V3 p2, float r2, float margin) {
  V3 dif = p2 - p1; float cd2 = dot(dif, dif), rs = margin + r1 + r2;
  if (cd2 > rs * rs) return 0;
  float cd = sqrtf(cd2), dist = cd - r1 - r2;
  V3 n = cd < 1e-15f ? v3(1, 0, 0) : dif * (1.0f / cd);
  raw_put(dst, dist, p1 + n * (r1 + 0.5f * dist), n, v3(0, 0, 0));
  return 1;
}
__device__ __forceinline__ int c_sphere_capsule(float* dst, V3 p1, float r1, V3 p2, const float* m2, const float* s2, float margin) {
  V3 ax = matcol(m2, 2);
  float x = clampf(dot(ax, p1 - p2), -s2[1], s2[1]);
  return c_sphere_sphere(dst, p1, r1, p2 + ax * x, s2[0], margin);
}
__device__ __noinline__ int c_sphere_box(float* dst, V3 p1, float r1, V3 p2, const float* m2, const float* s2, float margin) {
  V3 c = mulmatT(m2, p1 - p2);
  V3 cl = v3(clampf(c.x, -s2[0], s2[0]), clampf(c.y, -s2[1], s2[1]), clampf(c.z, -s2[2], s2[2]));
  V3 dv = c - cl; float dist = norm(dv);
  if (dist - r1 > margin) return 0;
  V3 nl, posl; float d;
  if (dist <= 1e-15f) {   // centre inside the box: leave through the nearest face (faces scanned -x,+x,-y,+y,-z,+z)
    float closest = 2.f * (s2[0] + s2[1] + s2[2]); int k = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      float face = ((i & 1) ? 1.f : -1.f) * s2[i >> 1], dd = fabsf(sel3(c, i >> 1) - face);
      if (closest > dd) { closest = dd; k = i; }
    }
    float sg = (k & 1) ? -1.f : 1.f;
    nl = v3((k >> 1) == 0 ? sg : 0.f, (k >> 1) == 1 ? sg : 0.f, (k >> 1) == 2 ? sg : 0.f);
    posl = c + nl * ((r1 - closest) * 0.5f);
    d = -closest - r1;
  } else {
    float inv = 1.0f / dist;
    V3 deepest = c - dv * (inv * r1);
    nl = (cl - c) * inv; posl = (cl + deepest) * 0.5f; d = dist - r1;
  }
  raw_put(dst, d, mulmat(m2, posl) + p2, mulmat(m2, nl), v3(0, 0, 0));
  return 1;
}
// signed distance of p (cylinder frame) to the solid cylinder; closest surface point and outward direction there
__device__ __forceinline__ float point_cylinder(V3 p, float r, float h, V3& cp, V3& n) {
  float rho = sqrtf(p.x * p.x + p.y * p.y), az = fabsf(p.z), sz = p.z < -1e-6f ? -1.f : 1.f;   // mid-plane ties leave upwards
  float ux = 1.f, uy = 0.f;
  if (rho > 1e-15f) { ux = p.x / rho; uy = p.y / rho; }
  if (az <= h && rho <= r) {
    if (h - az <= r - rho) { cp = v3(p.x, p.y, sz * h); n = v3(0, 0, sz); return -(h - az); }
    cp = v3(ux * r, uy * r, p.z); n = v3(ux, uy, 0); return -(r - rho);
  }
  if (rho <= r) { cp = v3(p.x, p.y, sz * h); n = v3(0, 0, sz); return az - h; }
  if (az <= h) { cp = v3(ux * r, uy * r, p.z); n = v3(ux, uy, 0); return rho - r; }
  cp = v3(ux * r, uy * r, sz * h);
  V3 v = p - cp; float dd = norm(v);
  n = v * (1.0f / dd);
  return dd;
}
__device__ __forceinline__ int c_sphere_cylinder(float* dst, V3 p1, float r1, V3 p2, const float* m2, const float* s2, float margin) {
  V3 cp, n; float dist = point_cylinder(mulmatT(m2, p1 - p2), s2[0], s2[1], cp, n) - r1;
  if (dist > margin) return 0;
  raw_put(dst, dist, mulmat(m2, cp + n * (dist * 0.5f)) + p2, mulmat(m2, n * -1.f), v3(0, 0, 0));
  return 1;
}
// ---- capsule vs capsule (mjc_CapsuleCapsule restated: closest points of two segments, parallel case by end tests)
__device__ __noinline__ int c_capsule_capsule(float* dst, V3 p1, const float* m1, const float* s1, V3 p2, const float* m2,
                                              const float* s2, float margin) {
  V3 a1 = matcol(m1, 2), a2 = matcol(m2, 2), dif = p1 - p2;
  float ma = dot(a1, a1), mb = -dot(a1, a2), mc = dot(a2, a2), u = -dot(a1, dif), v = dot(a2, dif), det = ma * mc - mb * mb;
  float l1 = s1[1], l2 = s2[1];
  if (fabsf(det) >= 1e-15f) {
    float x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > l1) { x1 = l1; x2 = (v - mb * l1) / mc; }
    else if (x1 < -l1) { x1 = -l1; x2 = (v + mb * l1) / mc; }
    if (x2 > l2) { x2 = l2; x1 = clampf((u - mb * l2) / ma, -l1, l1); }
    else if (x2 < -l2) { x2 = -l2; x1 = clampf((u + mb * l2) / ma, -l1, l1); }
    return c_sphere_sphere(dst, p1 + a1 * x1, s1[0], p2 + a2 * x2, s2[0], margin);
  }
  int n = 0;
  for (int s = 0; s < 2 && n < 2; s++) {
    V3 q1 = p1 + a1 * (s ? -l1 : l1);
    float x2 = clampf(dot(a2, q1 - p2), -l2, l2);
    n += c_sphere_sphere(dst + B2_RAW * n, q1, s1[0], p2 + a2 * x2, s2[0], margin);
  }
  return n;
}

// ---- capsule vs box
// derivative (up to a factor 2) of the squared distance between the point c + a t (box frame) and the box
__device__ __forceinline__ float seg_box_slope(V3 c, V3 a, const float* s, float t) {
  float g = 0.f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float ak = sel3(a, k), p = fmaf(ak, t, sel3(c, k));
    if (p > s[k]) g = fmaf(ak, p - s[k], g); else if (p < -s[k]) g = fmaf(ak, p + s[k], g);
  }
  return g;
}
__device__ __noinline__ int c_capsule_box(float* dst, V3 p1, const float* m1, const float* s1, V3 p2, const float* m2,
                                          const float* s2, float margin) {
  V3 ax = matcol(m1, 2);
  V3 cl = mulmatT(m2, p1 - p2), al = mulmatT(m2, ax);
  float l = s1[1], r = s1[0];
  {   // axis passes through the box: one contact at the chord point nearest the capsule centre
    float te = -l, tx = l; bool hit = true;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float ak = sel3(al, k), ck = sel3(cl, k);
      if (fabsf(ak) <= 1e-9f) { if (fabsf(ck) > s2[k]) hit = false; continue; }
      float u = (-s2[k] - ck) / ak, v = (s2[k] - ck) / ak;
      te = fmaxf(te, fminf(u, v)); tx = fminf(tx, fmaxf(u, v));
    }
    if (hit && te <= tx) return c_sphere_box(dst, p1 + ax * fminf(fmaxf(0.f, te), tx), r, p2, m2, s2, margin);
  }
  // smallest t in [-l, l] where the slope is >= 0 (the distance is convex in t)
  float tbest;
  if (seg_box_slope(cl, al, s2, -l) >= 0.f) tbest = -l;
  else if (seg_box_slope(cl, al, s2, l) < 0.f) tbest = l;
  else {
    float lo = -l, hi = l;
    for (int it = 0; it < 26; it++) { float mid = 0.5f * (lo + hi); if (seg_box_slope(cl, al, s2, mid) >= 0.f) hi = mid; else lo = mid; }
    // polish: stationary point of the quadratic piece just left of hi, kept inside the bracket
    float num = 0.f, den = 0.f, tm = lo;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float ak = sel3(al, k), ck = sel3(cl, k), p = fmaf(ak, tm, ck);
      if (p > s2[k]) { num = fmaf(ak, ck - s2[k], num); den = fmaf(ak, ak, den); }
      else if (p < -s2[k]) { num = fmaf(ak, ck + s2[k], num); den = fmaf(ak, ak, den); }
    }
    tbest = hi;
    if (den > 1e-12f) { float ts = -num / den; if (ts >= lo && ts <= hi) tbest = ts; }
  }
  int n1 = c_sphere_box(dst, p1 + ax * tbest, r, p2, m2, s2, margin);
  if (!n1) return 0;
  int nout = 0, kf = -1;
#pragma unroll
  for (int k = 0; k < 3; k++) if (fabsf(fmaf(sel3(al, k), tbest, sel3(cl, k))) > s2[k] + 1e-6f + 1e-5f * s2[k]) { nout++; kf = k; }
  if (nout != 1) return 1;
  float ta = -l, tb = l; bool ok = true;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (k == kf) continue;
    float ak = sel3(al, k), ck = sel3(cl, k);
    if (fabsf(ak) <= 1e-9f) { if (fabsf(ck) > s2[k]) ok = false; continue; }
    float u = (-s2[k] - ck) / ak, v = (s2[k] - ck) / ak;
    ta = fmaxf(ta, fminf(u, v)); tb = fminf(tb, fmaxf(u, v));
  }
  if (!ok || tb < ta) return 1;
  float t2 = (tbest - ta > tb - tbest) ? ta : tb;
  if (!(fabsf(t2 - tbest) > 1e-3f * l)) return 1;
  float* d2 = dst + B2_RAW;
  int n2 = c_sphere_box(d2, p1 + ax * t2, r, p2, m2, s2, margin);
  if (n2 && t2 < tbest) {   // contacts are emitted in ascending t
#pragma unroll
    for (int k = 0; k < B2_RAW; k++) { float t = dst[k]; dst[k] = d2[k]; d2[k] = t; }
  }
  return 1 + n2;
}

// ---- box vs box.  dst must have room for 8 raw contacts (80 floats); floats [32,80) double as polygon scratch.
__device__ __noinline__ int c_box_box(float* dst, V3 p1, const float* m1, const float* s1, V3 p2, const float* m2,
                                      const float* s2, float margin) {
  V3 A[3] = {matcol(m1, 0), matcol(m1, 1), matcol(m1, 2)}, B[3] = {matcol(m2, 0), matcol(m2, 1), matcol(m2, 2)};
  V3 p = p2 - p1;
  float Q[3][3], pa[3], pb[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    pa[i] = dot(p, A[i]); pb[i] = dot(p, B[i]);
#pragma unroll
    for (int j = 0; j < 3; j++) Q[i][j] = fabsf(dot(A[i], B[j]));
  }
  float best = -3.0e38f, bsign = 1.f; int code = -1;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float s = fabsf(pa[i]) - (s1[i] + s2[0] * Q[i][0] + s2[1] * Q[i][1] + s2[2] * Q[i][2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = i; bsign = pa[i] < 0.f ? -1.f : 1.f; }
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float s = fabsf(pb[j]) - (s2[j] + s1[0] * Q[0][j] + s1[1] * Q[1][j] + s1[2] * Q[2][j]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = 3 + j; bsign = pb[j] < 0.f ? -1.f : 1.f; }
  }
  float ebest = -3.0e38f; int ecode = -1; V3 en = v3(0, 0, 0);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      V3 L = cross(A[i], B[j]); float len = norm(L);
      if (len < 1e-6f) continue;
      L = L * (1.0f / len);
      float ra = s1[0] * fabsf(dot(A[0], L)) + s1[1] * fabsf(dot(A[1], L)) + s1[2] * fabsf(dot(A[2], L));
      float rb = s2[0] * fabsf(dot(B[0], L)) + s2[1] * fabsf(dot(B[1], L)) + s2[2] * fabsf(dot(B[2], L));
      float pl = dot(p, L), s = fabsf(pl) - (ra + rb);
      if (s > margin) return 0;
      if (s > ebest) { ebest = s; ecode = 3 * i + j; en = L * (pl < 0.f ? -1.f : 1.f); }
    }
  if (ecode >= 0 && ebest > best + 1e-4f + 0.02f * fabsf(best)) {
    int i = ecode / 3, j = ecode - 3 * i;
    V3 ea = p1, eb = p2, ua = v3(0, 0, 0), ub = v3(0, 0, 0); float la = 0.f, lb = 0.f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (a == i) { ua = A[a]; la = s1[a]; } else ea = ea + A[a] * (dot(en, A[a]) > 0.f ? s1[a] : -s1[a]);
      if (a == j) { ub = B[a]; lb = s2[a]; } else eb = eb + B[a] * (dot(en, B[a]) > 0.f ? -s2[a] : s2[a]);
    }
    V3 w = eb - ea; float uaub = dot(ua, ub), q1 = dot(ua, w), q2 = -dot(ub, w), dd = 1.f - uaub * uaub, alpha = 0.f, beta = 0.f;
    if (dd > 1e-12f) { alpha = (q1 + uaub * q2) / dd; beta = (uaub * q1 + q2) / dd; }
    alpha = clampf(alpha, -la, la); beta = clampf(beta, -lb, lb);
    raw_put(dst, ebest, ((ea + ua * alpha) + (eb + ub * beta)) * 0.5f, en, v3(0, 0, 0));
    return 1;
  }
  // face case: reference box owns the axis, the incident face of the other box is clipped against its side planes
  const bool ref1 = code < 3; const int fi = ref1 ? code : code - 3;
  V3 Rf[3], If[3]; float rs[3], is[3]; V3 rp, ip;
#pragma unroll
  for (int k = 0; k < 3; k++) { Rf[k] = ref1 ? A[k] : B[k]; If[k] = ref1 ? B[k] : A[k]; rs[k] = ref1 ? s1[k] : s2[k]; is[k] = ref1 ? s2[k] : s1[k]; }
  rp = ref1 ? p1 : p2; ip = ref1 ? p2 : p1;
  V3 rfa = fi == 0 ? Rf[0] : (fi == 1 ? Rf[1] : Rf[2]);
  V3 nout = rfa * (ref1 ? bsign : -bsign);
  float rsf = fi == 0 ? rs[0] : (fi == 1 ? rs[1] : rs[2]);
  int ii = 0; float mn = 3.0e38f, isg = 1.f;
#pragma unroll
  for (int k = 0; k < 3; k++) { float dn = dot(If[k], nout); if (-fabsf(dn) < mn) { mn = -fabsf(dn); ii = k; isg = dn > 0.f ? -1.f : 1.f; } }
  V3 ia = ii == 0 ? If[0] : (ii == 1 ? If[1] : If[2]), ib = ii == 0 ? If[1] : (ii == 1 ? If[2] : If[0]), ic = ii == 0 ? If[2] : (ii == 1 ? If[0] : If[1]);
  float sa = ii == 0 ? is[0] : (ii == 1 ? is[1] : is[2]), sb = ii == 0 ? is[1] : (ii == 1 ? is[2] : is[0]), sc = ii == 0 ? is[2] : (ii == 1 ? is[0] : is[1]);
  V3 r1a = fi == 0 ? Rf[1] : (fi == 1 ? Rf[2] : Rf[0]), r2a = fi == 0 ? Rf[2] : (fi == 1 ? Rf[0] : Rf[1]);
  float r1s = fi == 0 ? rs[1] : (fi == 1 ? rs[2] : rs[0]), r2s = fi == 0 ? rs[2] : (fi == 1 ? rs[0] : rs[1]);
  float* X = dst + 56; float* Y = dst + 32;
  V3 fc = (ip - rp) + ia * (isg * sa);
  st3(X + 0, fc + ib * sb + ic * sc); st3(X + 3, fc - ib * sb + ic * sc); st3(X + 6, fc - ib * sb - ic * sc); st3(X + 9, fc + ib * sb - ic * sc);
  int np = 4;
  for (int pl = 0; pl < 4 && np > 0; pl++) {
    V3 axv = pl < 2 ? r1a : r2a; float sg = (pl & 1) ? -1.f : 1.f, lim = pl < 2 ? r1s : r2s;
    const float* src = (pl & 1) ? Y : X; float* out = (pl & 1) ? X : Y; int nn = 0;
    for (int v = 0; v < np; v++) {
      V3 a = ld3(src + 3 * v), b = ld3(src + 3 * ((v + 1 == np) ? 0 : v + 1));
      float da = sg * dot(axv, a) - lim, db = sg * dot(axv, b) - lim;
      if (da <= 0.f && nn < 8) { st3(out + 3 * nn, a); nn++; }
      if (((da < 0.f && db > 0.f) || (da > 0.f && db < 0.f)) && nn < 8) { float t = da / (da - db); st3(out + 3 * nn, a + (b - a) * t); nn++; }
    }
    np = nn;
  }
  int cnt = 0; float fsign = ref1 ? 1.f : -1.f;
  for (int v = 0; v < np; v++) {
    V3 q = ld3(X + 3 * v);
    float depth = dot(q, nout) - rsf;
    if (depth > margin) continue;
    raw_put(dst + B2_RAW * cnt, depth, q + rp - nout * (depth * 0.5f), nout * fsign, v3(0, 0, 0));
    cnt++;
  }
  return cnt;
}

}  // namespace b2
#include "b2_mpr.cuh"
namespace b2 {

// ---- dispatch: geom 1 has the lower type id.  CVX = false compiles the convex (MPR) call out: the mere presence of that
// out-of-line fp64 function constrains the register allocation of the whole narrow-phase loop (quadruped: -4 %), so kernels of
// models without convex candidate pairs are built without it (b2_batch_create checks the model against the task's trait)
template <bool CVX = true>
__device__ __forceinline__ int collide_pair(int t1, int t2, V3 p1, const float* m1, const float* s1, V3 p2, const float* m2,
                                            const float* s2, float margin, float* dst, int maxn) {
  if (t1 == GT_PLANE) {
    V3 n = matcol(m1, 2);
    if (t2 == GT_SPHERE) return c_plane_sphere(dst, p1, n, p2, s2[0], margin, v3(0, 0, 0));
    if (t2 == GT_CAPSULE) return c_plane_capsule(dst, p1, n, p2, m2, s2, margin);
    if (t2 == GT_BOX) return c_plane_box(dst, p1, n, p2, m2, s2, margin, maxn);
    if (t2 == GT_CYLINDER) return c_plane_cylinder(dst, p1, n, p2, m2, s2, margin, maxn);
    return 0;
  }
  if (t1 == GT_SPHERE) {
    if (t2 == GT_SPHERE) return c_sphere_sphere(dst, p1, s1[0], p2, s2[0], margin);
    if (t2 == GT_CAPSULE) return c_sphere_capsule(dst, p1, s1[0], p2, m2, s2, margin);
    if (t2 == GT_BOX) return c_sphere_box(dst, p1, s1[0], p2, m2, s2, margin);
    if (t2 == GT_CYLINDER) return c_sphere_cylinder(dst, p1, s1[0], p2, m2, s2, margin);
    return 0;
  }
  if (t1 == GT_CAPSULE) {
    if (t2 == GT_CAPSULE) return c_capsule_capsule(dst, p1, m1, s1, p2, m2, s2, margin);
    if (t2 == GT_BOX) return c_capsule_box(dst, p1, m1, s1, p2, m2, s2, margin);
    if (t2 == GT_CYLINDER) { if (!CVX) return 0; return convex_far_apart(t1, p1, m1, s1, t2, p2, m2, s2, margin) ? 0 : c_convex_mpr(dst, t1, p1, m1, s1, t2, p2, m2, s2, margin); }
    return 0;
  }
  if (CVX && t1 == GT_CYLINDER && (t2 == GT_BOX || t2 == GT_CYLINDER))
    return convex_far_apart(t1, p1, m1, s1, t2, p2, m2, s2, margin) ? 0 : c_convex_mpr(dst, t1, p1, m1, s1, t2, p2, m2, s2, margin);
  if (t1 == GT_BOX && t2 == GT_BOX) return c_box_box(dst, p1, m1, s1, p2, m2, s2, margin);
  return 0;
}

}  // namespace b2
