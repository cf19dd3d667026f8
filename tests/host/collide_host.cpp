// Host build of the product's fp32 narrow-phase primitives (csrc/b2_collide.cuh) for CPU-side differential tests
// against the fp64 oracle.  Compiled by tests/test_collide_host.py with g++; the CUDA qualifiers are defined away.
#include <math.h>
#include <string.h>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define B2_HOST_BUILD 1
static inline void sincosf_(float a, float* s, float* c) { *s = sinf(a); *c = cosf(a); }
#include "../../mujoco_gymnasium_environments_b200/csrc/b2_collide.cuh"

extern "C" int h_collide_pair(int t1, int t2, const float* p1, const float* m1, const float* s1, const float* p2,
                              const float* m2, const float* s2, float margin, float* out80) {
  using namespace b2;
  return collide_pair(t1, t2, ld3(p1), m1, s1, ld3(p2), m2, s2, margin, out80, 8);
}
// the convex (MPR) path on any pair of sphere / capsule / cylinder / box
extern "C" int h_mpr_pair(int t1, int t2, const float* p1, const float* m1, const float* s1, const float* p2,
                          const float* m2, const float* s2, float margin, float* out10) {
  using namespace b2;
  return c_convex_mpr(out10, t1, ld3(p1), m1, s1, t2, ld3(p2), m2, s2, margin);
}
