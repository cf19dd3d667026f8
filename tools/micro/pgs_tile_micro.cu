// Micro-benchmark of the 4-row-block PGS sweep on a tiled symmetric A (same code shape as Engine::sweep_block):
// cycles per 4-row block for several orderings of the same arithmetic, vs warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
extern __shared__ __align__(16) float sm[];
__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }
__device__ __forceinline__ int a_index(int i, int j) { int t = tri(i >> 2) + (j >> 2); return 16 * t + 4 * ((i & 3) ^ ((t >> 1) & 3)) + (j & 3); }
struct Rows { float f[4], r[4], ainv[4], ad[4]; };

__device__ __forceinline__ void tile_raw(const float* A, int M, int b, float4 (&q)[4], bool& lower) {
  lower = M >= b; const int t = lower ? tri(M) + b : tri(b) + M; const int sw = (t >> 1) & 3;
  const float4* p = reinterpret_cast<const float4*>(A + 16 * t);
  q[0] = p[sw]; q[1] = p[1 ^ sw]; q[2] = p[2 ^ sw]; q[3] = p[3 ^ sw];
}
__device__ __forceinline__ void tile_sel(const float4 (&q)[4], bool lower, float (&C)[4][4]) {
  C[0][0] = q[0].x; C[1][1] = q[1].y; C[2][2] = q[2].z; C[3][3] = q[3].w;
  C[0][1] = lower ? q[0].y : q[1].x; C[1][0] = lower ? q[1].x : q[0].y;
  C[0][2] = lower ? q[0].z : q[2].x; C[2][0] = lower ? q[2].x : q[0].z;
  C[0][3] = lower ? q[0].w : q[3].x; C[3][0] = lower ? q[3].x : q[0].w;
  C[1][2] = lower ? q[1].z : q[2].y; C[2][1] = lower ? q[2].y : q[1].z;
  C[1][3] = lower ? q[1].w : q[3].y; C[3][1] = lower ? q[3].y : q[1].w;
  C[2][3] = lower ? q[2].w : q[3].z; C[3][2] = lower ? q[3].z : q[2].w;
}
// ---- V0: as in the engine (tile for b+1 loaded and selected at the top of block b)
__device__ __forceinline__ void block_v0(Rows& w, const float* A, int M, int b, int bn, float (&C)[4][4], float (&Cn)[4][4], float& imp, int lane) {
  { float4 q[4]; bool lo; tile_raw(A, M, bn, q, lo); tile_sel(q, lo, Cn); }
  float g0 = fmaf(-w.r[0], w.ainv[0], w.f[0]), g1 = fmaf(-w.r[1], w.ainv[1], w.f[1]);
  float g2 = fmaf(-w.r[2], w.ainv[2], w.f[2]), g3 = fmaf(-w.r[3], w.ainv[3], w.f[3]);
  float c10 = C[1][0] * w.ainv[1], c20 = C[2][0] * w.ainv[2], c21 = C[2][1] * w.ainv[2];
  float c30 = C[3][0] * w.ainv[3], c31 = C[3][1] * w.ainv[3], c32 = C[3][2] * w.ainv[3];
  float h1 = fmaf(c10, w.f[0], g1), h2 = fmaf(c21, w.f[1], fmaf(c20, w.f[0], g2));
  float h3 = fmaf(c32, w.f[2], fmaf(c31, w.f[1], fmaf(c30, w.f[0], g3)));
  float n0 = fmaxf(g0, 0.f), e0 = n0 - w.f[0];
  float d0 = __shfl_sync(FULL, e0, b);
  float p1 = fmaf(-c10, n0, h1), n1 = fmaxf(p1, 0.f), e1 = n1 - w.f[1];
  float d1 = __shfl_sync(FULL, e1, b);
  float p2 = fmaf(-c21, n1, fmaf(-c20, n0, h2)), n2 = fmaxf(p2, 0.f), e2 = n2 - w.f[2];
  float d2 = __shfl_sync(FULL, e2, b);
  float p3 = fmaf(-c32, n2, fmaf(-c31, n1, fmaf(-c30, n0, h3))), n3 = fmaxf(p3, 0.f), e3 = n3 - w.f[3];
  float d3 = __shfl_sync(FULL, e3, b);
  float ch = e0 * w.ad[0] * fmaf(0.5f, e0, w.f[0] - g0) + e1 * w.ad[1] * fmaf(0.5f, e1, w.f[1] - p1) +
             e2 * w.ad[2] * fmaf(0.5f, e2, w.f[2] - p2) + e3 * w.ad[3] * fmaf(0.5f, e3, w.f[3] - p3);
  const bool own = lane == b;
  imp -= own ? ch : 0.f;
  w.f[0] = own ? n0 : w.f[0]; w.f[1] = own ? n1 : w.f[1]; w.f[2] = own ? n2 : w.f[2]; w.f[3] = own ? n3 : w.f[3];
#pragma unroll
  for (int q = 0; q < 4; q++) w.r[q] = fmaf(C[q][3], d3, fmaf(C[q][2], d2, fmaf(C[q][1], d1, fmaf(C[q][0], d0, w.r[q]))));
}
__device__ __forceinline__ void sweep_v0(Rows& w, int n, const float* A, float& imp, int lane) {
  const int nb = (n + 3) >> 2; const int M = min(lane, nb - 1);
  float Ca[4][4], Cb[4][4];
  { float4 q[4]; bool lo; tile_raw(A, M, 0, q, lo); tile_sel(q, lo, Ca); }
#pragma unroll 1
  for (int b = 0; b < nb; b += 2) {
    block_v0(w, A, M, b, min(b + 1, nb - 1), Ca, Cb, imp, lane);
    if (b + 1 < nb) block_v0(w, A, M, b + 1, min(b + 2, nb - 1), Cb, Ca, imp, lane);
  }
}
// ---- V1: state is H_k = f_k - r_k/A_kk + sum_{j<k} c_kj f_j (the owner's chain starts from it directly); the scaled
// tile coefficients E[k][j] = C[k][j] / A_kk (own row k) are formed off the chain; raw tile loads run two blocks ahead.
struct RowsH { float f[4], H[4], ainv[4], ad[4], c[6]; };
__device__ __forceinline__ void block_v1(RowsH& w, const float* A, int M, int b, int b2, float4 (&qc)[4], bool loc, float4 (&q2)[4], bool& lo2,
                                         float& imp, int lane) {
  tile_raw(A, M, b2, q2, lo2);                         // two blocks ahead: latency fully hidden
  // owner chain straight from H
  float n0 = fmaxf(w.H[0], 0.f), e0 = n0 - w.f[0];
  float d0 = __shfl_sync(FULL, e0, b);
  float p1 = fmaf(-w.c[0], n0, w.H[1]), n1 = fmaxf(p1, 0.f), e1 = n1 - w.f[1];
  float d1 = __shfl_sync(FULL, e1, b);
  float p2 = fmaf(-w.c[2], n1, fmaf(-w.c[1], n0, w.H[2])), n2 = fmaxf(p2, 0.f), e2 = n2 - w.f[2];
  float d2 = __shfl_sync(FULL, e2, b);
  float p3 = fmaf(-w.c[5], n2, fmaf(-w.c[4], n1, fmaf(-w.c[3], n0, w.H[3]))), n3 = fmaxf(p3, 0.f), e3 = n3 - w.f[3];
  float d3 = __shfl_sync(FULL, e3, b);
  // coefficients of this lane's rows against block b (selected while the shuffles are in flight)
  float C[4][4]; tile_sel(qc, loc, C);
  const bool own = lane == b;
  float ch = e0 * w.ad[0] * fmaf(0.5f, e0, w.f[0] - w.H[0]) + e1 * w.ad[1] * fmaf(0.5f, e1, w.f[1] - p1) +
             e2 * w.ad[2] * fmaf(0.5f, e2, w.f[2] - p2) + e3 * w.ad[3] * fmaf(0.5f, e3, w.f[3] - p3);
  imp -= own ? ch : 0.f;
  w.f[0] = own ? n0 : w.f[0]; w.f[1] = own ? n1 : w.f[1]; w.f[2] = own ? n2 : w.f[2]; w.f[3] = own ? n3 : w.f[3];
  // H_k -= ainv_k * sum_j C[k][j] d_j ; for the owner only the strict upper triangle of the diagonal tile remains
  float s0 = fmaf(C[0][3], d3, fmaf(C[0][2], d2, fmaf(C[0][1], d1, own ? 0.f : C[0][0] * d0)));
  float s1 = fmaf(C[1][3], d3, fmaf(C[1][2], d2, own ? 0.f : fmaf(C[1][1], d1, C[1][0] * d0)));
  float s2 = fmaf(C[2][3], d3, own ? 0.f : fmaf(C[2][2], d2, fmaf(C[2][1], d1, C[2][0] * d0)));
  float s3 = own ? 0.f : fmaf(C[3][3], d3, fmaf(C[3][2], d2, fmaf(C[3][1], d1, C[3][0] * d0)));
  w.H[0] = fmaf(-w.ainv[0], s0, w.H[0]); w.H[1] = fmaf(-w.ainv[1], s1, w.H[1]);
  w.H[2] = fmaf(-w.ainv[2], s2, w.H[2]); w.H[3] = fmaf(-w.ainv[3], s3, w.H[3]);
}
__device__ __forceinline__ void sweep_v1(RowsH& w, int n, const float* A, float& imp, int lane) {
  const int nb = (n + 3) >> 2; const int M = min(lane, nb - 1);
  float4 qa[4], qb[4], qc[4]; bool la, lb, lc;
  tile_raw(A, M, 0, qa, la); tile_raw(A, M, min(1, nb - 1), qb, lb);
#pragma unroll 1
  for (int b = 0; b < nb; b += 3) {
    block_v1(w, A, M, b, min(b + 2, nb - 1), qa, la, qc, lc, imp, lane);
    if (b + 1 < nb) block_v1(w, A, M, b + 1, min(b + 3, nb - 1), qb, lb, qa, la, imp, lane);
    if (b + 2 < nb) block_v1(w, A, M, b + 2, min(b + 4, nb - 1), qc, lc, qb, lb, imp, lane);
  }
}

// ---- V2: H-state, raw tile for block b+1 loaded at the top of block b into the other of two buffers (no rotation),
// selected after the shuffles are issued
__device__ __forceinline__ void block_v2(RowsH& w, const float* A, int M, int b, int bn, const float4 (&qc)[4], bool loc, float4 (&qn)[4], bool& lon,
                                         float& imp, int lane) {
  tile_raw(A, M, bn, qn, lon);
  float n0 = fmaxf(w.H[0], 0.f), e0 = n0 - w.f[0];
  float d0 = __shfl_sync(FULL, e0, b);
  float p1 = fmaf(-w.c[0], n0, w.H[1]), n1 = fmaxf(p1, 0.f), e1 = n1 - w.f[1];
  float d1 = __shfl_sync(FULL, e1, b);
  float p2 = fmaf(-w.c[2], n1, fmaf(-w.c[1], n0, w.H[2])), n2 = fmaxf(p2, 0.f), e2 = n2 - w.f[2];
  float d2 = __shfl_sync(FULL, e2, b);
  float p3 = fmaf(-w.c[5], n2, fmaf(-w.c[4], n1, fmaf(-w.c[3], n0, w.H[3]))), n3 = fmaxf(p3, 0.f), e3 = n3 - w.f[3];
  float d3 = __shfl_sync(FULL, e3, b);
  float C[4][4]; tile_sel(qc, loc, C);
  const bool own = lane == b;
  float ch = e0 * w.ad[0] * fmaf(0.5f, e0, w.f[0] - w.H[0]) + e1 * w.ad[1] * fmaf(0.5f, e1, w.f[1] - p1) +
             e2 * w.ad[2] * fmaf(0.5f, e2, w.f[2] - p2) + e3 * w.ad[3] * fmaf(0.5f, e3, w.f[3] - p3);
  imp -= own ? ch : 0.f;
  w.f[0] = own ? n0 : w.f[0]; w.f[1] = own ? n1 : w.f[1]; w.f[2] = own ? n2 : w.f[2]; w.f[3] = own ? n3 : w.f[3];
  float s0 = fmaf(C[0][3], d3, fmaf(C[0][2], d2, fmaf(C[0][1], d1, own ? 0.f : C[0][0] * d0)));
  float s1 = fmaf(C[1][3], d3, fmaf(C[1][2], d2, own ? 0.f : fmaf(C[1][1], d1, C[1][0] * d0)));
  float s2 = fmaf(C[2][3], d3, own ? 0.f : fmaf(C[2][2], d2, fmaf(C[2][1], d1, C[2][0] * d0)));
  float s3 = own ? 0.f : fmaf(C[3][3], d3, fmaf(C[3][2], d2, fmaf(C[3][1], d1, C[3][0] * d0)));
  w.H[0] = fmaf(-w.ainv[0], s0, w.H[0]); w.H[1] = fmaf(-w.ainv[1], s1, w.H[1]);
  w.H[2] = fmaf(-w.ainv[2], s2, w.H[2]); w.H[3] = fmaf(-w.ainv[3], s3, w.H[3]);
}
__device__ __forceinline__ void sweep_v2(RowsH& w, int n, const float* A, float& imp, int lane) {
  const int nb = (n + 3) >> 2; const int M = min(lane, nb - 1);
  float4 qa[4], qb[4]; bool la, lb;
  tile_raw(A, M, 0, qa, la);
  const int nb2 = nb & ~1;
#pragma unroll 1
  for (int b = 0; b < nb2; b += 2) {
    block_v2(w, A, M, b, b + 1, qa, la, qb, lb, imp, lane);
    block_v2(w, A, M, b + 1, min(b + 2, nb - 1), qb, lb, qa, la, imp, lane);
  }
  if (nb & 1) block_v2(w, A, M, nb - 1, nb - 1, qa, la, qb, lb, imp, lane);
}

template <int V>
__global__ void k(int n, int iters, float* out, long long* cyc) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* A = sm + warp * 5120;
  for (int i = lane; i < 5120; i += 32) A[i] = 0.f;
  __syncwarp();
  for (int i = lane; i < n; i += 32) for (int j = 0; j <= i; j++) { float v = i == j ? 5.f : 0.001f * (((i * 7 + j) % 13) - 6); A[a_index(i, j)] = v; if ((i >> 2) == (j >> 2)) A[a_index(j, i)] = v; }
  __syncwarp();
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    float imp = 0.f;
    if (V == 0) {
      Rows w;
#pragma unroll
      for (int q = 0; q < 4; q++) { int i = 4 * lane + q; bool v = i < n; w.f[q] = v ? 0.1f * i : 0.f; w.r[q] = v ? 0.3f - 0.01f * i + acc * 1e-9f : 0.f; w.ad[q] = v ? A[a_index(i, i)] : 1.f; w.ainv[q] = 1.f / w.ad[q]; }
      sweep_v0(w, n, A, imp, lane);
      acc += imp + w.r[0] + w.f[0];
    } else {
      RowsH w;
#pragma unroll
      for (int q = 0; q < 4; q++) { int i = 4 * lane + q; bool v = i < n; w.f[q] = v ? 0.1f * i : 0.f; float r = v ? 0.3f - 0.01f * i + acc * 1e-9f : 0.f; w.ad[q] = v ? A[a_index(i, i)] : 1.f; w.ainv[q] = 1.f / w.ad[q]; w.H[q] = w.f[q] - r * w.ainv[q]; }
      int i0 = 4 * min(lane, ((n + 3) >> 2) - 1);
      w.c[0] = A[a_index(i0 + 1, i0)] * w.ainv[1]; w.c[1] = A[a_index(i0 + 2, i0)] * w.ainv[2]; w.c[2] = A[a_index(i0 + 2, i0 + 1)] * w.ainv[2];
      w.c[3] = A[a_index(i0 + 3, i0)] * w.ainv[3]; w.c[4] = A[a_index(i0 + 3, i0 + 1)] * w.ainv[3]; w.c[5] = A[a_index(i0 + 3, i0 + 2)] * w.ainv[3];
      w.H[1] = fmaf(w.c[0], w.f[0], w.H[1]); w.H[2] = fmaf(w.c[2], w.f[1], fmaf(w.c[1], w.f[0], w.H[2]));
      w.H[3] = fmaf(w.c[5], w.f[2], fmaf(w.c[4], w.f[1], fmaf(w.c[3], w.f[0], w.H[3])));
      if (V == 1) sweep_v1(w, n, A, imp, lane); else sweep_v2(w, n, A, imp, lane);
      acc += imp + w.H[0] + w.f[0];
    }
  }
  long long t1 = clock64();
  if (lane == 0) { out[blockIdx.x * (blockDim.x / 32) + warp] = acc; if (blockIdx.x == 0 && warp == 0) *cyc = t1 - t0; }
}
template <int V> void run(const char* name, float* out, long long* cyc) {
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  for (int n : {40, 72}) for (int warps : {1, 6}) {
    int iters = 200;
    k<V><<<148, 32 * warps, warps * 20480, 0>>>(n, iters, out, cyc); cudaDeviceSynchronize();
    k<V><<<148, 32 * warps, warps * 20480, 0>>>(n, iters, out, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    float o; cudaMemcpy(&o, out, 4, cudaMemcpyDeviceToHost);
    printf("%s n=%3d warps/SM=%2d: %8.1f cycles/iter, %6.1f cycles/block, %5.1f cycles/row  acc=%g (%s)\n", name, n, warps, (double)h / iters,
           (double)h / iters / ((n + 3) / 4), (double)h / iters / n, o, cudaGetErrorString(cudaGetLastError()));
  }
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  run<0>("v0-engine ", out, cyc);
  run<1>("v1-Hstate ", out, cyc);
  run<2>("v2-pingpong", out, cyc);
  return 0;
}
