// Micro-benchmark of the PGS row sweep (same code shape as Engine::sweep_slot): cycles per row vs warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define S_ 3
#define FULL 0xffffffffu
extern __shared__ float sm[];
__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }
struct Chain { int n; const float* A; float f[S_], r[S_], ainv[S_], ad[S_]; int adr[S_]; };
template <int S>
__device__ __forceinline__ void sweep_slot(Chain& c, float& improvement, int lane) {
  int nn = min(32, c.n - 32 * S);
  for (int ii = 0; ii < nn; ii++) {
    const int i = 32 * S + ii;
    float dl = fmaxf(fmaf(-c.r[S], c.ainv[S], c.f[S]), 0.f) - c.f[S];
    float dlb = __shfl_sync(FULL, dl, ii);
    if (lane == ii) { improvement -= dl * fmaf(0.5f * dl, c.ad[S], c.r[S]); c.f[S] += dl; }
#pragma unroll
    for (int s2 = 0; s2 < S_; s2++) {
      const int col = lane + 32 * s2;
      if (col < c.n) c.r[s2] = fmaf(c.A[c.adr[s2]], dlb, c.r[s2]);
      c.adr[s2] += (s2 > S) ? 1 : (s2 < S) ? i + 1 : ((col > i) ? 1 : i + 1);
    }
  }
}
__global__ void k(int n, int iters, float* out, long long* cyc) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* A = sm + warp * 1024;
  for (int i = lane; i < 1024; i += 32) A[i] = 0.001f * ((i * 7) % 13);
  for (int i = lane; i < n; i += 32) A[tri(i) + i] = 5.f;
  __syncwarp();
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    Chain c; c.n = n; c.A = A;
#pragma unroll
    for (int s = 0; s < S_; s++) { int i = lane + 32 * s; bool v = i < n; c.f[s] = v ? 0.1f * i : 0.f; c.r[s] = v ? 0.3f - 0.01f * i + acc * 1e-9f : 0.f; c.ad[s] = v ? A[tri(i) + i] : 1.f; c.ainv[s] = 1.f / c.ad[s]; c.adr[s] = tri(i); }
    float imp = 0.f;
    sweep_slot<0>(c, imp, lane); sweep_slot<1>(c, imp, lane); sweep_slot<2>(c, imp, lane);
    acc += imp + c.r[0] + c.f[0];
  }
  long long t1 = clock64();
  if (lane == 0) { out[blockIdx.x * (blockDim.x / 32) + warp] = acc; if (blockIdx.x == 0 && warp == 0) *cyc = t1 - t0; }
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int n : {8, 16, 33, 40, 64}) for (int warps : {1, 6, 18}) {
    int iters = 200;
    k<<<148, 32 * warps, warps * 4096, 0>>>(n, iters, out, cyc); cudaDeviceSynchronize();
    k<<<148, 32 * warps, warps * 4096, 0>>>(n, iters, out, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("n=%d warps/SM=%d: %.1f cycles/iter, %.1f cycles/row  (%s)\n", n, warps, (double)h / iters, (double)h / iters / n, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
