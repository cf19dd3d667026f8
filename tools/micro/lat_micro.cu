// Latency probes: dependent SHFL chain, dependent FFMA chain, and variants of the PGS row recurrence.
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
extern __shared__ float sm[];
__global__ void k(int mode, int iters, float* out, long long* cyc) {
  int lane = threadIdx.x & 31;
  float* A = sm + (threadIdx.x >> 5) * 2048;
  for (int i = lane; i < 2048; i += 32) A[i] = 0.001f * ((i * 7) % 13) + (i % 41 == 0 ? 5.f : 0.f);
  __syncwarp();
  float x = 0.1f * lane, f = 0.2f, r0 = 0.3f - 0.01f * lane, r1 = 0.1f, ainv = 0.2f, negf = -0.2f, acc = 0.f;
  const int n = 32;
  long long t0 = clock64();
  if (mode == 0) { for (int it = 0; it < iters; it++) for (int ii = 0; ii < n; ii++) x = __shfl_sync(FULL, x, ii) + 1.0f; }
  else if (mode == 1) { for (int it = 0; it < iters; it++) for (int ii = 0; ii < n; ii++) x = fmaf(x, 1.0001f, 0.5f); }
  else if (mode == 2) {   // minimal chain: FMUL, FMNMX, SHFL, FFMA(A from smem, row-major square)
    for (int it = 0; it < iters; it++) for (int ii = 0; ii < n; ii++) {
      float dl = fmaxf(-r0 * ainv, negf); float b = __shfl_sync(FULL, dl, ii);
      r0 = fmaf(A[ii * 32 + lane], b, r0); if (lane == ii) negf -= dl;
    } x = r0 + negf;
  } else if (mode == 3) { // same but broadcast through shared memory instead of SHFL
    volatile float* bc = A + 1500;
    for (int it = 0; it < iters; it++) for (int ii = 0; ii < n; ii++) {
      float dl = fmaxf(-r0 * ainv, negf); if (lane == ii) { bc[ii & 1] = dl; negf -= dl; } __syncwarp();
      float b = bc[ii & 1]; r0 = fmaf(A[ii * 32 + lane], b, r0);
    } x = r0 + negf;
  } else if (mode == 4) { // 2 rows per step handled by the owner pair analytically (block of 2): lane ii owns rows 2ii, 2ii+1
    float ra = r0, rb = r1, fa = f, fb = f, aab = 0.01f, ainva = ainv, ainvb = ainv;
    for (int it = 0; it < iters; it++) for (int ii = 0; ii < n / 2; ii++) {
      float da = fmaxf(-ra * ainva, -fa); float rb2 = fmaf(aab, da, rb); float db = fmaxf(-rb2 * ainvb, -fb);
      float ba = __shfl_sync(FULL, da, ii), bb = __shfl_sync(FULL, db, ii);
      if (lane == ii) { fa += da; fb += db; }
      const float* Ar = A + (2 * ii) * 64;
      ra = fmaf(Ar[lane], ba, ra); rb = fmaf(Ar[32 + lane], ba, rb); ra = fmaf(Ar[64 + lane], bb, ra); rb = fmaf(Ar[96 + lane], bb, rb);
    } x = ra + rb + fa + fb;
  } else if (mode == 5) { // block of 4 rows per owner lane
    float r[4] = {r0, r1, r0 * 0.5f, r1 * 0.5f}, ff[4] = {f, f, f, f};
    for (int it = 0; it < iters; it++) for (int ii = 0; ii < n / 4; ii++) {
      float d[4], rr[4] = {r[0], r[1], r[2], r[3]};
#pragma unroll
      for (int q = 0; q < 4; q++) { d[q] = fmaxf(-rr[q] * ainv, -ff[q]);
#pragma unroll
        for (int q2 = q + 1; q2 < 4; q2++) rr[q2] = fmaf(0.01f * (q + q2), d[q], rr[q2]); }
      float b[4];
#pragma unroll
      for (int q = 0; q < 4; q++) b[q] = __shfl_sync(FULL, d[q], ii);
      if (lane == ii) {
#pragma unroll
        for (int q = 0; q < 4; q++) ff[q] += d[q]; }
      const float* Ar = A + (4 * ii) * 128;
#pragma unroll
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int t = 0; t < 4; t++) r[t] = fmaf(Ar[q * 128 + t * 32 + lane], b[q], r[t]);
    } x = r[0] + r[1] + r[2] + r[3] + ff[0];
  }
  long long t1 = clock64();
  acc = x;
  if (lane == 0) { out[blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)] = acc; if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0; }
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[] = {"dependent SHFL+FADD", "dependent FFMA", "row: FMUL,FMNMX,SHFL,FFMA(LDS)", "row: smem broadcast", "2-row blocks", "4-row blocks"};
  for (int mode = 0; mode < 6; mode++) for (int warps : {1, 18}) {
    int iters = 300;
    k<<<148, 32 * warps, warps * 8192, 0>>>(mode, iters, out, cyc); cudaDeviceSynchronize();
    k<<<148, 32 * warps, warps * 8192, 0>>>(mode, iters, out, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps/SM=%2d: %.1f cycles per row  (%s)\n", names[mode], warps, (double)h / iters / 32, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
