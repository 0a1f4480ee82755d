// Microbenchmark: dependent-chain LATENCIES (one warp, cycles per op via clock64) of DFMA, FFMA,
// mma.sync F64 (m8n8k4), 64-bit shuffle, LDS.64, MUFU.RCP-based fp64 reciprocal; and throughput of
// DFMA alone / DMMA alone / both interleaved (are they the same pipe?). Development aid.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e));return 1;}}while(0)

__global__ void lat_kernel(double* out, long long* cyc, int iters) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 1) & 1023);
  __syncthreads();
  const int lane = threadIdx.x;
  long long t0, t1;
  double a = 1.0 + lane * 1e-9, b = 1.0000001, c = 1e-9;
  // DFMA chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
  t1 = clock64();
  if (lane == 0) cyc[0] = t1 - t0;
  float f = 1.0f + lane * 1e-6f;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { f = fmaf(f, 1.0000001f, 1e-9f); f = fmaf(f, 1.0000001f, 1e-9f); f = fmaf(f, 1.0000001f, 1e-9f); f = fmaf(f, 1.0000001f, 1e-9f); }
  t1 = clock64();
  if (lane == 0) cyc[1] = t1 - t0;
  // 64-bit shuffle chain
  double s = a;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { s = __shfl_sync(0xffffffffu, s, (lane + 1) & 31); s = __shfl_sync(0xffffffffu, s, (lane + 3) & 31); s = __shfl_sync(0xffffffffu, s, (lane + 5) & 31); s = __shfl_sync(0xffffffffu, s, (lane + 7) & 31); }
  t1 = clock64();
  if (lane == 0) cyc[2] = t1 - t0;
  // LDS.64 pointer chase
  int idx = lane;
  double v = 0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { v = sm[idx]; idx = (int)v; v = sm[idx]; idx = (int)v; v = sm[idx]; idx = (int)v; v = sm[idx]; idx = (int)v; }
  t1 = clock64();
  if (lane == 0) cyc[3] = t1 - t0;
  // DMMA chain (accumulator dependent)
  double c0 = 0, c1 = 0, ma = 1.0 + lane * 1e-3, mb = 0.5;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(ma), "d"(mb));
  }
  t1 = clock64();
  if (lane == 0) cyc[4] = t1 - t0;
  // fp64 reciprocal chain (fp32 seed + 3 Newton steps)
  double r = 1.0 + lane * 1e-3;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float xf = (float)r, rf;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(xf));
      double y = (double)rf;
      y = fma(y, fma(-r, y, 1.0), y); y = fma(y, fma(-r, y, 1.0), y); y = fma(y, fma(-r, y, 1.0), y);
      r = y + 0.5;
    }
  }
  t1 = clock64();
  if (lane == 0) cyc[5] = t1 - t0;
  // IEEE division chain
  double dv = 1.0 + lane * 1e-3;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { dv = 1.0 / dv + 0.5; dv = 1.0 / dv + 0.5; dv = 1.0 / dv + 0.5; dv = 1.0 / dv + 0.5; }
  t1 = clock64();
  if (lane == 0) cyc[6] = t1 - t0;
  out[threadIdx.x] = a + f + s + v + c0 + c1 + r + dv;
}

// throughput: mode 0 DFMA only, 1 DMMA only, 2 both interleaved
__global__ void thr_kernel(double* out, int iters, int mode) {
  double a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001 + i;
  double c[8][2]; for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
  const double b = 1.0001, cc = 0.5, ma = 1.0 + threadIdx.x * 1e-3, mb = 0.5;
  for (int it = 0; it < iters; ++it) {
    if (mode != 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, cc);
    }
    if (mode != 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(ma), "d"(mb));
    }
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i] + c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double* out; long long* cyc; CK(cudaMalloc(&out, 148 * 8 * 1024 * 8)); CK(cudaMalloc(&cyc, 64));
  const int iters = 4096;
  lat_kernel<<<1, 32>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
  lat_kernel<<<1, 32>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
  long long h[8]; CK(cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost));
  const char* nm[7] = {"DFMA", "FFMA", "SHFL.64 (2x SHFL)", "LDS.64 + F2I chase", "DMMA m8n8k4 (acc chain)", "fp64 rcp (seed+3 Newton) + DADD", "fp64 IEEE div + DADD"};
  for (int i = 0; i < 7; ++i) printf("latency %-34s: %.1f cycles/op\n", nm[i], (double)h[i] / (4.0 * iters));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 16; warps *= 2)
    for (int mode = 0; mode < 3; ++mode) {
      const int it2 = 20000;
      thr_kernel<<<sms, warps * 32>>>(out, it2, mode); cudaDeviceSynchronize();
      cudaEventRecord(e0); thr_kernel<<<sms, warps * 32>>>(out, it2, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double dfma = mode != 1 ? (double)sms * warps * 32 * 8.0 * it2 : 0, dmma = mode != 0 ? (double)sms * warps * 8.0 * it2 * 256 : 0;
      printf("throughput %2d warps/SM mode %d (%s): %.3f ms  DFMA %.1f TFLOP/s  DMMA %.1f TFLOP/s\n", warps, mode,
             mode == 0 ? "DFMA" : mode == 1 ? "DMMA" : "both", ms, 2 * dfma / ms / 1e9, 2 * dmma / ms / 1e9);
    }
  return 0;
}
