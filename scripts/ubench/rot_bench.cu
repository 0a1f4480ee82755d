// Microbenchmark of the single-warp look-ahead chain routines (development aid): hot timings, alone
// and with the other 15 warps hammering shared memory (the trailing update's access pattern).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lcp_physics_b200/csrc/lcp_lu.cuh"
using namespace lcpb200;

template <typename T, int NBX>
__global__ void __launch_bounds__(512, 1) bench(long long* out, int reps, int ld, int contend) {
  T* sm = smem_base<T>();
  const int o_rdiag = 64 * ld, o_stage = o_rdiag + 64, o_perm = o_stage + 128, o_flag = o_perm + 64, o_lt = o_flag + 8, o_junk = o_lt + 1200;
  MPtr<T, 0> D; D.off = 0; D.g = nullptr;
  auto fill = [&]() {
    for (int t = threadIdx.x; t < 64 * ld; t += blockDim.x) {
      int i = t / ld, j = t % ld;
      sm[t] = (i == j) ? T(4) + T(0.01) * i : T(0.3) / (1 + ((i * 7 + j * 13) % 11));
    }
    __syncthreads();
  };
  __shared__ volatile int stop;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) out[i] = 0; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = 0; r < reps; ++r) {
    fill();
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (warp == 15) {
      long long t0 = clock64();
      diag_lu_rot<T, 0, NBX>(D, ld, o_perm * (int)(sizeof(T) / 4), o_rdiag, o_flag * (int)(sizeof(T) / 4), o_stage, o_lt, NBX + 4, 0);
      long long t1 = clock64();
      lookahead_u12<T, 0, NBX>(D, ld, o_lt, NBX + 4);
      long long t2 = clock64();
      lookahead_l21<T, 0, NBX>(D, ld, o_rdiag);
      long long t3 = clock64();
      diag_lu_rot<T, 0, NBX>(D.plus((long long)NBX * ld + NBX), ld, o_perm * (int)(sizeof(T) / 4), o_rdiag, o_flag * (int)(sizeof(T) / 4), o_stage, o_lt, NBX + 4, 1);
      long long t4 = clock64();
      if (lane == 0) { out[0] += t1 - t0; out[1] += t2 - t1; out[2] += t3 - t2; out[3] += t4 - t3; stop = 1; }
    } else if (contend) {
      // bulk-like traffic: vector loads + FMAs on a private junk area until the chain warp is done
      using V = typename VecOf<T>::type;
      constexpr int VC = VecOf<T>::VC;
      T acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const T* base = sm + o_junk + (threadIdx.x & ~7) * VC;
      int it = 0;
      while (!stop && it < 200000) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          T v[VC];
          vec_get<T>(*reinterpret_cast<const V*>(base + ((it + q) & 15) * 512 * VC / 8), v);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fma(v[k % VC], acc[(k + 1) & 7], acc[k]);
        }
        ++it;
      }
      if (acc[0] == T(12345.678)) out[7] = 1;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[6] = (long long)(sm[5 * ld + 7] * 1000);
}

template <typename T, int NBX>
void run(const char* name, int ld, long long* d) {
  long long h[8];
  const int reps = 20, smem = 200 * 1024;
  cudaFuncSetAttribute(bench<T, NBX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int contend = 0; contend < 2; ++contend) {
    bench<T, NBX><<<1, 512, smem>>>(d, reps, ld, contend);
    cudaDeviceSynchronize(); cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("%s NB=%d contend=%d: diag %lld  u12 %lld  l21 %lld  diag+fused_update %lld  (chk %lld) %s\n", name, NBX, contend,
           h[0] / reps, h[1] / reps, h[2] / reps, h[3] / reps, h[6], cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  long long* d; cudaMalloc(&d, 128);
  run<float, 32>("float ", 260, d);
  run<float, 16>("float ", 260, d);
  run<double, 16>("double", 162, d);
  return 0;
}
