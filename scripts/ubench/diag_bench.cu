// Microbenchmark of the single-warp diagonal-block routines (development aid).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lcp_physics_b200/csrc/lcp_lu.cuh"
using namespace lcpb200;

template <typename T>
__global__ void __launch_bounds__(512, 1) bench(long long* out, int reps, int ld, int big) {
  constexpr int NB = Blk<T>::NB;
  T* sm = smem_base<T>();
  // layout: [0, NB*ld) block; then rmaxs NB; rdiag NB; stage NB*36; perm NB ints; flag
  const int o_blk = 0, o_rmaxs = NB * ld, o_rdiag = o_rmaxs + NB, o_stage = o_rdiag + NB, o_perm = o_stage + NB * 40, o_flag = o_perm + NB;
  MPtr<T, 0> D; D.off = o_blk; D.g = nullptr;
  long long t_lu = 0, t_il = 0, t_iu = 0;
  if (threadIdx.x == 0) { out[4] = 0; out[5] = 0; }
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    for (int t = threadIdx.x; t < NB * ld; t += blockDim.x) {
      int i = t / ld, j = t % ld;
      sm[t] = (i == j) ? T(4) + T(0.01) * i : T(0.3) / (1 + ((i * 7 + j * 13) % 11));
    }
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x < 32) diag_lu_warp<T, 0, NB>(D, ld, o_perm * (int)(sizeof(T) / 4), o_rmaxs, o_rdiag, o_flag * (int)(sizeof(T) / 4));
    long long t1 = clock64();
    __syncthreads();
    long long t2 = clock64();
    if (threadIdx.x < 32) diag_inverse_lower_inplace<T, 0, NB>(D, ld);
    long long t3 = clock64();
    __syncthreads();
    long long t4 = clock64();
    if (threadIdx.x < 32) diag_inverse_upper_inplace<T, 0, NB>(D, ld, o_rdiag);
    long long t5 = clock64();
    __syncthreads();
    for (int t = threadIdx.x; t < NB * ld; t += blockDim.x) {
      int i = t / ld, j = t % ld;
      sm[t] = (i == j) ? T(4) + T(0.01) * i : T(0.3) / (1 + ((i * 7 + j * 13) % 11));
    }
    __syncthreads();
    long long t6 = clock64();
    int kd = 0;
    if (threadIdx.x < 32) kd = diag_lu_regs<T, 0, NB>(D, ld, o_perm * (int)(sizeof(T) / 4), o_rdiag) ? NB : -1;
    long long t7 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) { t_lu += t1 - t0; t_il += t3 - t2; t_iu += t5 - t4; out[4] += t7 - t6; out[5] = kd; }
  }
  if (threadIdx.x == 0) { out[0] = t_lu / reps; out[1] = t_il / reps; out[2] = t_iu / reps; out[3] = (long long)(sm[5 * ld + 7] * 1000); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h[8];
  for (int big = 0; big < 2; ++big) {
    int smem = big ? 200 * 1024 : 48 * 1024;
    cudaFuncSetAttribute(bench<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bench<float><<<1, 512, smem>>>(d, 20, 260, big);
    cudaDeviceSynchronize(); cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    printf("float  NB=32 smem=%dK: lu %lld  inv_lower %lld  inv_upper %lld cycles (chk %lld) | regs-shfl lu %lld (kdone %lld) err=%s\n", smem / 1024, h[0], h[1], h[2], h[3], h[4] / 20, h[5], cudaGetErrorString(cudaGetLastError()));
    bench<double><<<1, 512, smem>>>(d, 20, 162, big);
    cudaDeviceSynchronize(); cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    printf("double NB=16 smem=%dK: lu %lld  inv_lower %lld  inv_upper %lld cycles (chk %lld) | regs-shfl lu %lld (kdone %lld) err=%s\n", smem / 1024, h[0], h[1], h[2], h[3], h[4] / 20, h[5], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
