// Microbenchmark of the rolled single-warp diagonal-block LU (development aid): where do the cycles go?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lcp_physics_b200/csrc/lcp_lu.cuh"
using namespace lcpb200;

// instrumented copy of diag_lu_rot's main loop (no pivot search): tm[0]=stage, tm[1]=test+rcp, tm[2]=update
template <typename T, int NB, int VARIANT>
__device__ __noinline__ void rot_probe(T* D, int ld, T* stage, long long* tm) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, NV = NB / VC, SL = NB + VC;
  const int lane = threadIdx.x & 31;
  const int li = lane < NB ? lane : NB - 1;
  T* row = D + (size_t)li * ld;
  T a[NB];
  T rm = 0;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    T t[VC];
    vec_get<T>(*reinterpret_cast<const V*>(row + c * VC), t);
#pragma unroll
    for (int q = 0; q < VC; ++q) { a[c * VC + q] = t[q]; rm = fmax(rm, fabs(t[q])); }
  }
  const T tau = T(1e-4);
  T myr = 0;
  T u[NB];
  long long ta = 0, tb = 0, tc = 0;
#pragma unroll 1
  for (int k = 0; k < NB; ++k) {
    long long c0 = clock64();
    T* const st = stage + (k & 1) * SL;
    if (lane == k) {
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        T t[VC];
#pragma unroll
        for (int q = 0; q < VC; ++q) t[q] = a[c * VC + q];
        *reinterpret_cast<V*>(st + c * VC) = vec_make(t);
      }
      st[NB] = rm;
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      T t[VC];
      vec_get<T>(*reinterpret_cast<const V*>(st + c * VC), t);
#pragma unroll
      for (int q = 0; q < VC; ++q) u[c * VC + q] = t[q];
    }
    T piv = u[0];
    const T rs = st[NB];
    long long c1 = clock64();
    bool bad = !(fabs(piv) >= tau * rs && fabs(piv) > T(0));
    if (bad) piv = T(1);
    const T r = fast_rcp(piv);
    if (lane == k) myr = r;
    const bool alive = lane > k && lane < NB;
    const T l = a[0] * r;
    long long c2 = clock64();
    if (VARIANT == 0) {
      if (alive) {
        D[(size_t)lane * ld + k] = l;
#pragma unroll
        for (int j = 1; j < NB; ++j) a[j - 1] = fma(-l, u[j], a[j]);
      }
    } else {
      if (alive) D[(size_t)lane * ld + k] = l;
      const T ll = alive ? l : T(0);
      if (lane != k || true) {
#pragma unroll
        for (int j = 1; j < NB; ++j) { T v = fma(-ll, u[j], a[j]); a[j - 1] = (lane <= k) ? a[j - 1] : v; }
      }
    }
    long long c3 = clock64();
    ta += c1 - c0; tb += c2 - c1; tc += c3 - c2;
  }
  if (lane < NB) {
#pragma unroll
    for (int j = 0; j < NB; ++j)
      if (lane + j < NB) D[(size_t)lane * ld + lane + j] = a[j];
    stage[2 * SL + lane] = myr;
  }
  if (lane == 0) { tm[0] += ta; tm[1] += tb; tm[2] += tc; }
}

template <typename T>
__global__ void __launch_bounds__(512, 1) bench(long long* out, int reps, int ld) {
  constexpr int NB = Blk<T>::NB;
  T* sm = smem_base<T>();
  const int o_blk = 0, o_rdiag = NB * ld, o_stage = o_rdiag + NB, o_perm = o_stage + 128, o_flag = o_perm + NB;
  MPtr<T, 0> D; D.off = o_blk; D.g = nullptr;
  auto fill = [&]() {
    for (int t = threadIdx.x; t < NB * ld; t += blockDim.x) {
      int i = t / ld, j = t % ld;
      sm[t] = (i == j) ? T(4) + T(0.01) * i : T(0.3) / (1 + ((i * 7 + j * 13) % 11));
    }
    __syncthreads();
  };
  if (threadIdx.x == 0) for (int i = 0; i < 16; ++i) out[i] = 0;
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    fill();
    long long t0 = clock64();
    if (threadIdx.x < 32) diag_lu_rot<T, 0, NB>(D, ld, o_perm * (int)(sizeof(T) / 4), o_rdiag, o_flag * (int)(sizeof(T) / 4), o_stage);
    long long t1 = clock64();
    __syncthreads();
    long long t2 = clock64();
    diag_inverse_job<T, 0, NB>(D, ld, o_rdiag, (threadIdx.x >> 5) & 1, (threadIdx.x >> 5) < 2);
    long long t3 = clock64();
    __syncthreads();
    fill();
    long long t4 = clock64();
    if (threadIdx.x < 32) rot_probe<T, NB, 0>(sm, ld, sm + o_stage, out + 4);
    long long t5 = clock64();
    __syncthreads();
    fill();
    long long t6 = clock64();
    if (threadIdx.x < 32) rot_probe<T, NB, 1>(sm, ld, sm + o_stage, out + 8);
    long long t7 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) { out[0] += t1 - t0; out[1] += t3 - t2; out[2] += t5 - t4; out[3] += t7 - t6; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[12] = (long long)(sm[5 * ld + 7] * 1000);
}

int main() {
  long long* d; cudaMalloc(&d, 128);
  long long h[16];
  const int reps = 20;
  for (int big = 0; big < 2; ++big) {
    int smem = big ? 200 * 1024 : 48 * 1024;
    cudaFuncSetAttribute(bench<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bench<float><<<1, 512, smem>>>(d, reps, 260);
    cudaDeviceSynchronize(); cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
    printf("float  smem=%dK: rot %lld inv %lld | probe0 %lld (stage %lld rcp %lld upd %lld) probe1 %lld (stage %lld rcp %lld upd %lld) chk %lld err=%s\n", smem / 1024,
           h[0] / reps, h[1] / reps, h[2] / reps, h[4] / reps, h[5] / reps, h[6] / reps, h[3] / reps, h[8] / reps, h[9] / reps, h[10] / reps, h[12], cudaGetErrorString(cudaGetLastError()));
    bench<double><<<1, 512, smem>>>(d, reps, 162);
    cudaDeviceSynchronize(); cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
    printf("double smem=%dK: rot %lld inv %lld | probe0 %lld (stage %lld rcp %lld upd %lld) probe1 %lld (stage %lld rcp %lld upd %lld) chk %lld err=%s\n", smem / 1024,
           h[0] / reps, h[1] / reps, h[2] / reps, h[4] / reps, h[5] / reps, h[6] / reps, h[3] / reps, h[8] / reps, h[9] / reps, h[10] / reps, h[12], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
