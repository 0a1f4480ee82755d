// lcp_device.cuh -- device building blocks of the B200 batched LCP solver.
//
// One CTA owns one scene at a time. Matrices are addressed through generic
// pointers (shared or global/L2 workspace -- the host planner decides), vectors
// live in shared memory. Every helper assumes its inputs are visible on entry
// and ends with a __syncthreads() unless its comment says otherwise.
//
// Reference algorithm: lcp_physics/lcp/solvers/pdipm.py (see include/lcpb200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lcpb200 {

constexpr unsigned FULL = 0xffffffffu;

template <typename T> struct Blk { static constexpr int NB = (sizeof(T) == 4) ? 32 : 16; };

// ---------------------------------------------------------------- reductions
// torch.min/torch.max propagate NaN; fmin/fmax do not. Mirror torch.
template <typename T> __device__ __forceinline__ T nan_min(T a, T b) {
  return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
}
template <typename T> __device__ __forceinline__ T nan_max(T a, T b) {
  return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}
struct OpSum { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a + b; } };
struct OpMin { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return nan_min(a, b); } };
struct OpMax { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return nan_max(a, b); } };

template <typename T, typename Op>
__device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// A team = the threads [0, nt) of the CTA that run a phase together, with their own barrier:
// bar 0 = the whole CTA (__syncthreads), otherwise a named barrier over the first nt threads (nt a
// multiple of 32). Lets the bulk of the CTA run a phase while another warp does something else.
struct Team {
  int tid, nt, bar;
  __device__ __forceinline__ void sync() const {
    if (bar == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nt) : "memory");
  }
  __device__ __forceinline__ static Team cta() { Team t; t.tid = threadIdx.x; t.nt = blockDim.x; t.bar = 0; return t; }
};

// Block-wide reduction of up to 6 values at once; result broadcast to all threads.
// red: shared scratch of >= NV*32 elements. Ends with a barrier that makes `red` reusable.
template <typename T, int NV, typename Op>
__device__ __forceinline__ void block_reduce(T (&v)[NV], Op op, T ident, T* red, const Team& tm) {
  const int lane = tm.tid & 31, warp = tm.tid >> 5, nw = (tm.nt + 31) >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) v[q] = warp_reduce(v[q], op);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) red[q * 32 + warp] = v[q];
  }
  tm.sync();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    T t = (lane < nw) ? red[q * 32 + lane] : ident;
    v[q] = warp_reduce(t, op);
  }
  tm.sync();
}

template <typename T, int NV, typename Op>
__device__ __forceinline__ void block_reduce(T (&v)[NV], Op op, T ident, T* red) {
  block_reduce<T, NV, Op>(v, op, ident, red, Team::cta());
}

// ---------------------------------------------------------------- GEMV
// out[r] = epi(r, sum_j A[r*lda+j] * x[j]),  r in [0,M): one warp per row, coalesced along j.
template <typename T, typename Epi>
__device__ __forceinline__ void gemv_rows(const T* __restrict__ A, int lda, int M, int N,
                                          const T* x, Epi epi, const Team& tm) {
  const int lane = tm.tid & 31, warp = tm.tid >> 5, nw = tm.nt >> 5;
  for (int r = warp; r < M; r += nw) {
    const T* row = A + (size_t)r * lda;
    T acc = 0;
    for (int j = lane; j < N; j += 32) acc += row[j] * x[j];
    acc = warp_reduce(acc, OpSum());
    if (lane == 0) epi(r, acc);
  }
  tm.sync();
}

// out[j] = epi(j, sum_i A[i*lda+j] * w[i]),  j in [0,N): columns across lanes, row range
// split over thread groups, partials combined through `scratch` (>= blockDim.x elements).
template <typename T, typename Epi>
__device__ __forceinline__ void gemv_cols(const T* __restrict__ A, int lda, int M, int N,
                                          const T* w, T* scratch, Epi epi, const Team& tm) {
  const int NT = tm.nt;
  for (int j0 = 0; j0 < N; j0 += NT) {
    const int nj = min(N - j0, NT);
    const int njp = (nj + 31) & ~31;
    const int parts = NT / njp;                 // >= 1
    const int part = tm.tid / njp, jj = tm.tid - part * njp;
    T acc = 0;
    if (part < parts && jj < nj) {
      const T* col = A + j0 + jj;
      for (int i = part; i < M; i += parts) acc += col[(size_t)i * lda] * w[i];
    }
    scratch[tm.tid] = acc;
    tm.sync();
    if (tm.tid < nj) {
      T t = 0;
      for (int q = 0; q < parts; ++q) t += scratch[q * njp + tm.tid];
      epi(j0 + tm.tid, t);
    }
    tm.sync();
  }
}

// ---------------------------------------------------------------- small GEMMs
// C[M,N] (ldc) = beta*C + alpha * A[M,K] * op(B);  TRANSB ? B is [N,K] : B is [K,N].
// 4x4 register tiles, columns fastest across threads. All pointers generic.
template <typename T, bool TRANSB>
__device__ __forceinline__ void gemm_tiled(T* __restrict__ C, int ldc, const T* __restrict__ A, int lda,
                                           const T* __restrict__ Bm, int ldb, int M, int N, int K,
                                           T alpha, T beta) {
  const int tr = (M + 3) >> 2, tc = (N + 3) >> 2;
  for (int t = threadIdx.x; t < tr * tc; t += blockDim.x) {
    const int ti = t / tc, tj = t - ti * tc;
    const int i0 = ti * 4, j0 = tj * 4;
    T acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0;
    int ir[4], jc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { ir[r] = min(i0 + r, M - 1); jc[r] = min(j0 + r, N - 1); }
    for (int k = 0; k < K; ++k) {
      T a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = A[(size_t)ir[r] * lda + k];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = TRANSB ? Bm[(size_t)jc[c] * ldb + k] : Bm[(size_t)k * ldb + jc[c]];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] += a[r] * b[c];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (i0 + r < M && j0 + c < N) {
          T* dst = C + (size_t)(i0 + r) * ldc + j0 + c;
          *dst = (beta == T(0) ? T(0) : beta * *dst) + alpha * acc[r][c];
        }
  }
  __syncthreads();
}

// ---------------------------------------------------------------- in-place inverse (Gauss-Jordan, no pivoting)
// A[n,n] (lda) <- A^{-1}. Sets *flag = 1 on a zero / non-finite pivot. Used for Q (pdipm.py:362,
// the reference LU-factors Q; we keep Q^{-1} explicitly so every later Q-solve is a GEMV) and
// for the small e x e block A Q^{-1} A^T (pdipm.py:387).
template <typename T>
__device__ __forceinline__ void invert_inplace(T* A, int lda, int n, int* flag, T* colk /* n scratch (shared) */) {
  for (int k = 0; k < n; ++k) {
    const T piv = A[(size_t)k * lda + k];
    if (threadIdx.x == 0 && !(piv != T(0) && isfinite((double)piv))) *flag = 1;
    const T r = T(1) / piv;
    // save column k (multipliers) then rewrite it
    for (int i = threadIdx.x; i < n; i += blockDim.x) colk[i] = A[(size_t)i * lda + k];
    __syncthreads();
    // scale pivot row
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      T v = A[(size_t)k * lda + j];
      A[(size_t)k * lda + j] = (j == k) ? r : v * r;
    }
    __syncthreads();
    // eliminate column k from all other rows
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int i = t / n, j = t - i * n;
      if (i == k) continue;
      const T f = colk[i];
      const T pk = A[(size_t)k * lda + j];
      if (j == k) A[(size_t)i * lda + j] = -f * pk;            // pk == r here
      else A[(size_t)i * lda + j] -= f * pk;
    }
    __syncthreads();
  }
}

}  // namespace lcpb200
