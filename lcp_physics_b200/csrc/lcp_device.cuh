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

// Block-wide reduction of up to 4 values at once; result broadcast to all threads.
// red: shared scratch of >= 4*32 elements. Ends with a barrier that makes `red` reusable.
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

// ---------------------------------------------------------------- blocked pivot-free LU with inverted diagonal blocks
// One warp: factor the NB x NB block at D (ld) as P_b L U and overwrite it with
// [strict-lower(inv L) \ upper(inv U)] (unit diagonal of L implicit). Lane i holds physical row i
// during elimination; perm[i] = physical row chosen as i-th pivot. Returns (warp-uniform) whether
// the permutation is not the identity.
//
// Pivoting: the natural row is kept unless its pivot is below tau * (largest entry the row had when
// the block was loaded) -- a shuffle-free scale, so the common path costs two broadcasts per pivot;
// only then a full |max| search over the unused rows of the block runs (getf2 restricted to the
// block). Measured (DESIGN.md "Pivoting"): eager swaps inside a block HURT fp32 trajectory parity,
// never swapping leaves exact-zero pivots (0/0 -> NaN) on converged scenes.
// This routine is a single-warp dependent chain (the critical path of the whole factorisation):
// the two triangular inverses are interleaved in one loop so their chains overlap.
template <typename T, int NB>
__device__ __forceinline__ bool diag_block_factor_invert(T* D, int ld, int kb, int* perm) {
  (void)kb;
  const int lane = threadIdx.x & 31;
  const int li = lane < NB ? lane : NB - 1;      // lanes >= NB mirror the last row and never store
  T a[NB], x[NB];
  T rmax = 0;
#pragma unroll
  for (int j = 0; j < NB; ++j) { a[j] = D[(size_t)li * ld + j]; rmax = fmax(rmax, fabs(a[j])); }
  const T tau = (sizeof(T) == 4) ? T(1e-4) : T(1e-8);
  unsigned alive = (NB == 32) ? FULL : ((1u << NB) - 1u);
  bool done = lane >= NB;
  int myord = lane;
  bool moved = false;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int nat = __ffs(alive) - 1;            // natural pivot row (uniform)
    int pl = nat;
    T ukk = __shfl_sync(FULL, a[k], nat);
    const T rs = __shfl_sync(FULL, rmax, nat);
    if (!(fabs(ukk) >= tau * rs && fabs(ukk) > T(0))) {     // rare: search the unused rows
      T best = done ? T(-1) : fabs(a[k]);
      if (best != best) best = INFINITY;         // NaN: take it, everything is NaN anyway
      int bi = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const T ov = __shfl_xor_sync(FULL, best, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      pl = bi;
      ukk = __shfl_sync(FULL, a[k], pl);
    }
    alive &= ~(1u << pl);
    if (lane == k) myord = pl;
    moved |= (pl != k);
    if (lane == pl) done = true;
    const T r = T(1) / ukk;
    const bool upd = !done;
    const T l = a[k] * r;
    if (upd) a[k] = l;
#pragma unroll
    for (int j = k + 1; j < NB; ++j) {
      const T ukj = __shfl_sync(FULL, a[j], pl);
      if (upd) a[j] = fma(-l, ukj, a[j]);
    }
  }
  if (moved) {                                   // lane i <- row of its pivot lane
#pragma unroll
    for (int j = 0; j < NB; ++j) a[j] = __shfl_sync(FULL, a[j], myord);
  }
  if (lane < NB) perm[lane] = myord;
  // reciprocal of the own diagonal entry of U, off the dependent chain
  T dg = T(1);
#pragma unroll
  for (int j = 0; j < NB; ++j)
    if (j == lane) dg = a[j];
  const T rown = T(1) / dg;
#pragma unroll
  for (int j = 0; j < NB; ++j) x[j] = (j == lane) ? T(1) : T(0);
  // inv(L) by row-oriented forward substitution (ascending k) and inv(U) by row-oriented back
  // substitution (descending kk), interleaved; they touch disjoint halves of x in every lane.
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    {
      const bool below = lane > k;
      const T lik = a[k];
#pragma unroll
      for (int j = 0; j < k; ++j) {
        const T xkj = __shfl_sync(FULL, x[j], k);
        if (below) x[j] = fma(-lik, xkj, x[j]);
      }
      if (below) x[k] -= lik;
    }
    {
      const int kk = NB - 1 - k;
      const T rk = __shfl_sync(FULL, rown, kk);
      const bool above = lane < kk;
      const T uik = a[kk];
#pragma unroll
      for (int j = kk; j < NB; ++j) {
        if (lane == kk) x[j] *= rk;
        const T xkj = __shfl_sync(FULL, x[j], kk);
        if (above) x[j] = fma(-uik, xkj, x[j]);
      }
    }
  }
  // the unit diagonal of inv(L) is implicit: x[lane] currently holds inv(U)[lane][lane]
  if (lane < NB) {
#pragma unroll
    for (int j = 0; j < NB; ++j) D[(size_t)lane * ld + j] = x[j];
  }
  return moved;
}

// A[m,m] (lda) <- blocked LU (pdipm.py:431), P A = L U with P block-diagonal: partial pivoting
// is restricted to the NB rows of the current diagonal block (the reference pivots over the
// whole column on CPU tensors and not at all on CUDA tensors, pdipm.py:18; block-local pivoting
// removes the exact-zero pivots of the pivot-free variant at one warp's cost). perm[i] = source
// row of row i (shared, m ints; `flag` one shared int). Diagonal NB x NB blocks are left
// INVERTED so that the later vector solves are block GEMVs instead of dependent chains.
template <typename T>
__device__ __forceinline__ void lu_blocked(T* A, int lda, int m, int* perm, int* flag) {
  constexpr int NB = Blk<T>::NB;
  const int tid = threadIdx.x, NT = blockDim.x;
  for (int k0 = 0; k0 < m; k0 += NB) {
    const int kb = min(NB, m - k0);
    T* D = A + (size_t)k0 * lda + k0;
    if (tid < 32) {
      const bool moved = diag_block_factor_invert<T, NB>(D, lda, kb, perm + k0);
      if (tid == 0) *flag = moved ? 1 : 0;
    }
    __syncthreads();
    if (*flag) {
      // apply the block's row interchanges to the columns left and right of the diagonal block;
      // each thread owns one column, so no barrier is needed between its reads and writes
      for (int c = tid; c < m - kb; c += NT) {
        const int col = c < k0 ? c : c + kb;
        T tmp[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r)
          if (r < kb) tmp[r] = A[(size_t)(k0 + perm[k0 + r]) * lda + col];
#pragma unroll
        for (int r = 0; r < NB; ++r)
          if (r < kb) A[(size_t)(k0 + r) * lda + col] = tmp[r];
      }
      __syncthreads();
    }
    const int r0 = k0 + kb, rem = m - r0;
    if (rem <= 0) break;
    // panels: U12 = inv(L11) A12 (one column per task), L21 = A21 inv(U11) (one row per task)
    for (int t = tid; t < 2 * rem; t += NT) {
      T a[NB], y[NB];
      if (t < rem) {
        T* col = A + (size_t)k0 * lda + r0 + t;
#pragma unroll
        for (int r = 0; r < NB; ++r) a[r] = (r < kb) ? col[(size_t)r * lda] : T(0);
#pragma unroll
        for (int r = 0; r < NB; ++r) {
          T acc = a[r];
#pragma unroll
          for (int q = 0; q < r; ++q)
            if (r < kb) acc += D[(size_t)r * lda + q] * a[q];
          y[r] = acc;
        }
#pragma unroll
        for (int r = 0; r < NB; ++r)
          if (r < kb) col[(size_t)r * lda] = y[r];
      } else {
        T* row = A + (size_t)(r0 + t - rem) * lda + k0;
#pragma unroll
        for (int c = 0; c < NB; ++c) a[c] = (c < kb) ? row[c] : T(0);
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          T acc = 0;
#pragma unroll
          for (int r = 0; r <= c; ++r)
            if (c < kb) acc += a[r] * D[(size_t)r * lda + c];
          y[c] = acc;
        }
#pragma unroll
        for (int c = 0; c < NB; ++c)
          if (c < kb) row[c] = y[c];
      }
    }
    __syncthreads();
    // trailing update A22 -= L21 U12, 4x4 register tiles
    const int tt = (rem + 3) >> 2;
    for (int t = tid; t < tt * tt; t += NT) {
      const int ti = t / tt, tj = t - ti * tt;
      const int i0 = r0 + 4 * ti, j0 = r0 + 4 * tj;
      int ir[4], jc[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { ir[r] = min(i0 + r, m - 1); jc[r] = min(j0 + r, m - 1); }
      T acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0;
      for (int k = 0; k < kb; ++k) {
        T l[4], u[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) l[r] = A[(size_t)ir[r] * lda + k0 + k];
#pragma unroll
        for (int c = 0; c < 4; ++c) u[c] = A[(size_t)(k0 + k) * lda + jc[c]];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] += l[r] * u[c];
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (i0 + r < m && j0 + c < m) A[(size_t)(i0 + r) * lda + j0 + c] -= acc[r][c];
    }
    __syncthreads();
  }
}

// v[m] (shared) <- (LU)^{-1} v using the factors left by lu_blocked. Right-looking block
// substitution: diagonal step = multiply by the stored block inverse (one warp), update step
// = one thread per remaining row.
template <typename T>
__device__ __forceinline__ void lu_solve_vec(const T* __restrict__ A, int lda, int m, const int* perm, T* v,
                                             T* tmp /* m scratch (shared) */) {
  constexpr int NB = Blk<T>::NB;
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
  // v <- P v (block-local row interchanges)
  for (int i = tid; i < m; i += NT) tmp[i] = v[(i / NB) * NB + perm[i]];
  __syncthreads();
  for (int i = tid; i < m; i += NT) v[i] = tmp[i];
  __syncthreads();
  // L y = v
  for (int k0 = 0; k0 < m; k0 += NB) {
    const int kb = min(NB, m - k0);
    if (tid < 32) {
      T acc = 0;
      if (lane < kb) {
        acc = v[k0 + lane];
        const T* row = A + (size_t)(k0 + lane) * lda + k0;
        for (int q = 0; q < lane; ++q) acc += row[q] * v[k0 + q];
      }
      __syncwarp();
      if (lane < kb) v[k0 + lane] = acc;
    }
    __syncthreads();
    for (int i = k0 + kb + tid; i < m; i += NT) {
      const T* row = A + (size_t)i * lda + k0;
      T acc = v[i];
      for (int c = 0; c < kb; ++c) acc -= row[c] * v[k0 + c];
      v[i] = acc;
    }
    __syncthreads();
  }
  // U x = y
  const int last = ((m - 1) / NB) * NB;
  for (int k0 = last; k0 >= 0; k0 -= NB) {
    const int kb = min(NB, m - k0);
    if (tid < 32) {
      T acc = 0;
      if (lane < kb) {
        const T* row = A + (size_t)(k0 + lane) * lda + k0;
        for (int c = lane; c < kb; ++c) acc += row[c] * v[k0 + c];
      }
      __syncwarp();
      if (lane < kb) v[k0 + lane] = acc;
    }
    __syncthreads();
    for (int i = tid; i < k0; i += NT) {
      const T* row = A + (size_t)i * lda + k0;
      T acc = v[i];
      for (int c = 0; c < kb; ++c) acc -= row[c] * v[k0 + c];
      v[i] = acc;
    }
    __syncthreads();
  }
}

}  // namespace lcpb200
