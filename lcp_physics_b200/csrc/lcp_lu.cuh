// lcp_lu.cuh -- blocked LU factorisation / solves of the m x m Schur matrix T, one CTA per scene.
//
// The matrix is held through a TView: rows [0,m1) in a `main` array (all mp columns), rows
// [m1,mp) in a `low` array (columns [0,m1) only). When m1 == mp everything is in `main` (which may
// be shared memory, or an L2-resident workspace for very large problems). When the full square
// does not fit one CTA's shared memory (m = 256 fp32 is 256 KiB > 227 KiB) the planner picks
// m1 = mp/2: the first m1 pivots are eliminated on the L-shaped region that IS resident, the
// trailing block S22 = T22 - L21 U12 is accumulated in REGISTERS straight from the L2 copy of T22,
// U12 (final by then) is spilled to an L2 workspace, S22 takes its place in `main`, and the second
// half is factored there. Factors therefore live: L11\U11, L21, L22\U22 in shared memory, U12 in L2.
//
// mp is m padded to a multiple of NB with an identity block, so no partial blocks exist.
// Pivoting: threshold partial pivoting restricted to each NB x NB diagonal block (lcp_device.cuh).
#pragma once
#include "lcp_device.cuh"

namespace lcpb200 {

template <typename T> struct VecOf;
template <> struct VecOf<float> { using type = float4; static constexpr int VC = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int VC = 2; };

template <typename T> __device__ __forceinline__ void vec_get(const float4& v, T (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
template <typename T> __device__ __forceinline__ void vec_get(const double2& v, T (&o)[2]) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ float4 vec_make(const float (&o)[4]) { return make_float4(o[0], o[1], o[2], o[3]); }
__device__ __forceinline__ double2 vec_make(const double (&o)[2]) { return make_double2(o[0], o[1]); }

// All dynamic shared memory of the solver kernels. Declared at namespace scope so that every device
// function (also non-inlined ones) can form pointers the compiler PROVES are shared (LDS/STS, 32-bit
// addresses) instead of carrying generic 64-bit pointers through structs.
extern __shared__ __align__(16) unsigned char lcpb200_smem[];
template <typename T> __device__ __forceinline__ T* smem_base() { return reinterpret_cast<T*>(lcpb200_smem); }

// MODE 0: T fully in shared memory; 1: split (main + low in shared memory, U12 in L2); 2: T in L2.
template <typename T, int MODE>
struct TView {
  int main_off, low_off;  // shared-memory offsets in elements (MODE 0/1)
  T* main_g;              // L2 workspace (MODE 2)
  int ld, ldl;
  T* u12;                 // L2 spill of U12: [m1, mp-m1] row-major (MODE 1)
  int mp, m1;
  // main: rows [0,m1), columns [0,mp);  low: rows [m1,mp), columns [0,m1)
  __device__ __forceinline__ T* main() const { return MODE == 2 ? main_g : smem_base<T>() + main_off; }
  __device__ __forceinline__ T* low() const { return smem_base<T>() + low_off; }
};

// ---------------------------------------------------------------------------------------------
// C[rows r_lo..r_hi) x [c_lo..c_hi)  -=  L[rows, k0..k0+NB) * U[k0..k0+NB, cols]   (rank-NB update)
// Thread tile TR x 2VC; warp tile (4 TR) x (16 VC): lane = (g = lane>>3, cg = lane&7) owns rows
// rbase + g + 4 r and the two column vectors cbase + VC cg, cbase + 8 VC + VC cg, which makes the
// L loads (vectors along k, 4 consecutive rows per instruction) and the U loads (128 contiguous
// bytes per instruction, broadcast to the 4 row groups) bank-conflict-free for ld = 4 (mod 32) words.
// rowsrc: functor i -> pointer to row i of the matrix holding BOTH the L panel and C.
// urow:   functor k -> pointer to row k0+k of the matrix holding U (always `main`).
template <typename T, int TR, typename RowFn>
__device__ __forceinline__ void rank_nb_update(RowFn rowp, const T* __restrict__ Umain, int ldu, int k0,
                                               int r_lo, int r_hi, int c_lo, int c_hi) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, NB = Blk<T>::NB;
  constexpr int WR = 4 * TR, WC = 16 * VC;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int g = lane >> 3, cg = lane & 7;
  const int ntr = (r_hi - r_lo + WR - 1) / WR, ntc = (c_hi - c_lo + WC - 1) / WC;
  for (int wt = warp; wt < ntr * ntc; wt += nw) {
    const int tr_ = wt / ntc, tc_ = wt - tr_ * ntc;
    const int rbase = r_lo + tr_ * WR + g;
    const int c0 = c_lo + tc_ * WC + VC * cg, c1 = c0 + 8 * VC;
    const bool v0 = c0 < c_hi, v1 = c1 < c_hi;
    T* rp[TR];
    bool rv[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int i = rbase + 4 * r;
      rv[r] = i < r_hi;
      rp[r] = rowp(rv[r] ? i : r_hi - 1);
    }
    T acc[TR][2 * VC];
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int c = 0; c < 2 * VC; ++c) acc[r][c] = 0;
#pragma unroll 2
    for (int kc = 0; kc < NB; kc += VC) {
      T l[TR][VC];
#pragma unroll
      for (int r = 0; r < TR; ++r) vec_get<T>(*reinterpret_cast<const V*>(rp[r] + k0 + kc), l[r]);
#pragma unroll
      for (int kk = 0; kk < VC; ++kk) {
        const T* ur = Umain + (size_t)(k0 + kc + kk) * ldu;
        T u[2 * VC];
        {
          T lo[VC], hi[VC];
          if (v0) vec_get<T>(*reinterpret_cast<const V*>(ur + c0), lo); else { for (int q = 0; q < VC; ++q) lo[q] = 0; }
          if (v1) vec_get<T>(*reinterpret_cast<const V*>(ur + c1), hi); else { for (int q = 0; q < VC; ++q) hi[q] = 0; }
#pragma unroll
          for (int q = 0; q < VC; ++q) { u[q] = lo[q]; u[VC + q] = hi[q]; }
        }
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
          for (int c = 0; c < 2 * VC; ++c) acc[r][c] = fma(l[r][kk], u[c], acc[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      if (!rv[r]) continue;
      if (v0) {
        T cur[VC];
        vec_get<T>(*reinterpret_cast<const V*>(rp[r] + c0), cur);
#pragma unroll
        for (int q = 0; q < VC; ++q) cur[q] -= acc[r][q];
        *reinterpret_cast<V*>(rp[r] + c0) = vec_make(cur);
      }
      if (v1) {
        T cur[VC];
        vec_get<T>(*reinterpret_cast<const V*>(rp[r] + c1), cur);
#pragma unroll
        for (int q = 0; q < VC; ++q) cur[q] -= acc[r][VC + q];
        *reinterpret_cast<V*>(rp[r] + c1) = vec_make(cur);
      }
    }
  }
}

// Pick the thread-tile height so that the warp tiles roughly fill the CTA.
template <typename T, typename RowFn>
__device__ __forceinline__ void rank_nb_update_auto(RowFn rowp, const T* Umain, int ldu, int k0, int r_lo, int r_hi,
                                                    int c_lo, int c_hi) {
  constexpr int VC = VecOf<T>::VC;
  if (r_hi <= r_lo || c_hi <= c_lo) return;
  const int nw = blockDim.x >> 5;
  const int ntc = (c_hi - c_lo + 16 * VC - 1) / (16 * VC);
  const int rows = r_hi - r_lo;
  const int t8 = ((rows + 31) / 32) * ntc;
  const int t4 = ((rows + 15) / 16) * ntc;
  // rounds(t) * cost(tile): prefer the taller tile unless it leaves warps idle
  const int c8 = ((t8 + nw - 1) / nw) * 8, c4 = ((t4 + nw - 1) / nw) * 4;
  const int t2 = ((rows + 7) / 8) * ntc;
  const int c2 = ((t2 + nw - 1) / nw) * 2;
  if (c8 <= c4 && c8 <= c2) rank_nb_update<T, 8>(rowp, Umain, ldu, k0, r_lo, r_hi, c_lo, c_hi);
  else if (c4 <= c2) rank_nb_update<T, 4>(rowp, Umain, ldu, k0, r_lo, r_hi, c_lo, c_hi);
  else rank_nb_update<T, 2>(rowp, Umain, ldu, k0, r_lo, r_hi, c_lo, c_hi);
}

// ---------------------------------------------------------------------------------------------
// Diagonal block (NB x NB at D, leading dimension ld, in place), executed by ONE warp with lane i
// owning row i IN MEMORY (no per-thread register arrays, run-time loops: ~1 KB of code, so this
// single-warp dependent chain -- the critical path of the whole factorisation -- neither spills nor
// thrashes the instruction cache; the fully unrolled shuffle version it replaces took 66k cycles
// per block, 52% of the forward kernel, mostly in instruction-fetch and shuffle stalls).
//   1. P_b L U  by right-looking elimination, reciprocal pivot scaling (getf2), threshold pivoting
//      restricted to the rows of the block: the natural row is kept unless its pivot is below
//      tau * (largest entry the row had when the block was loaded); only then the largest |entry|
//      of the column among the remaining rows is swapped in (rows, scales and perm move together).
//      Measured (DESIGN.md "Pivoting"): eager swaps inside a block HURT fp32 trajectory parity,
//      never swapping leaves exact-zero pivots (0/0 -> NaN) on converged scenes.
//   2. both triangular inverses in place and interleaved: inv(L) by ascending, inv(U) by
//      descending right-looking substitution (row scaling of inv(U) deferred to the end).
// Result: D = [strict-lower(inv L) \ upper(inv U)], perm[i] = source row of row i.
// rmaxs: NB scratch elements (shared). Returns (warp-uniform) whether rows were interchanged.
template <typename T, int NB>
__device__ __forceinline__ bool diag_block_inplace(T* D, int ld, int* perm, T* rmaxs) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC;
  const int lane = threadIdx.x & 31;
  const bool act = lane < NB;
  T* row = D + (size_t)(act ? lane : 0) * ld;
  const T tau = (sizeof(T) == 4) ? T(1e-4) : T(1e-8);
  {
    T rm = 0;
    for (int j0 = 0; j0 < NB; j0 += VC) {
      T a[VC];
      vec_get<T>(*reinterpret_cast<const V*>(row + j0), a);
#pragma unroll
      for (int q = 0; q < VC; ++q) rm = fmax(rm, fabs(a[q]));
    }
    if (act) { rmaxs[lane] = rm; perm[lane] = lane; }
  }
  __syncwarp();
  bool moved = false;
  for (int k = 0; k < NB; ++k) {
    T piv = D[(size_t)k * ld + k];
    const T rs = rmaxs[k];
    if (!(fabs(piv) >= tau * rs && fabs(piv) > T(0))) {          // rare: partial pivoting inside the block
      T best = (act && lane >= k) ? fabs(row[k]) : T(-1);
      if (best != best) best = INFINITY;
      int bi = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const T ov = __shfl_xor_sync(FULL, best, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (bi != k) {
        if (act) {
          const T t0 = D[(size_t)k * ld + lane], t1 = D[(size_t)bi * ld + lane];
          D[(size_t)k * ld + lane] = t1;
          D[(size_t)bi * ld + lane] = t0;
        }
        if (lane == 0) {
          const T r0 = rmaxs[k]; rmaxs[k] = rmaxs[bi]; rmaxs[bi] = r0;
          const int p0 = perm[k]; perm[k] = perm[bi]; perm[bi] = p0;
        }
        moved = true;
        __syncwarp();
        piv = D[(size_t)k * ld + k];
      }
    }
    const T r = T(1) / piv;
    if (act && lane > k) {
      const T l = row[k] * r;
      const T* prow = D + (size_t)k * ld;
      for (int j0 = (k / VC) * VC; j0 < NB; j0 += VC) {
        T u[VC], a[VC];
        vec_get<T>(*reinterpret_cast<const V*>(prow + j0), u);
        vec_get<T>(*reinterpret_cast<const V*>(row + j0), a);
#pragma unroll
        for (int q = 0; q < VC; ++q) {
          if (j0 + q > k) a[q] = fma(-l, u[q], a[q]);
          else if (j0 + q == k) a[q] = l;
        }
        *reinterpret_cast<V*>(row + j0) = vec_make(a);
      }
    }
    __syncwarp();
  }
  // ---- inverses, in place
  for (int k = 0; k < NB; ++k) {
    const int kk = NB - 1 - k;
    if (act && lane > k) {                       // inv(L): rows below k
      const T mlt = row[k];
      const T* prow = D + (size_t)k * ld;
      for (int j0 = 0; j0 <= k; j0 += VC) {
        T x[VC], a[VC];
        vec_get<T>(*reinterpret_cast<const V*>(prow + j0), x);
        vec_get<T>(*reinterpret_cast<const V*>(row + j0), a);
#pragma unroll
        for (int q = 0; q < VC; ++q) {
          if (j0 + q < k) a[q] = fma(-mlt, x[q], a[q]);
          else if (j0 + q == k) a[q] = -mlt;
        }
        *reinterpret_cast<V*>(row + j0) = vec_make(a);
      }
    }
    if (act && lane < kk) {                      // inv(U) (unscaled rows Z): rows above kk
      const T* prow = D + (size_t)kk * ld;
      const T f = row[kk] * (T(1) / prow[kk]);
      for (int j0 = (kk / VC) * VC; j0 < NB; j0 += VC) {
        T z[VC], a[VC];
        vec_get<T>(*reinterpret_cast<const V*>(prow + j0), z);
        vec_get<T>(*reinterpret_cast<const V*>(row + j0), a);
#pragma unroll
        for (int q = 0; q < VC; ++q) {
          if (j0 + q > kk) a[q] = fma(-f, z[q], a[q]);
          else if (j0 + q == kk) a[q] = -f;
        }
        *reinterpret_cast<V*>(row + j0) = vec_make(a);
      }
    }
    __syncwarp();
  }
  if (act) {                                     // inv(U)[i][j] = Z[i][j] / U[i][i], inv(U)[i][i] = 1 / U[i][i]
    const T ri = T(1) / row[lane];
    for (int j0 = (lane / VC) * VC; j0 < NB; j0 += VC) {
      T a[VC];
      vec_get<T>(*reinterpret_cast<const V*>(row + j0), a);
#pragma unroll
      for (int q = 0; q < VC; ++q) {
        if (j0 + q > lane) a[q] *= ri;
        else if (j0 + q == lane) a[q] = ri;
      }
      *reinterpret_cast<V*>(row + j0) = vec_make(a);
    }
  }
  return moved;
}

// ---------------------------------------------------------------------------------------------
// Panel solves of one block step. D = diagonal block (already [inv L \ inv U], lcp_device.cuh).
//   U12 columns  [c_lo, c_hi) of main rows k0..k0+NB:   U12 = inv(L11) A12   (one column per thread)
//   L21 rows     [r_lo, r_hi):                           L21 = A21 inv(U11)   (one row per thread)
template <typename T, typename RowFn>
__device__ __forceinline__ void panel_solves(RowFn rowp, T* Umain, int ldu, int k0, int r_lo, int r_hi, int c_lo,
                                             int c_hi) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, NB = Blk<T>::NB;
  const T* D = Umain + (size_t)k0 * ldu + k0;
  const int ncol = max(c_hi - c_lo, 0), nrow = max(r_hi - r_lo, 0);
  for (int t = threadIdx.x; t < ncol + nrow; t += blockDim.x) {
    T a[NB], y[NB];
    if (t < ncol) {
      T* col = Umain + (size_t)k0 * ldu + c_lo + t;
#pragma unroll
      for (int r = 0; r < NB; ++r) a[r] = col[(size_t)r * ldu];
#pragma unroll
      for (int r = 0; r < NB; ++r) {
        T acc = a[r];
        // inv(L11)[r][0..r): vector loads along the row (broadcast across the warp)
#pragma unroll
        for (int q0 = 0; q0 < r; q0 += VC) {
          T lv[VC];
          vec_get<T>(*reinterpret_cast<const V*>(D + (size_t)r * ldu + q0), lv);
#pragma unroll
          for (int q = 0; q < VC; ++q)
            if (q0 + q < r) acc = fma(lv[q], a[q0 + q], acc);
        }
        y[r] = acc;
      }
#pragma unroll
      for (int r = 0; r < NB; ++r) col[(size_t)r * ldu] = y[r];
    } else {
      T* row = rowp(r_lo + t - ncol) + k0;
#pragma unroll
      for (int c0 = 0; c0 < NB; c0 += VC) {
        T tv[VC];
        vec_get<T>(*reinterpret_cast<const V*>(row + c0), tv);
#pragma unroll
        for (int q = 0; q < VC; ++q) { a[c0 + q] = tv[q]; y[c0 + q] = 0; }
      }
      // y[c] = sum_{r<=c} a[r] inv(U11)[r][c]: sweep rows of inv(U11), vector loads along the row
#pragma unroll
      for (int r = 0; r < NB; ++r) {
#pragma unroll
        for (int c0 = (r / VC) * VC; c0 < NB; c0 += VC) {
          T uv[VC];
          vec_get<T>(*reinterpret_cast<const V*>(D + (size_t)r * ldu + c0), uv);
#pragma unroll
          for (int q = 0; q < VC; ++q)
            if (c0 + q >= r) y[c0 + q] = fma(a[r], uv[q], y[c0 + q]);
        }
      }
#pragma unroll
      for (int c0 = 0; c0 < NB; c0 += VC) {
        T tv[VC];
#pragma unroll
        for (int q = 0; q < VC; ++q) tv[q] = y[c0 + q];
        *reinterpret_cast<V*>(row + c0) = vec_make(tv);
      }
    }
  }
}

// Apply the block's row interchanges to the columns outside the diagonal block.
template <typename T, typename RowFn>
__device__ __forceinline__ void apply_block_perm(RowFn rowp, int k0, const int* perm, int c_begin, int c_end) {
  constexpr int NB = Blk<T>::NB;
  // each thread owns one column: no barrier between its reads and writes
  for (int c = c_begin + threadIdx.x; c < c_end - NB; c += blockDim.x) {
    const int col = c < k0 ? c : c + NB;
    T tmp[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) tmp[r] = rowp(k0 + perm[k0 + r])[col];
#pragma unroll
    for (int r = 0; r < NB; ++r) rowp(k0 + r)[col] = tmp[r];
  }
}

// ---------------------------------------------------------------------------------------------
// Factor one square region held entirely in `main`-style storage: rows/cols [0, sz) of the matrix
// whose (0,0) is at A (ld), optionally with extra "low" rows [sz, sz+nlow) that only carry the L
// panel (columns [0, sz)) -- the split case's phase 1. perm/flag: shared.
// shadow (may be null): rows [0,sz) x [0,shadow_cols) of another array that must follow the row
// interchanges (the L21 rows of the split's second half).
template <typename T>
__device__ __forceinline__ void lu_region(T* A, int ld, int sz, int ncols_total, T* low, int ldl, int nlow, int* perm,
                                          int* flag, T* rmaxs, T* shadow = nullptr, int lds = 0, int shadow_cols = 0) {
  constexpr int NB = Blk<T>::NB;
  auto rowp = [=](int i) -> T* { return i < sz ? A + (size_t)i * ld : low + (size_t)(i - sz) * ldl; };
  for (int k0 = 0; k0 < sz; k0 += NB) {
    if (threadIdx.x < 32) {
      const bool moved = diag_block_inplace<T, NB>(A + (size_t)k0 * ld + k0, ld, perm + k0, rmaxs);
      if (threadIdx.x == 0) *flag = moved ? 1 : 0;
    }
    __syncthreads();
    if (*flag) {
      apply_block_perm<T>(rowp, k0, perm, 0, ncols_total);
      if (shadow) {
        for (int c = threadIdx.x; c < shadow_cols; c += blockDim.x) {
          T tmp[NB];
#pragma unroll
          for (int r = 0; r < NB; ++r) tmp[r] = shadow[(size_t)(k0 + perm[k0 + r]) * lds + c];
#pragma unroll
          for (int r = 0; r < NB; ++r) shadow[(size_t)(k0 + r) * lds + c] = tmp[r];
        }
      }
      __syncthreads();
    }
    const int r0 = k0 + NB;
    if (r0 >= ncols_total && r0 >= sz + nlow) break;
    panel_solves<T>(rowp, A, ld, k0, r0, sz + nlow, r0, ncols_total);
    __syncthreads();
    // trailing update: rows [r0, sz) x cols [r0, ncols_total)  and  low rows [sz, sz+nlow) x cols [r0, sz)
    rank_nb_update_auto<T>(rowp, A, ld, k0, r0, sz, r0, ncols_total);
    if (nlow > 0) rank_nb_update_auto<T>(rowp, A, ld, k0, sz, sz + nlow, r0, sz);
    __syncthreads();
  }
}

// Schur block of the split: S22 = T22 - L21 U12, accumulated in registers (K = m1), T22 = R22 + diag
// read from the L2 copy `R` (ld = ldr, already offset to (m1,m1)); U12 is spilled to v.u12 and S22
// written over it in main[0..n2) x [m1, mp).
template <typename T, int MODE>
__device__ __forceinline__ void split_schur(const TView<T, MODE>& v, const T* __restrict__ R22, int ldr,
                                            const T* dinv, int m_real) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC;
  constexpr int TR = 4, WR = 4 * TR, WC = 16 * VC;
  const int m1 = v.m1, n2 = v.mp - v.m1;
  T* const vmain = v.main();
  T* const vlow = v.low();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int g = lane >> 3, cg = lane & 7;
  const int ntr = (n2 + WR - 1) / WR, ntc = (n2 + WC - 1) / WC;
  // every warp owns at most MAXT tiles whose accumulators stay in registers across the spill
  constexpr int MAXT = 1;
  T acc[MAXT][TR][2 * VC];
  int nt = 0;
  for (int wt = warp; wt < ntr * ntc && nt < MAXT; wt += nw, ++nt) {
    const int tr_ = wt / ntc, tc_ = wt - tr_ * ntc;
    const int rbase = tr_ * WR + g, c0 = tc_ * WC + VC * cg, c1 = c0 + 8 * VC;
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int c = 0; c < 2 * VC; ++c) acc[nt][r][c] = 0;
    const T* lp[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) lp[r] = vlow + (size_t)min(rbase + 4 * r, n2 - 1) * v.ldl;
#pragma unroll 2
    for (int kc = 0; kc < m1; kc += VC) {
      T l[TR][VC], u[VC][2 * VC];
#pragma unroll
      for (int r = 0; r < TR; ++r) vec_get<T>(*reinterpret_cast<const V*>(lp[r] + kc), l[r]);
#pragma unroll
      for (int kk = 0; kk < VC; ++kk) {
        const T* ur = vmain + (size_t)(kc + kk) * v.ld + m1;
        T lo[VC], hi[VC];
        vec_get<T>(*reinterpret_cast<const V*>(ur + min(c0, n2 - VC)), lo);
        vec_get<T>(*reinterpret_cast<const V*>(ur + min(c1, n2 - VC)), hi);
#pragma unroll
        for (int q = 0; q < VC; ++q) { u[kk][q] = lo[q]; u[kk][VC + q] = hi[q]; }
      }
#pragma unroll
      for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int kk = 0; kk < VC; ++kk)
#pragma unroll
          for (int c = 0; c < 2 * VC; ++c) acc[nt][r][c] = fma(l[r][kk], u[kk][c], acc[nt][r][c]);
    }
  }
  __syncthreads();                       // all reads of U12 done
  // spill U12 (final) to L2: main rows [0,m1) x cols [m1,mp) -> u12 [m1, n2]
  for (int t = threadIdx.x; t < m1 * (n2 / VC); t += blockDim.x) {
    const int i = t / (n2 / VC), jv = t - i * (n2 / VC);
    *reinterpret_cast<V*>(v.u12 + (size_t)i * n2 + jv * VC) = *reinterpret_cast<const V*>(vmain + (size_t)i * v.ld + m1 + jv * VC);
  }
  __syncthreads();
  // S22 = T22 - acc -> main[0..n2) x [m1, mp)
  nt = 0;
  for (int wt = warp; wt < ntr * ntc && nt < MAXT; wt += nw, ++nt) {
    const int tr_ = wt / ntc, tc_ = wt - tr_ * ntc;
    const int rbase = tr_ * WR + g, c0 = tc_ * WC + VC * cg, c1 = c0 + 8 * VC;
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int i = rbase + 4 * r;
      if (i >= n2) continue;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int c = hh ? c1 : c0;
        if (c >= n2) continue;
        T out[VC];
#pragma unroll
        for (int q = 0; q < VC; ++q) {
          const int gi = m1 + i, gj = m1 + c + q;       // global indices in the padded matrix
          T t22 = (gi < m_real && gj < m_real) ? R22[(size_t)i * ldr + c + q] : T(0);
          if (gi == gj) t22 += (gi < m_real) ? dinv[gi] : T(1);
          out[q] = t22 - acc[nt][r][hh * VC + q];
        }
        *reinterpret_cast<V*>(vmain + (size_t)i * v.ld + m1 + c) = vec_make(out);
      }
    }
  }
  __syncthreads();
}

// Full factorisation of the view (T must already be loaded: main rows [0,m1) all columns, low rows
// [m1,mp) columns [0,m1)). R22/ldr/dinv/m_real are only used by the split.
template <typename T, int MODE>
__device__ __forceinline__ void lu_factor_view(const TView<T, MODE>& v, const T* R22, int ldr, const T* dinv,
                                               int m_real, int* perm, int* flag, T* rmaxs) {
  if (MODE != 1) {
    lu_region<T>(v.main(), v.ld, v.mp, v.mp, (T*)nullptr, 0, 0, perm, flag, rmaxs);
    return;
  }
  lu_region<T>(v.main(), v.ld, v.m1, v.mp, v.low(), v.ldl, v.mp - v.m1, perm, flag, rmaxs);
  split_schur<T, MODE>(v, R22, ldr, dinv, m_real);
  lu_region<T>(v.main() + v.m1, v.ld, v.mp - v.m1, v.mp - v.m1, (T*)nullptr, 0, 0, perm + v.m1, flag, rmaxs, v.low(),
               v.ldl, v.m1);
}

// ---------------------------------------------------------------------------------------------
// x[mp] (shared) <- T^{-1} x using the factors above. tmp: mp scratch (shared).
template <typename T, int MODE>
__device__ __forceinline__ void lu_solve_view(const TView<T, MODE>& v, const int* perm, T* x, T* tmp) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, NB = Blk<T>::NB;
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
  const int mp = v.mp, m1 = v.m1, n2 = mp - m1;
  const T* const vmain = v.main();
  const T* const vlow = (MODE == 1) ? v.low() : vmain;
  for (int i = tid; i < mp; i += NT) tmp[i] = x[(i / NB) * NB + perm[i]];
  __syncthreads();
  for (int i = tid; i < mp; i += NT) x[i] = tmp[i];
  __syncthreads();
  // block row pointer helpers: L/U entries of logical row i, columns of logical block k0
  auto lrow = [&](int i, int k0) -> const T* {      // L part (i > block)
    if (k0 < m1) return (i < m1 ? vmain + (size_t)i * v.ld : vlow + (size_t)(i - m1) * v.ldl) + k0;
    return vmain + (size_t)(i - m1) * v.ld + k0;   // S22 region: row i-m1, column m1 + (k0-m1)
  };
  auto dblk = [&](int k0) -> const T* {
    return k0 < m1 ? vmain + (size_t)k0 * v.ld + k0 : vmain + (size_t)(k0 - m1) * v.ld + k0;
  };
  // ---- L y = x (right-looking)
  for (int k0 = 0; k0 < mp; k0 += NB) {
    if (tid < 32) {
      T acc = 0;
      if (lane < NB) {
        acc = x[k0 + lane];
        const T* row = dblk(k0) + (size_t)lane * v.ld;
        for (int q = 0; q < lane; ++q) acc = fma(row[q], x[k0 + q], acc);
      }
      __syncwarp();
      if (lane < NB) x[k0 + lane] = acc;
    }
    __syncthreads();
    for (int i = k0 + NB + tid; i < mp; i += NT) {
      const T* row = lrow(i, k0);
      T acc = x[i];
#pragma unroll
      for (int c0 = 0; c0 < NB; c0 += VC) {
        T a[VC], y[VC];
        vec_get<T>(*reinterpret_cast<const V*>(row + c0), a);
        vec_get<T>(*reinterpret_cast<const V*>(x + k0 + c0), y);
#pragma unroll
        for (int q = 0; q < VC; ++q) acc = fma(-a[q], y[q], acc);
      }
      x[i] = acc;
    }
    __syncthreads();
  }
  // ---- U x = y: second-half blocks, then the U12 coupling (from L2), then first-half blocks
  for (int k0 = mp - NB; k0 >= 0; k0 -= NB) {
    if (tid < 32) {
      T acc = 0;
      if (lane < NB) {
        const T* row = dblk(k0) + (size_t)lane * v.ld;
        for (int c = lane; c < NB; ++c) acc = fma(row[c], x[k0 + c], acc);
      }
      __syncwarp();
      if (lane < NB) x[k0 + lane] = acc;
    }
    __syncthreads();
    // rows above inside the same triangular factor
    const int top = k0 < m1 ? 0 : m1;
    for (int i = top + tid; i < k0; i += NT) {
      const T* row = (k0 < m1 ? vmain + (size_t)i * v.ld : vmain + (size_t)(i - m1) * v.ld) + k0;
      T acc = x[i];
#pragma unroll
      for (int c0 = 0; c0 < NB; c0 += VC) {
        T a[VC], y[VC];
        vec_get<T>(*reinterpret_cast<const V*>(row + c0), a);
        vec_get<T>(*reinterpret_cast<const V*>(x + k0 + c0), y);
#pragma unroll
        for (int q = 0; q < VC; ++q) acc = fma(-a[q], y[q], acc);
      }
      x[i] = acc;
    }
    __syncthreads();
    if (k0 == m1 && m1 < mp) {
      // x1 -= U12 x2, U12 [m1, n2] in L2: one warp per row, coalesced
      const int warp = tid >> 5, nw = NT >> 5;
      for (int i = warp; i < m1; i += nw) {
        const T* row = v.u12 + (size_t)i * n2;
        T acc = 0;
        for (int j = lane * VC; j < n2; j += 32 * VC) {
          T a[VC], y[VC];
          vec_get<T>(*reinterpret_cast<const V*>(row + j), a);
          vec_get<T>(*reinterpret_cast<const V*>(x + m1 + j), y);
#pragma unroll
          for (int q = 0; q < VC; ++q) acc = fma(a[q], y[q], acc);
        }
        acc = warp_reduce(acc, OpSum());
        if (lane == 0) x[i] -= acc;
      }
      __syncthreads();
    }
  }
}

}  // namespace lcpb200
