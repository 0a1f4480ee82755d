// lcp_lu.cuh -- blocked LU factorisation / solves of the m x m Schur matrix T, one CTA per scene.
//
// Storage (TView): rows [0,m1) in a `main` array (all mp columns), rows [m1,mp) in a `low` array
// (columns [0,m1) only). MODE 0: m1 == mp, everything in shared memory. MODE 1 (split): the full
// square does not fit one CTA's shared memory (m = 256 fp32 is 256 KiB > 227 KiB): m1 = mp/2, the
// first m1 pivots are eliminated on the L-shaped region that IS resident, the trailing block
// S22 = T22 - L21 U12 is accumulated in REGISTERS straight from the L2 copy of T22, U12 (final by
// then) is spilled to an L2 workspace, S22 takes its place in `main`, and the second half is factored
// there. Factors then live: L11\U11, L21, L22\U22 in shared memory, U12 in L2. MODE 2: `main` is an
// L2-resident workspace (problems too large for either).
//
// mp is m padded to a multiple of NB with an identity block, so no partial blocks exist.
// Pivoting: threshold partial pivoting restricted to each NB x NB diagonal block.
//
// Register discipline: with ~220 KB of shared memory carved out, L1 is a few KB, so a spilled
// register costs an L2 round trip. Every heavy routine is therefore its own __noinline__ function
// (own register allocation); matrices are passed as MPtr (shared-memory OFFSET, or a global pointer
// in MODE 2) and vectors as offsets, so that inside each function the compiler can still prove the
// accesses are to shared memory (LDS/STS with 32-bit addresses).
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

// All dynamic shared memory of the solver kernels, at namespace scope so every device function can
// form provably-shared pointers.
extern __shared__ __align__(16) unsigned char lcpb200_smem[];
template <typename T> __device__ __forceinline__ T* smem_base() { return reinterpret_cast<T*>(lcpb200_smem); }
__device__ __forceinline__ int* smem_int(int o_i) { return reinterpret_cast<int*>(lcpb200_smem) + o_i; }

// Matrix base: shared-memory element offset (MODE 0/1) or global pointer (MODE 2).
template <typename T, int MODE>
struct MPtr {
  long long off;          // elements relative to smem_base<T>() (may be negative after re-basing)
  T* g;
  __device__ __forceinline__ T* get() const { return MODE == 2 ? g : smem_base<T>() + off; }
  __device__ __forceinline__ MPtr plus(long long d) const { MPtr r; r.off = off + d; r.g = g + d; return r; }
};

template <typename T, int MODE>
struct TView {
  MPtr<T, MODE> main_;    // rows [0,m1), columns [0,mp)
  MPtr<T, 0> low_;        // rows [m1,mp), columns [0,m1)   (MODE 1 only, always shared)
  int ld, ldl;
  T* u12;                 // L2 spill of U12: [m1, mp-m1] row-major (MODE 1)
  int mp, m1;
  __device__ __forceinline__ T* main() const { return main_.get(); }
  __device__ __forceinline__ T* low() const { return low_.get(); }
};

// Shared vectors used by the factorisation. *_i offsets are in 4-byte ints from the start of the
// shared array, the others in elements of T.
struct LuVec {
  int o_perm_i, o_flag_i, o_rmaxs, o_rdiag, o_stage, o_lt, ldlt, lds;
};

// ---------------------------------------------------------------------------------------------
// C[rows r_lo..r_hi) x [c_lo..c_hi)  -=  L[rows, k0..k0+NB) * U[k0..k0+NB, cols]   (rank-NB update)
// C and the L panel live in the array `Cb` (row i at Cb + i*ldc), U in `Ub` (row k at Ub + k*ldu).
// Thread tile TR x 2VC; warp tile (4 TR) x (16 VC): lane = (g = lane>>3, cg = lane&7) owns rows
// rbase + g + 4 r and the two column vectors cbase + VC cg, cbase + 8 VC + VC cg, which makes the
// L loads (vectors along k, 4 consecutive rows per instruction) and the U loads (128 contiguous
// bytes per instruction, broadcast to the 4 row groups) bank-conflict-free for ld = 4 (mod 32) words.
template <typename T, int CMODE, int UMODE, int TR>
__device__ __noinline__ void rank_nb_update(MPtr<T, CMODE> Cb, int ldc, MPtr<T, UMODE> Ub, int ldu, int k0, int r_lo,
                                            int r_hi, int c_lo, int c_hi, int wid, int nwk, int wshift) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, NB = Blk<T>::NB;
  constexpr int WR = 4 * TR, WC = 16 * VC;
  T* const C = Cb.get();
  const T* const U = Ub.get();
  const int lane = threadIdx.x & 31;
  const int g = lane >> 3, cg = lane & 7;
  const int ntr = (r_hi - r_lo + WR - 1) / WR, ntc = (c_hi - c_lo + WC - 1) / WC;
  // tiles are dealt to the nwk participating warps starting at warp (wshift mod nwk), so that
  // back-to-back calls without a barrier in between continue where the previous one stopped
  int w0 = wid - wshift % nwk;
  if (w0 < 0) w0 += nwk;
  for (int wt = w0; wt < ntr * ntc; wt += nwk) {
    const int tr_ = wt / ntc, tc_ = wt - tr_ * ntc;
    const int rbase = r_lo + tr_ * WR + g;
    const int c0 = c_lo + tc_ * WC + VC * cg, c1 = c0 + 8 * VC;
    const bool v0 = c0 < c_hi, v1 = c1 < c_hi;
    // rows beyond r_hi are clamped (their results are not stored)
    T* const rp0 = C + (size_t)min(rbase, r_hi - 1) * ldc;
    int roff[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) roff[r] = (min(rbase + 4 * r, r_hi - 1) - min(rbase, r_hi - 1)) * ldc;
    T acc[TR][2 * VC];
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int c = 0; c < 2 * VC; ++c) acc[r][c] = 0;
#pragma unroll 2
    for (int kc = 0; kc < NB; kc += VC) {
      T l[TR][VC];
#pragma unroll
      for (int r = 0; r < TR; ++r) vec_get<T>(*reinterpret_cast<const V*>(rp0 + roff[r] + k0 + kc), l[r]);
#pragma unroll
      for (int kk = 0; kk < VC; ++kk) {
        const T* ur = U + (size_t)(k0 + kc + kk) * ldu;
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
      if (rbase + 4 * r >= r_hi) continue;
      T* rp = rp0 + roff[r];
      if (v0) {
        T cur[VC];
        vec_get<T>(*reinterpret_cast<const V*>(rp + c0), cur);
#pragma unroll
        for (int q = 0; q < VC; ++q) cur[q] -= acc[r][q];
        *reinterpret_cast<V*>(rp + c0) = vec_make(cur);
      }
      if (v1) {
        T cur[VC];
        vec_get<T>(*reinterpret_cast<const V*>(rp + c1), cur);
#pragma unroll
        for (int q = 0; q < VC; ++q) cur[q] -= acc[r][VC + q];
        *reinterpret_cast<V*>(rp + c1) = vec_make(cur);
      }
    }
  }
}

// Pick the thread-tile height so that the warp tiles roughly fill the participating warps
// (wid in [0, nwk); a warp with wid < 0 does not take part). Returns the number of tiles dealt,
// the caller adds it to wshift of its next call.
template <typename T, int CMODE, int UMODE>
__device__ __forceinline__ int rank_nb_update_auto(MPtr<T, CMODE> Cb, int ldc, MPtr<T, UMODE> Ub, int ldu, int k0,
                                                   int r_lo, int r_hi, int c_lo, int c_hi, int wid, int nwk,
                                                   int wshift) {
  constexpr int VC = VecOf<T>::VC;
  if (r_hi <= r_lo || c_hi <= c_lo) return 0;
  const int ntc = (c_hi - c_lo + 16 * VC - 1) / (16 * VC);
  const int rows = r_hi - r_lo;
  const int t8 = ((rows + 31) / 32) * ntc, t4 = ((rows + 15) / 16) * ntc, t2 = ((rows + 7) / 8) * ntc;
  // rounds(t) * cost(tile); the constant models the per-tile load/store overhead
  const int c8 = ((t8 + nwk - 1) / nwk) * 17, c4 = ((t4 + nwk - 1) / nwk) * 9, c2 = ((t2 + nwk - 1) / nwk) * 5;
  if (c8 <= c4 && c8 <= c2) {
    if (wid >= 0) rank_nb_update<T, CMODE, UMODE, 8>(Cb, ldc, Ub, ldu, k0, r_lo, r_hi, c_lo, c_hi, wid, nwk, wshift);
    return t8;
  } else if (c4 <= c2) {
    if (wid >= 0) rank_nb_update<T, CMODE, UMODE, 4>(Cb, ldc, Ub, ldu, k0, r_lo, r_hi, c_lo, c_hi, wid, nwk, wshift);
    return t4;
  }
  if (wid >= 0) rank_nb_update<T, CMODE, UMODE, 2>(Cb, ldc, Ub, ldu, k0, r_lo, r_hi, c_lo, c_hi, wid, nwk, wshift);
  return t2;
}

// Reciprocal off the slow IEEE-division path: hardware approximation + Newton steps (<= 1 ulp for
// float, ~1 ulp for double); falls back to a true division outside the seed's range.
__device__ __forceinline__ float fast_rcp(float x) {
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(x));
  const float r = fmaf(r0, fmaf(-x, r0, 1.0f), r0);
  return (fabsf(r0) < INFINITY) ? r : r0;        // x == 0 / denormal: keep the infinity (no NaN)
}
__device__ __forceinline__ double fast_rcp(double x) {
  float xf = (float)x, rf;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(xf));
  if (!(fabsf(rf) < INFINITY) || rf == 0.0f) return 1.0 / x;     // outside the float range: exact path
  double r = (double)rf;
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}

// Packed FP32 FMA (sm_100 FFMA2): d{0,1} = a{0,1} * b{0,1} + d{0,1} in ONE instruction. The pairs
// must sit in even-aligned adjacent registers (ptxas then emits no moves). Used only in the
// single-warp chain routines, which are bound by instruction issue (~1 instruction / 3 cycles).
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long d, a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
// out{0,1} = a{0,1} * b{0,1} + c{0,1}
__device__ __forceinline__ void ffma2_to(float& o0, float& o1, float a0, float a1, float b0, float b1, float c0, float c1) {
  unsigned long long d, a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(d));
}

// ---------------------------------------------------------------------------------------------
// Diagonal block (ONE warp -- the critical path of the factorisation): P_b L U of the NB x NB block
// with rows in REGISTERS and a ROLLED pivot loop. After eliminating a column every live row shifts
// its registers left by one as part of the FMA that updates them (a[j-1] = a[j] - l u[j]), so the
// current column is always a[0]: no dynamic register index, no unrolled pivot loop. The loop body is
// ~1.5 KB of code: a single warp running a 22 KB fully unrolled variant spent 25-30k cycles per
// block waiting on instruction fetch (ncu: no_instruction / short_scoreboard), this one ~5k.
// Pivoting: threshold partial pivoting inside the block -- the natural row is kept unless its pivot
// is below tau * (initial row scale) AND below tau * (largest candidate of the column in the
// block); only then rows are interchanged (registers via shuffles, the L part in memory).
// Measured (DESIGN.md "Pivoting"): eager swaps HURT fp32 trajectory parity, never swapping leaves
// exact-zero pivots on converged scenes.
// Outputs: D = L\U in place, rdiag[j] = 1/U[j][j], perm[i] = source row of row i, *flag = moved.
// Steps [k_begin, k_end) of the rolled factorisation with W live columns: at step k only NB - k <= W
// leading registers of a row are live, so the row broadcast and the row update touch W entries only
// (the caller runs four phases W = NB, 3NB/4, NB/2, NB/4: a single warp issues ~1 instruction per
// 3 cycles, instruction count is what the chain costs).
template <typename T, int NB, int W>
__device__ __forceinline__ void diag_rot_steps(int k_begin, int k_end, T (&a)[NB], T (&u)[NB], T& rinv, T& rm, T& myr,
                                               bool& moved, T* D, int ld, T* stage, int* perm, T* Lt, int ldlt) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, WV = W / VC, SL = NB + VC;
  const int lane = threadIdx.x & 31;
  const float tau = (sizeof(T) == 4) ? 1e-4f : 1e-8f;
#pragma unroll 1
  for (int k = k_begin; k < k_end; ++k) {
    T* const st = stage + (k & 1) * SL;
    {
      // the pivot row is broadcast through shared memory: one predicated vector store per chunk by
      // the pivot lane (a branch around the group costs a divergence round trip), broadcast loads by all
      const bool mine = lane == k;
#pragma unroll
      for (int c = 0; c < WV; ++c) {
        T t[VC];
#pragma unroll
        for (int q = 0; q < VC; ++q) t[q] = a[c * VC + q];
        if (mine) *reinterpret_cast<V*>(st + c * VC) = vec_make(t);
      }
      if (mine) st[NB] = rinv;
      if (mine) st[NB + 1] = rm;
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < WV; ++c) {
      T t[VC];
      vec_get<T>(*reinterpret_cast<const V*>(st + c * VC), t);
#pragma unroll
      for (int q = 0; q < VC; ++q) u[c * VC + q] = t[q];
    }
    T piv = u[0];
    T r = st[NB];
    const T rs = st[NB + 1];
    // Threshold partial pivoting, three tiers: (1) pivot >= tau * row scale: accept (the common
    // case, nothing extra on the chain); (2) else compare with the pivot column's largest candidate,
    // one REDUX on the float bit patterns (non-negative floats order like unsigned integers);
    // (3) only if that fails too search the arg-max row and interchange.
    bool suspect = !(fabs(piv) >= T(tau) * rs && fabs(piv) > T(0));
    if (suspect) {
      float f = fabsf((float)a[0]);
      if (f != f) f = INFINITY;
      const unsigned cb = __reduce_max_sync(FULL, (lane >= k && lane < NB) ? __float_as_uint(f) : 0u);
      suspect = !(fabsf((float)piv) >= 2.f * tau * __uint_as_float(cb) && fabs(piv) > T(0));
    }
    if (suspect) {                                               // rare, warp-uniform
      T best = (lane >= k && lane < NB) ? fabs(a[0]) : T(-1);
      if (best != best) best = INFINITY;
      int bi = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const T ov = __shfl_xor_sync(FULL, best, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (bi != k && !(fabs(piv) >= T(tau) * best && fabs(piv) > T(0))) {
        // interchange rows k and bi: live registers (same shift on both), scale, L part (both
        // copies), perm
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const T fk = __shfl_sync(FULL, a[j], k), fb = __shfl_sync(FULL, a[j], bi);
          if (lane == k) a[j] = fb; else if (lane == bi) a[j] = fk;
          u[j] = fb;                                               // the new pivot row
        }
        {
          const T fk = __shfl_sync(FULL, rm, k), fb = __shfl_sync(FULL, rm, bi);
          if (lane == k) rm = fb; else if (lane == bi) rm = fk;
        }
        if (lane < k) {
          const T t0 = D[(size_t)k * ld + lane], t1 = D[(size_t)bi * ld + lane];
          D[(size_t)k * ld + lane] = t1;
          D[(size_t)bi * ld + lane] = t0;
          Lt[(size_t)lane * ldlt + k] = t1;
          Lt[(size_t)lane * ldlt + bi] = t0;
        }
        if (lane == 0) { const int p0 = perm[k]; perm[k] = perm[bi]; perm[bi] = p0; }
        moved = true;
        __syncwarp();
        piv = u[0];
        r = fast_rcp(piv);
      }
    }
    if (lane == k) myr = r;
    const bool alive = lane > k && lane < NB;
    const T l = a[0] * r;
    if (alive) D[(size_t)lane * ld + k] = l;                     // multiplier: final entry of L
    if (alive) Lt[(size_t)k * ldlt + lane] = l;                  // ... and of the transposed copy
#pragma unroll
    for (int j = 1; j < W; ++j)
      if (alive) a[j - 1] = fma(-l, u[j], a[j]);
    rinv = fast_rcp(a[0]);
  }
}

// fp32 variant with packed FMAs: two pivots per loop iteration so that every FFMA2 works on
// even-aligned register pairs. Pivot k (OFF = 0) sits in a[0]: rows are updated in place (the dead
// a[0] rides along in the first pair). Pivot k+1 (OFF = 1) sits in a[1]: the update writes two
// registers lower, a[j-2] = a[j] - l u[j], which is the shift for both pivots at once. A row that was
// pivot at an odd step therefore ends with its U row starting at a[1] (see the final store).
template <int NB, int W, int OFF>
__device__ __forceinline__ void diag_one_step2(int k, float (&a)[NB], float (&u)[NB], float& rinv, float& rm, float& myr,
                                               bool& moved, float* D, int ld, float* stage, int* perm, float* Lt, int ldlt) {
  constexpr int VC = 4, WV = W / VC, SL = NB + VC;
  const int lane = threadIdx.x & 31;
  const float tau = 1e-4f;
  float* const st = stage + (k & 1) * SL;
  {
    const bool mine = lane == k;
#pragma unroll
    for (int c = 0; c < WV; ++c)
      if (mine) *reinterpret_cast<float4*>(st + c * VC) = make_float4(a[c * VC], a[c * VC + 1], a[c * VC + 2], a[c * VC + 3]);
    if (mine) st[NB] = rinv;
    if (mine) st[NB + 1] = rm;
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < WV; ++c) {
    const float4 t = *reinterpret_cast<const float4*>(st + c * VC);
    u[c * VC] = t.x; u[c * VC + 1] = t.y; u[c * VC + 2] = t.z; u[c * VC + 3] = t.w;
  }
  float piv = u[OFF];
  float r = st[NB];
  const float rs = st[NB + 1];
  bool suspect = !(fabsf(piv) >= tau * rs && fabsf(piv) > 0.f);   // tier 1 (see diag_rot_steps)
  if (suspect) {                                                   // tier 2: one REDUX over the column
    float f = fabsf(a[OFF]);
    if (f != f) f = INFINITY;
    const unsigned cb = __reduce_max_sync(FULL, (lane >= k && lane < NB) ? __float_as_uint(f) : 0u);
    suspect = !(fabsf(piv) >= 2.f * tau * __uint_as_float(cb) && fabsf(piv) > 0.f);
  }
  if (suspect) {                                                   // tier 3: arg-max search, interchange
    float best = (lane >= k && lane < NB) ? fabsf(a[OFF]) : -1.f;
    if (best != best) best = INFINITY;
    int bi = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(FULL, best, o);
      const int oi = __shfl_xor_sync(FULL, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (bi != k && !(fabsf(piv) >= tau * best && fabsf(piv) > 0.f)) {
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const float fk = __shfl_sync(FULL, a[j], k), fb = __shfl_sync(FULL, a[j], bi);
        if (lane == k) a[j] = fb; else if (lane == bi) a[j] = fk;
        u[j] = fb;
      }
      {
        const float fk = __shfl_sync(FULL, rm, k), fb = __shfl_sync(FULL, rm, bi);
        if (lane == k) rm = fb; else if (lane == bi) rm = fk;
      }
      if (lane < k) {
        const float t0 = D[(size_t)k * ld + lane], t1 = D[(size_t)bi * ld + lane];
        D[(size_t)k * ld + lane] = t1;
        D[(size_t)bi * ld + lane] = t0;
        Lt[(size_t)lane * ldlt + k] = t1;
        Lt[(size_t)lane * ldlt + bi] = t0;
      }
      if (lane == 0) { const int p0 = perm[k]; perm[k] = perm[bi]; perm[bi] = p0; }
      moved = true;
      __syncwarp();
      piv = u[OFF];
      r = fast_rcp(piv);
    }
  }
  if (lane == k) myr = r;
  const bool alive = lane > k && lane < NB;
  const float l = a[OFF] * r;
  if (alive) D[(size_t)lane * ld + k] = l;
  if (alive) Lt[(size_t)k * ldlt + lane] = l;
  const float nl = -l;
  if (OFF == 0) {
#pragma unroll
    for (int i = 0; i < W / 2; ++i)
      if (alive) ffma2(a[2 * i], a[2 * i + 1], nl, nl, u[2 * i], u[2 * i + 1]);
    rinv = fast_rcp(a[1]);
  } else {
#pragma unroll
    for (int i = 1; i < W / 2; ++i)
      if (alive) ffma2_to(a[2 * i - 2], a[2 * i - 1], nl, nl, u[2 * i], u[2 * i + 1], a[2 * i], a[2 * i + 1]);
    rinv = fast_rcp(a[0]);
  }
}

template <int NB, int W>
__device__ __forceinline__ void diag_rot_steps2(int k_begin, int k_end, float (&a)[NB], float (&u)[NB], float& rinv, float& rm,
                                                float& myr, bool& moved, float* D, int ld, float* stage, int* perm, float* Lt,
                                                int ldlt) {
#pragma unroll 1
  for (int k = k_begin; k < k_end; k += 2) {
    diag_one_step2<NB, W, 0>(k, a, u, rinv, rm, myr, moved, D, ld, stage, perm, Lt, ldlt);
    diag_one_step2<NB, W, 1>(k + 1, a, u, rinv, rm, myr, moved, D, ld, stage, perm, Lt, ldlt);
  }
}

// Lt: transposed copy of the strictly-lower part (Lt[k][i] = L[i][k], leading dimension ldlt), so
// that the look-ahead U12 piece reads columns of L as contiguous vectors.
template <typename T, int MODE, int NB>
__device__ __noinline__ void diag_lu_rot(MPtr<T, MODE> Db, int ld, int o_perm_i, int o_rdiag, int o_flag_i,
                                         int o_stage, int o_lt, int ldlt, int fuse_update) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, NV = NB / VC;
  T* const D = Db.get();
  T* const stage = smem_base<T>() + o_stage;     // 2 x (NB + VC): pivot row + reciprocal + row scale
  T* const Lt = smem_base<T>() + o_lt;
  int* const perm = smem_int(o_perm_i);
  const int lane = threadIdx.x & 31;
  const int li = lane < NB ? lane : NB - 1;      // lanes >= NB mirror the last row and never store
  T* row = D + (size_t)li * ld;
  T a[NB];
  T u[NB];
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    T t[VC];
    vec_get<T>(*reinterpret_cast<const V*>(row + c * VC), t);
#pragma unroll
    for (int q = 0; q < VC; ++q) a[c * VC + q] = t[q];
  }
  if (fuse_update) {
    // look-ahead: this block still lacks the previous step's rank-NB update,
    // block -= L21 (same rows, NB columns to the left) * U12 (same columns, NB rows above)
    const T* lrow = row - NB;
    const T* ub = D - (size_t)NB * ld;
#pragma unroll 1
    for (int p = 0; p < NB; ++p) {
      const T l = lrow[p];
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        T t[VC];
        vec_get<T>(*reinterpret_cast<const V*>(ub + (size_t)p * ld + c * VC), t);
        if constexpr (sizeof(T) == 4) {
          const float nl = -l;
#pragma unroll
          for (int q = 0; q < VC; q += 2) ffma2(a[c * VC + q], a[c * VC + q + 1], nl, nl, t[q], t[q + 1]);
        } else {
#pragma unroll
          for (int q = 0; q < VC; ++q) a[c * VC + q] = fma(-l, t[q], a[c * VC + q]);
        }
      }
    }
  }
  if (lane < NB) perm[lane] = lane;
  bool moved = false;
  T myr = 0;
  // rinv: every lane keeps the reciprocal of its leading entry ready (computed right after a row
  // update, off the step-to-step dependence chain); the pivot lane publishes it with its row and
  // its row scale rm = max |row| of the block row before elimination.
  T rm = 0;
#pragma unroll
  for (int j = 0; j < NB; ++j) rm = fmax(rm, fabs(a[j]));
  T rinv = fast_rcp(a[0]);
  constexpr int Q4 = NB / 4;
  if constexpr (sizeof(T) == 4) {
    diag_rot_steps2<NB, NB>(0, Q4, a, u, rinv, rm, myr, moved, D, ld, stage, perm, Lt, ldlt);
    diag_rot_steps2<NB, 3 * Q4>(Q4, 2 * Q4, a, u, rinv, rm, myr, moved, D, ld, stage, perm, Lt, ldlt);
    diag_rot_steps2<NB, 2 * Q4>(2 * Q4, 3 * Q4, a, u, rinv, rm, myr, moved, D, ld, stage, perm, Lt, ldlt);
    diag_rot_steps2<NB, Q4>(3 * Q4, NB, a, u, rinv, rm, myr, moved, D, ld, stage, perm, Lt, ldlt);
    // a row that was pivot at an even step: a[j] = U[i][i + j]; at an odd step: a[j] = U[i][i - 1 + j], j >= 1
    if (lane < NB) {
      const int off = lane & 1;
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (j >= off && lane - off + j < NB) D[(size_t)lane * ld + lane - off + j] = a[j];
      (smem_base<T>() + o_rdiag)[lane] = myr;
    }
  } else {
    diag_rot_steps<T, NB, NB>(0, Q4, a, u, rinv, rm, myr, moved, D, ld, stage, perm, Lt, ldlt);
    diag_rot_steps<T, NB, 3 * Q4>(Q4, 2 * Q4, a, u, rinv, rm, myr, moved, D, ld, stage, perm, Lt, ldlt);
    diag_rot_steps<T, NB, 2 * Q4>(2 * Q4, 3 * Q4, a, u, rinv, rm, myr, moved, D, ld, stage, perm, Lt, ldlt);
    diag_rot_steps<T, NB, Q4>(3 * Q4, NB, a, u, rinv, rm, myr, moved, D, ld, stage, perm, Lt, ldlt);
    // row i stopped shifting after step i: a[j] = U[i][i + j]
    if (lane < NB) {
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (lane + j < NB) D[(size_t)lane * ld + lane + j] = a[j];
      (smem_base<T>() + o_rdiag)[lane] = myr;
    }
  }
  if (lane == 0) *smem_int(o_flag_i) = moved ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Panel solves of one block step by SUBSTITUTION with the L\U diagonal block (in place, one
// register array per thread):
//   panel_cols: U12 columns [c_lo, c_hi) of rows k0..k0+NB of `Ub`:  U12 = L11^{-1} A12  (one column
//       per thread, left-looking: rows of L11 are read as broadcast vectors)
//   (a rolled tri_piece variant of panel_rows was measured: 14 % slower than this unrolled form, whose
//   instruction count is lower although a quarter of its samples wait on instruction fetch)
//   panel_rows: L21 rows [r_lo, r_hi) of `Cb`:  L21 = A21 U11^{-1}  (one row per thread, right-looking:
//       rows of U11 are read as broadcast vectors, rdiag = 1/diag(U11))
// tix in [0, nthr): index of this thread among the participating ones (tix < 0: not taking part).
template <typename T, int MODE>
__device__ __noinline__ void panel_cols(MPtr<T, MODE> Ub, int ldu, int k0, int c_lo, int c_hi, int tix, int nthr) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, NB = Blk<T>::NB;
  T* const U = Ub.get();
  const T* D = U + (size_t)k0 * ldu + k0;
  if (tix < 0) return;
  for (int t = c_lo + tix; t < c_hi; t += nthr) {
    T a[NB];
    T* col = U + (size_t)k0 * ldu + t;
#pragma unroll
    for (int r = 0; r < NB; ++r) a[r] = col[(size_t)r * ldu];
#pragma unroll
    for (int r = 1; r < NB; ++r) {
      T acc = a[r];
#pragma unroll
      for (int q0 = 0; q0 < r; q0 += VC) {
        T lv[VC];
        vec_get<T>(*reinterpret_cast<const V*>(D + (size_t)r * ldu + q0), lv);
#pragma unroll
        for (int q = 0; q < VC; ++q)
          if (q0 + q < r) acc = fma(-lv[q], a[q0 + q], acc);
      }
      a[r] = acc;
    }
#pragma unroll
    for (int r = 1; r < NB; ++r) col[(size_t)r * ldu] = a[r];
  }
}

template <typename T, int CMODE, int UMODE>
__device__ __noinline__ void panel_rows(MPtr<T, CMODE> Cb, int ldc, MPtr<T, UMODE> Ub, int ldu, int k0, int r_lo,
                                        int r_hi, int o_rdiag, int tix_, int nthr, int tshift) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, NB = Blk<T>::NB;
  T* const C = Cb.get();
  const T* D = Ub.get() + (size_t)k0 * ldu + k0;
  const T* const rdiag = smem_base<T>() + o_rdiag;
  if (tix_ < 0 || r_hi <= r_lo) return;
  // rotate the thread ids so that row tasks land on the threads the column tasks left idle
  int tix = tix_ - (tshift % nthr);
  if (tix < 0) tix += nthr;
  for (int t = r_lo + tix; t < r_hi; t += nthr) {
    T a[NB];
    T* row = C + (size_t)t * ldc + k0;
#pragma unroll
    for (int c0 = 0; c0 < NB; c0 += VC) {
      T tv[VC];
      vec_get<T>(*reinterpret_cast<const V*>(row + c0), tv);
#pragma unroll
      for (int q = 0; q < VC; ++q) a[c0 + q] = tv[q];
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      a[r] *= rdiag[r];
#pragma unroll
      for (int c0 = (r / VC) * VC; c0 < NB; c0 += VC) {
        T uv[VC];
        vec_get<T>(*reinterpret_cast<const V*>(D + (size_t)r * ldu + c0), uv);
#pragma unroll
        for (int q = 0; q < VC; ++q)
          if (c0 + q > r) a[c0 + q] = fma(-a[r], uv[q], a[c0 + q]);
      }
    }
#pragma unroll
    for (int c0 = 0; c0 < NB; c0 += VC) {
      T tv[VC];
#pragma unroll
      for (int q = 0; q < VC; ++q) tv[q] = a[c0 + q];
      *reinterpret_cast<V*>(row + c0) = vec_make(tv);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Look-ahead pieces (ONE warp each, compact rolled loops -- see diag_lu_rot for why): the two
// panel blocks the next diagonal block depends on. Both are private triangular solves per lane
// (no cross-lane traffic): lane t holds NB running values a[j] in registers, pivot p finalises
// x = a[p] (* scale[p]), and a[j] -= x * coef[p][j] for j > p, with the coefficient rows read as
// broadcast vector loads. Pivots go in groups of VC so that the loads stay vector-aligned; after a
// group the registers shift down by VC (static indices only, loop rolled over the groups). Entries
// past the block only ever reach registers that are already dead.
//   lookahead_u12: block (k, k+1) = L11^{-1} A12 in place; lane j owns COLUMN j, coef = Lt (the
//       transposed copy of L11 that diag_lu_rot leaves).
//   lookahead_l21: block (k+1, k) = A21 U11^{-1} in place; lane i owns ROW i, coef = U11 (the
//       diagonal block itself), scale = rdiag.
template <typename T, int NB, bool SCALE, typename Store>
__device__ __forceinline__ void tri_piece(T (&a)[NB], const T* coef, int ldc, const T* scale, Store store) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, NV = NB / VC;
#pragma unroll 1
  for (int g = 0; g < NV; ++g) {
    // vector c of a coefficient row covers positions (g + c) VC ..; past the block (c >= NV - g) only
    // dead registers would be updated, so those loads are clamped onto the last in-block vector --
    // no read ever leaves the block (the neighbouring block is being written by the other piece),
    // and no branch enters the unrolled body
    int coff[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) coff[c] = min(c, NV - 1 - g) * VC;
#pragma unroll
    for (int s_ = 0; s_ < VC; ++s_) {
      const int p = g * VC + s_;
      T x = a[s_];
      if (SCALE) x *= scale[p];
      store(p, x);
      const T* cr = coef + (size_t)p * ldc + g * VC;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        if (c * VC + VC - 1 <= s_) continue;
        T t[VC];
        vec_get<T>(*reinterpret_cast<const V*>(cr + coff[c]), t);
        if constexpr (sizeof(T) == 4) {
          // packed pairs; an entry <= s_ inside a pair is already dead, updating it is harmless
          const float nx = -x;
#pragma unroll
          for (int q = 0; q < VC; q += 2)
            if (c * VC + q + 1 > s_) ffma2(a[c * VC + q], a[c * VC + q + 1], nx, nx, t[q], t[q + 1]);
        } else {
#pragma unroll
          for (int q = 0; q < VC; ++q)
            if (c * VC + q > s_) a[c * VC + q] = fma(-x, t[q], a[c * VC + q]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j + VC < NB; ++j) a[j] = a[j + VC];
  }
}

template <typename T, int MODE, int NB>
__device__ __noinline__ void lookahead_u12(MPtr<T, MODE> Db, int ld, int o_lt, int ldlt) {
  T* const B = Db.get() + NB;                     // block (k, k+1): same rows, next NB columns
  const T* const Lt = smem_base<T>() + o_lt;
  const int lane = threadIdx.x & 31;
  const int lj = lane < NB ? lane : NB - 1;
  T a[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) a[i] = B[(size_t)i * ld + lj];
  __syncwarp();                                    // lanes >= NB mirror the last column: reads before its stores
  tri_piece<T, NB, false>(a, Lt, ldlt, nullptr, [&](int p, T x) { if (lane < NB) B[(size_t)p * ld + lj] = x; });
  __syncwarp();
}

template <typename T, int MODE, int NB>
__device__ __noinline__ void lookahead_l21(MPtr<T, MODE> Db, int ld, int o_rdiag) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, NV = NB / VC;
  const T* const D = Db.get();
  const T* const rdiag = smem_base<T>() + o_rdiag;
  const int lane = threadIdx.x & 31;
  const int li = lane < NB ? lane : NB - 1;
  T* const row = Db.get() + (size_t)(NB + li) * ld;   // block (k+1, k): next NB rows, same columns
  T a[NB];
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    T t[VC];
    vec_get<T>(*reinterpret_cast<const V*>(row + c * VC), t);
#pragma unroll
    for (int q = 0; q < VC; ++q) a[c * VC + q] = t[q];
  }
  __syncwarp();                                    // all rows are in registers before any store
  tri_piece<T, NB, true>(a, D, ld, rdiag, [&](int p, T x) { if (lane < NB) row[p] = x; });
  __syncwarp();
}

// Apply the block's row interchanges (rows k0..k0+NB of `Cb`) to columns [c_begin, c_end), except
// the diagonal block's own columns when skip_diag. Each thread owns one column: no barrier between
// its reads and writes. perm (region base, int units): perm[k0 + r] = block-local source row.
template <typename T, int CMODE>
__device__ __noinline__ void apply_block_perm(MPtr<T, CMODE> Cb, int ldc, int k0, int o_perm_i, int c_begin, int c_end,
                                              int skip_diag) {
  constexpr int NB = Blk<T>::NB;
  T* const C = Cb.get();
  const int* const perm = smem_int(o_perm_i);
  for (int col = c_begin + threadIdx.x; col < c_end; col += blockDim.x) {
    if (skip_diag && col >= k0 && col < k0 + NB) continue;
    T tmp[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) tmp[r] = C[(size_t)(k0 + perm[k0 + r]) * ldc + col];
#pragma unroll
    for (int r = 0; r < NB; ++r) C[(size_t)(k0 + r) * ldc + col] = tmp[r];
  }
}

// ---------------------------------------------------------------------------------------------
// Factor one square region: rows/cols [0, sz) of the matrix at Ab (ld) with ncols_total columns,
// optionally with extra "low" rows [sz, sz+nlow) held in Lb (ldl) that only carry the L panel
// (columns [0, sz)) -- the split case's phase 1.
//
// The diagonal block is a serial job for ONE warp (~10-16k cycles) while panels + trailing update
// of a step keep 16 warps busy for about as long, so the two are overlapped (look-ahead): in step k
//   warps 0,1: U12 / L21 of the NEXT block only (blocks (k,k+1), (k+1,k)), one each -> named barrier 1
//   warp 0   : update of block (k+1,k+1) fused with its L\U factorisation (perm / flag of block k+1)
//   warps 2+ : panels of step k for everything else -> named barrier 1
//   warps 1+ : trailing update of everything except block (k+1,k+1)
// and one CTA barrier closes the step. The row interchanges found in block k+1 are applied to the
// rest of its block row (and to the shadow array Sb) at the start of step k+1.
// The diagonal blocks are left as L\U; their inverses are formed afterwards for all blocks in
// parallel (lu_invert_diag_blocks). rdiag is indexed by the GLOBAL row (offset o_rdiag + row0).
// Sb (shadow_cols > 0): rows [0,sz) x [0,shadow_cols) of another shared array that must follow the
// row interchanges (the L21 rows of the split's second half). row0: global index of this region's
// first row (perm / rdiag position).
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <typename T, int MODE>
__device__ __forceinline__ void lu_region(MPtr<T, MODE> Ab, int ld, int sz, int ncols_total, MPtr<T, 0> Lb, int ldl,
                                          int nlow, LuVec lv, int row0, long long* prof, MPtr<T, 0> Sb, int ldsh,
                                          int shadow_cols, bool first_done) {
  constexpr int NB = Blk<T>::NB;
  // Roles. The serial look-ahead chain (role 0) runs on the LAST warp (the SMSP arbiter prefers the
  // highest warp id), the L21-piece warp (role 1) on another SMSP, the bulk on the rest. (Keeping
  // the chain's SMSP free of bulk warps was measured: chain -5%, bulk -12%, net loss.)
  const int lane = threadIdx.x & 31, nw = blockDim.x >> 5, pw = threadIdx.x >> 5;
  const int chain_w = nw - 1, piece_w = nw - 2;
  auto is_quiet = [&](int) { return false; };
  int bidx = 0, nbulk = 0;
  for (int w = 0; w < nw; ++w) {
    const bool bulk = w != chain_w && w != piece_w && !is_quiet(w);
    if (bulk && w < pw) ++bidx;
    if (bulk) ++nbulk;
  }
  // role index "warp": 0 = chain, 1 = piece, 2.. = bulk, -1 = quiet
  const int warp = pw == chain_w ? 0 : (pw == piece_w ? 1 : (is_quiet(pw) ? -1 : 2 + bidx));
  const bool is_timer = warp == 2 && lane == 0;
  long long t0 = 0;
  auto lap = [&](int idx) {
    if (prof && is_timer) { const long long t = clock64(); prof[idx] += t - t0; t0 = t; }
  };
  if (prof && is_timer) t0 = clock64();
  const int o_perm_i = lv.o_perm_i + row0;
  const int o_rdiag = lv.o_rdiag + row0;
  // low rows are addressed as rows [sz, sz+nlow) of a re-based array
  const MPtr<T, 0> Lrb = Lb.plus(-(long long)sz * ldl);
  if (warp == 0 && !first_done) {                 // first_done: the caller factored block 0 already
    const long long c0 = prof ? clock64() : 0;
    diag_lu_rot<T, MODE, NB>(Ab, ld, o_perm_i, o_rdiag, lv.o_flag_i, lv.o_stage, lv.o_lt, lv.ldlt, 0);
    if (prof && lane == 0) { prof[11] += 1; if (*smem_int(lv.o_flag_i)) prof[10] += 1; prof[6] += clock64() - c0; }
  }
  __syncthreads();
  lap(13);
  for (int k0 = 0, kb = 0; k0 < sz; k0 += NB, ++kb) {
    const int r0 = k0 + NB;
    const bool has_next = r0 < sz;
    if (*smem_int(lv.o_flag_i + (kb & 1))) {
      apply_block_perm<T, MODE>(Ab, ld, k0, o_perm_i, 0, ncols_total, 1);
      if (shadow_cols > 0) apply_block_perm<T, 0>(Sb, ldsh, k0, o_perm_i, 0, shadow_cols, 0);
      __syncthreads();
    }
    lap(13);                                                     // (a barrier blocks at its first consumer)
    if (has_next && (warp == 0 || warp == 1)) {
      // ---- look-ahead group: warp 0 -> U12 block, warp 1 -> L21 block (in parallel)
      const long long c0 = prof ? clock64() : 0;
      const MPtr<T, MODE> Db = Ab.plus((long long)k0 * ld + k0);
      if (warp == 0) lookahead_u12<T, MODE, NB>(Db, ld, lv.o_lt, lv.ldlt);
      else lookahead_l21<T, MODE, NB>(Db, ld, o_rdiag + k0);
      named_bar_sync(2, 64);                                     // the two pieces see each other
      if (warp == 0) {
        named_bar_arrive(1, blockDim.x);                         // release them to everyone (warp 1 syncs below)
        // (a separate two-warp rank_nb_update of block (k+1,k+1) was measured: its cold code costs more
        // than the compact fused loop inside diag_lu_rot saves)
        const long long c1 = prof ? clock64() : 0;
        diag_lu_rot<T, MODE, NB>(Ab.plus((long long)r0 * ld + r0), ld, o_perm_i + r0, o_rdiag + r0,
                                 lv.o_flag_i + ((kb + 1) & 1), lv.o_stage, lv.o_lt, lv.ldlt, 1);
        if (prof && lane == 0) {
          prof[11] += 1; if (*smem_int(lv.o_flag_i + ((kb + 1) & 1))) prof[10] += 1;
          const long long c2 = clock64(); prof[12] += c1 - c0; prof[6] += c2 - c1;
        }
      }
    }
    if (!(has_next && warp == 0)) {
      // ---- panels: bulk warps; update: bulk + piece warp; the last step: every warp
      int tix, nthr, wid, nwk;
      if (has_next) {
        tix = warp >= 2 ? (warp - 2) * 32 + lane : -1; nthr = nbulk * 32;
        wid = warp >= 1 ? warp - 1 : -1; nwk = nbulk + 1;
      } else {
        tix = threadIdx.x; nthr = blockDim.x; wid = pw; nwk = nw;
      }
      const int skip = has_next ? NB : 0;                       // block column / row k+1: look-ahead group
      const int ncol_tasks = max(ncols_total - r0 - skip, 0), nrow_tasks = max(sz - r0 - skip, 0);
      panel_cols<T, MODE>(Ab, ld, k0, r0 + skip, ncols_total, tix, nthr);
      panel_rows<T, MODE, MODE>(Ab, ld, Ab, ld, k0, r0 + skip, sz, o_rdiag + k0, tix, nthr, ncol_tasks);
      if (nlow > 0)
        panel_rows<T, 0, MODE>(Lrb, ldl, Ab, ld, k0, sz, sz + nlow, o_rdiag + k0, tix, nthr, ncol_tasks + nrow_tasks);
      if (has_next) named_bar_sync(1, blockDim.x); else __syncthreads();
      lap(7);
      // trailing update minus block (k+1,k+1): block row k+1 right of it, the rows below, the low rows
      int dealt = 0;
      if (has_next) {
        dealt += rank_nb_update_auto<T, MODE, MODE>(Ab, ld, Ab, ld, k0, r0, r0 + NB, r0 + NB, ncols_total, wid, nwk, dealt);
        dealt += rank_nb_update_auto<T, MODE, MODE>(Ab, ld, Ab, ld, k0, r0 + NB, sz, r0, ncols_total, wid, nwk, dealt);
      }
      if (nlow > 0) dealt += rank_nb_update_auto<T, 0, MODE>(Lrb, ldl, Ab, ld, k0, sz, sz + nlow, r0, sz, wid, nwk, dealt);
      lap(8);
    }
    __syncthreads();
    lap(13);                                                     // warps 1+ waiting for the look-ahead warp
  }
}

// Diagonal-block inverses: one warp per triangle, lane j owns COLUMN j of the result and runs the
// whole substitution for it in registers (L x_j = e_j / U x_j = e_j); the L\U entries are read as
// broadcast vector loads -- no shuffles, no cross-lane dependence. Called by ALL warps of the CTA
// (a warp without a job passes active = false): a CTA barrier separates every job's reads of the
// block from the in-place stores.
//   lower: X = inv(L) (unit diagonal implicit, strict lower part stored);
//   upper: X = inv(U) (rdiag = 1/diag(U)).
template <typename T, int MODE, int NB>
__device__ __noinline__ void diag_inverse_job(MPtr<T, MODE> Db, int ld, int o_rdiag, bool upper, bool active) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC;
  T* const D = Db.get();
  const T* const rdiag = smem_base<T>() + o_rdiag;
  const int lane = threadIdx.x & 31;
  const int lj = lane < NB ? lane : NB - 1;
  T x[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) x[j] = 0;
  if (active) {
    if (!upper) {
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        T acc0 = (i == lj) ? T(1) : T(0), acc1 = 0;
#pragma unroll
        for (int q0 = 0; q0 < i; q0 += VC) {
          T l[VC];
          vec_get<T>(*reinterpret_cast<const V*>(D + (size_t)i * ld + q0), l);
#pragma unroll
          for (int q = 0; q < VC; ++q)
            if (q0 + q < i) { if (q & 1) acc1 = fma(-l[q], x[q0 + q], acc1); else acc0 = fma(-l[q], x[q0 + q], acc0); }
        }
        x[i] = acc0 + acc1;
      }
    } else {
#pragma unroll
      for (int i = NB - 1; i >= 0; --i) {
        T acc0 = (i == lj) ? T(1) : T(0), acc1 = 0;
#pragma unroll
        for (int q0 = (i / VC) * VC; q0 < NB; q0 += VC) {
          T uu[VC];
          vec_get<T>(*reinterpret_cast<const V*>(D + (size_t)i * ld + q0), uu);
#pragma unroll
          for (int q = 0; q < VC; ++q)
            if (q0 + q > i) { if (q & 1) acc1 = fma(-uu[q], x[q0 + q], acc1); else acc0 = fma(-uu[q], x[q0 + q], acc0); }
        }
        x[i] = (acc0 + acc1) * rdiag[i];
      }
    }
  }
  __syncthreads();                               // every job has read its block
  if (active && lane < NB) {
#pragma unroll
    for (int i = 0; i < NB; ++i)
      if (upper ? (i <= lane) : (i > lane)) D[(size_t)i * ld + lane] = x[i];
  }
}

// All diagonal blocks of the factored view, both triangles, one warp per job.
template <typename T, int MODE>
__device__ __forceinline__ void lu_invert_diag_blocks(const TView<T, MODE>& v, int o_rdiag) {
  constexpr int NB = Blk<T>::NB;
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nblk = v.mp / NB, njobs = 2 * nblk;
  for (int j0 = 0; j0 < njobs; j0 += nw) {
    const int job = j0 + warp;
    const bool active = job < njobs;
    const int b = active ? job >> 1 : 0, k0 = b * NB;
    // logical block (k0,k0): rows/cols < m1 in main at (k0,k0); the rest at (k0-m1, k0) (S22 region)
    const MPtr<T, MODE> Db = v.main_.plus(k0 < v.m1 ? (long long)k0 * v.ld + k0 : (long long)(k0 - v.m1) * v.ld + k0);
    diag_inverse_job<T, MODE, NB>(Db, v.ld, o_rdiag + k0, (job & 1) != 0, active);
    __syncthreads();
  }
}

// Schur block of the split: S22 = T22 - L21 U12, accumulated in registers (K = m1), T22 = R22 + diag
// read from the L2 copy `R22` (ld = ldr, already offset to (m1,m1)); U12 is spilled to v.u12 and S22
// written over it in main[0..n2) x [m1, mp). o_dinv: shared vector 1/d (padded with ones).
template <typename T, int MODE>
__device__ __noinline__ void split_schur(TView<T, MODE> v, const T* __restrict__ R22, int ldr, int o_dinv, int m_real) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC;
  constexpr int TR = 4, WR = 4 * TR, WC = 16 * VC;
  const int m1 = v.m1, n2 = v.mp - v.m1;
  T* const vmain = v.main();
  T* const vlow = v.low();
  const T* const dinv = smem_base<T>() + o_dinv;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 3, cg = lane & 7;
  const int ntr = (n2 + WR - 1) / WR, ntc = (n2 + WC - 1) / WC;
  // every warp owns at most one tile whose accumulators stay in registers across the spill
  // (the planner only selects the split when ntr * ntc <= number of warps)
  T acc[TR][2 * VC];
  const int wt = warp;
  const bool have = wt < ntr * ntc;
  const int tr_ = have ? wt / ntc : 0, tc_ = have ? wt - tr_ * ntc : 0;
  const int rbase = tr_ * WR + g, c0 = tc_ * WC + VC * cg, c1 = c0 + 8 * VC;
#pragma unroll
  for (int r = 0; r < TR; ++r)
#pragma unroll
    for (int c = 0; c < 2 * VC; ++c) acc[r][c] = 0;
  if (have) {
    const T* lp[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) lp[r] = vlow + (size_t)min(rbase + 4 * r, n2 - 1) * v.ldl;
#pragma unroll 2
    for (int kc = 0; kc < m1; kc += VC) {
      T l[TR][VC];
#pragma unroll
      for (int r = 0; r < TR; ++r) vec_get<T>(*reinterpret_cast<const V*>(lp[r] + kc), l[r]);
#pragma unroll
      for (int kk = 0; kk < VC; ++kk) {
        const T* ur = vmain + (size_t)(kc + kk) * v.ld + m1;
        T u[2 * VC];
        {
          T lo[VC], hi[VC];
          vec_get<T>(*reinterpret_cast<const V*>(ur + min(c0, n2 - VC)), lo);
          vec_get<T>(*reinterpret_cast<const V*>(ur + min(c1, n2 - VC)), hi);
#pragma unroll
          for (int q = 0; q < VC; ++q) { u[q] = lo[q]; u[VC + q] = hi[q]; }
        }
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
          for (int c = 0; c < 2 * VC; ++c) acc[r][c] = fma(l[r][kk], u[c], acc[r][c]);
      }
    }
  }
  // T22 = R22 + diag comes from the L2 copy: issue those loads now, they complete under the spill
  T t22v[TR][2 * VC];
#pragma unroll
  for (int r = 0; r < TR; ++r)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int q = 0; q < VC; ++q) {
        const int i = rbase + 4 * r, c = (hh ? c1 : c0) + q;
        const int gi = m1 + i, gj = m1 + c;             // global indices in the padded matrix
        t22v[r][hh * VC + q] = (have && i < n2 && c < n2 && gi < m_real && gj < m_real) ? R22[(size_t)i * ldr + c] : T(0);
      }
  __syncthreads();                       // all reads of U12 done
  // spill U12 (final) to L2: main rows [0,m1) x cols [m1,mp) -> u12 [m1, n2]
  for (int t = threadIdx.x; t < m1 * (n2 / VC); t += blockDim.x) {
    const int i = t / (n2 / VC), jv = t - i * (n2 / VC);
    *reinterpret_cast<V*>(v.u12 + (size_t)i * n2 + jv * VC) = *reinterpret_cast<const V*>(vmain + (size_t)i * v.ld + m1 + jv * VC);
  }
  __syncthreads();
  // S22 = T22 - acc -> main[0..n2) x [m1, mp)
  if (have) {
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
          const int gi = m1 + i, gj = m1 + c + q;
          T t22 = t22v[r][hh * VC + q];
          if (gi == gj) t22 += dinv[gi];
          out[q] = t22 - acc[r][hh * VC + q];
        }
        *reinterpret_cast<V*>(vmain + (size_t)i * v.ld + m1 + c) = vec_make(out);
      }
    }
  }
  __syncthreads();
}

// Full factorisation of the view (T must already be loaded: main rows [0,m1) all columns, low rows
// [m1,mp) columns [0,m1)). R22/ldr/o_dinv/m_real are only used by the split.
template <typename T, int MODE>
__device__ __forceinline__ void lu_factor_view(const TView<T, MODE>& v, const T* R22, int ldr, int o_dinv, int m_real,
                                               LuVec lv, long long* prof, bool first_done = false) {
  MPtr<T, 0> none; none.off = 0; none.g = nullptr;
  if (MODE != 1) {
    lu_region<T, MODE>(v.main_, v.ld, v.mp, v.mp, none, 0, 0, lv, 0, prof, none, 0, 0, first_done);
  } else {
    lu_region<T, MODE>(v.main_, v.ld, v.m1, v.mp, v.low_, v.ldl, v.mp - v.m1, lv, 0, prof, none, 0, 0, first_done);
    split_schur<T, MODE>(v, R22, ldr, o_dinv, m_real);
    lu_region<T, MODE>(v.main_.plus(v.m1), v.ld, v.mp - v.m1, v.mp - v.m1, none, 0, 0, lv, v.m1, prof, v.low_, v.ldl, v.m1, false);
  }
  long long t0 = (prof && threadIdx.x == 0) ? clock64() : 0;
  lu_invert_diag_blocks<T, MODE>(v, lv.o_rdiag);
  if (prof && threadIdx.x == 0) prof[9] += clock64() - t0;
}

// ---------------------------------------------------------------------------------------------
// x[mp] (shared, element offset o_x) <- T^{-1} x using the factors above. o_tmp: mp scratch elements.
template <typename T, int MODE>
__device__ __noinline__ void lu_solve_view(TView<T, MODE> v, int o_perm_i, int o_x, int o_tmp) {
  using V = typename VecOf<T>::type;
  constexpr int VC = VecOf<T>::VC, NB = Blk<T>::NB;
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
  const int mp = v.mp, m1 = v.m1, n2 = mp - m1;
  const T* const vmain = v.main();
  const T* const vlow = (MODE == 1) ? v.low() : vmain;
  const int* const perm = smem_int(o_perm_i);
  T* const x = smem_base<T>() + o_x;
  T* const tmp = smem_base<T>() + o_tmp;
  for (int i = tid; i < mp; i += NT) tmp[i] = x[(i / NB) * NB + perm[i]];
  __syncthreads();
  for (int i = tid; i < mp; i += NT) x[i] = tmp[i];
  __syncthreads();
  auto lrow = [&](int i, int k0) -> const T* {      // L entries of logical row i, block column k0
    if (k0 < m1) return (i < m1 ? vmain + (size_t)i * v.ld : vlow + (size_t)(i - m1) * v.ldl) + k0;
    return vmain + (size_t)(i - m1) * v.ld + k0;   // S22 region: row i-m1, column m1 + (k0-m1)
  };
  auto dblk = [&](int k0) -> const T* {
    return k0 < m1 ? vmain + (size_t)k0 * v.ld + k0 : vmain + (size_t)(k0 - m1) * v.ld + k0;
  };
  // Diagonal blocks hold explicit inverses, so a block step is a 32x32 mat-vec: NT/NB threads per
  // row (partial dot products + shuffle reduction) instead of one thread per row. The block result
  // goes to the other vector (x -> tmp going down, tmp -> x coming back), so reads and writes of a
  // step never alias.
  const int tpr = NT / NB;                         // threads per row: power of two, <= 32
  const int drow = tid / tpr, dpart = tid - drow * tpr;
  // ---- L y = x (right-looking); y accumulates in tmp
  for (int k0 = 0; k0 < mp; k0 += NB) {
    {
      const T* row = dblk(k0) + (size_t)drow * v.ld;
      T acc = 0;
      for (int q = dpart; q < drow; q += tpr) acc = fma(row[q], x[k0 + q], acc);
      for (int o = tpr >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
      if (dpart == 0) tmp[k0 + drow] = x[k0 + drow] + acc;
    }
    __syncthreads();
    for (int i = k0 + NB + tid; i < mp; i += NT) {
      const T* row = lrow(i, k0);
      T acc = x[i], a2 = 0;
#pragma unroll
      for (int c0 = 0; c0 < NB; c0 += 2 * VC) {
        T a[VC], y[VC], b[VC], z[VC];
        vec_get<T>(*reinterpret_cast<const V*>(row + c0), a);
        vec_get<T>(*reinterpret_cast<const V*>(tmp + k0 + c0), y);
        vec_get<T>(*reinterpret_cast<const V*>(row + c0 + VC), b);
        vec_get<T>(*reinterpret_cast<const V*>(tmp + k0 + c0 + VC), z);
#pragma unroll
        for (int q = 0; q < VC; ++q) { acc = fma(-a[q], y[q], acc); a2 = fma(-b[q], z[q], a2); }
      }
      x[i] = acc + a2;
    }
    __syncthreads();
  }
  // ---- U x = y (y in tmp): second-half blocks, then the U12 coupling (from L2), then first-half blocks
  for (int k0 = mp - NB; k0 >= 0; k0 -= NB) {
    {
      const T* row = dblk(k0) + (size_t)drow * v.ld;
      T acc = 0;
      for (int c = drow + dpart; c < NB; c += tpr) acc = fma(row[c], tmp[k0 + c], acc);
      for (int o = tpr >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
      if (dpart == 0) x[k0 + drow] = acc;
    }
    __syncthreads();
    const int top = k0 < m1 ? 0 : m1;
    for (int i = top + tid; i < k0; i += NT) {
      const T* row = (k0 < m1 ? vmain + (size_t)i * v.ld : vmain + (size_t)(i - m1) * v.ld) + k0;
      T acc = tmp[i], a2 = 0;
#pragma unroll
      for (int c0 = 0; c0 < NB; c0 += 2 * VC) {
        T a[VC], y[VC], b[VC], z[VC];
        vec_get<T>(*reinterpret_cast<const V*>(row + c0), a);
        vec_get<T>(*reinterpret_cast<const V*>(x + k0 + c0), y);
        vec_get<T>(*reinterpret_cast<const V*>(row + c0 + VC), b);
        vec_get<T>(*reinterpret_cast<const V*>(x + k0 + c0 + VC), z);
#pragma unroll
        for (int q = 0; q < VC; ++q) { acc = fma(-a[q], y[q], acc); a2 = fma(-b[q], z[q], a2); }
      }
      tmp[i] = acc + a2;
    }
    __syncthreads();
    if (MODE == 1 && k0 == m1) {
      // x1 -= U12 x2, U12 [m1, n2] in L2: one warp per row, 4 rows in flight, coalesced
      const int warp = tid >> 5, nw = NT >> 5;
      for (int i0 = warp * 4; i0 < m1; i0 += nw * 4) {
        T acc[4] = {0, 0, 0, 0};
        for (int j = lane * VC; j < n2; j += 32 * VC) {
          T y[VC];
          vec_get<T>(*reinterpret_cast<const V*>(x + m1 + j), y);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            T a[VC];
            vec_get<T>(*reinterpret_cast<const V*>(v.u12 + (size_t)min(i0 + q4, m1 - 1) * n2 + j), a);
#pragma unroll
            for (int q = 0; q < VC; ++q) acc[q4] = fma(a[q], y[q], acc[q4]);
          }
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) acc[q4] = warp_reduce(acc[q4], OpSum());
        if (lane == 0) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            if (i0 + q4 < m1) tmp[i0 + q4] -= acc[q4];
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace lcpb200
