// lcp_solver.cuh -- per-scene PDIPM forward / implicit-diff backward kernels.
//
// Restates (B200-native, one CTA per scene, persistent grid):
//   pdipm.py:357-408 pre_factor_kkt, :414-454 factor_kkt, :325-354 solve_kkt,
//   :49-179 forward, :182-186 get_step, lcp.py:22-64 LCPFunction.forward/backward.
#pragma once
#include "lcp_device.cuh"

namespace lcpb200 {

// Launch plan, computed on the host (plan.cu) and passed by value.
struct Plan {
  int n, m, e;
  int nt;                     // threads per CTA
  int grid;                   // CTAs (persistent)
  int ldT, ldG, ldQi;         // leading dimensions of T, G copy, Qinv
  int T_smem, G_smem, Qi_smem;
  int off_T, off_G, off_Qi, off_vec;   // shared offsets, in elements
  int smem_bytes;
  long long ws_per_cta;       // workspace elements per CTA
  long long w_Qi, w_R, w_T, w_X, w_XA, w_S11, w_V, w_W;   // workspace offsets (elements)
};

// Shared-memory vectors. `carve` is used with base == nullptr on the host to size the block.
template <typename T>
struct Vecs {
  T *x, *s, *z, *y, *d;
  T *rx, *rz, *ry;
  T *hz, *hy, *te;            // Schur rhs / solution pieces
  T *tn, *tn2;                // n-length temporaries
  T *dxa, *dsa, *dza, *dya;   // affine direction
  T *dxc, *dsc, *dzc, *dyc;   // corrector direction
  T *rs2;
  T *scratch;                 // blockDim.x elements
  T *red;                     // 128 elements
  int *perm;                  // m ints: block-local LU row interchanges
  __host__ __device__ long long carve(T* base, int n, int m, int e, int nt) {
    long long o = 0;
#define LCPB200_TAKE(ptr, cnt) do { ptr = base + o; o += ((cnt) + 3) & ~3; } while (0)
    LCPB200_TAKE(x, n); LCPB200_TAKE(s, m); LCPB200_TAKE(z, m); LCPB200_TAKE(y, e); LCPB200_TAKE(d, m);
    LCPB200_TAKE(rx, n); LCPB200_TAKE(rz, m); LCPB200_TAKE(ry, e);
    LCPB200_TAKE(hz, m); LCPB200_TAKE(hy, e); LCPB200_TAKE(te, e);
    LCPB200_TAKE(tn, n); LCPB200_TAKE(tn2, n);
    LCPB200_TAKE(dxa, n); LCPB200_TAKE(dsa, m); LCPB200_TAKE(dza, m); LCPB200_TAKE(dya, e);
    LCPB200_TAKE(dxc, n); LCPB200_TAKE(dsc, m); LCPB200_TAKE(dzc, m); LCPB200_TAKE(dyc, e);
    LCPB200_TAKE(rs2, m);
    LCPB200_TAKE(scratch, (nt > n ? nt : n) > m ? (nt > n ? nt : n) : m); LCPB200_TAKE(red, 128);
    { T* pp; LCPB200_TAKE(pp, m); perm = reinterpret_cast<int*>(pp); }
#undef LCPB200_TAKE
    return o;
  }
};

template <typename T>
struct SceneCtx {
  int n, m, e;
  const T *Q, *G, *A, *F;     // this scene's inputs (G may point at the shared copy)
  int ldG;
  T *Qi; int ldQi;
  T *Tm; int ldT;
  T *R, *X, *XA, *S11, *Vm, *W;
  int* lu_flag;               // shared int used by lu_blocked
  Vecs<T> v;
};

// ------------------------------------------------------------------ pre_factor_kkt (pdipm.py:357-408)
// Returns false (uniformly) when Q is singular.
template <typename T>
__device__ bool prefactor(SceneCtx<T>& c, int* flag) {
  const int n = c.n, m = c.m, e = c.e, tid = threadIdx.x, NT = blockDim.x;
  if (tid == 0) *flag = 0;
  for (int t = tid; t < n * n; t += NT) { int i = t / n, j = t - i * n; c.Qi[(size_t)i * c.ldQi + j] = c.Q[t]; }
  __syncthreads();
  invert_inplace(c.Qi, c.ldQi, n, flag, c.v.scratch);          // :362  Q^{-1} instead of LU(Q)
  const bool singular = (*flag != 0);
  __syncthreads();
  if (singular) return false;
  // X = Q^{-1} G^T ; R = G X + F                                 :378-379
  gemm_tiled<T, true>(c.X, m, c.Qi, c.ldQi, c.G, c.ldG, n, m, n, T(1), T(0));
  for (int t = tid; t < m * m; t += NT) c.R[t] = c.F[t];
  __syncthreads();
  gemm_tiled<T, false>(c.R, m, c.G, c.ldG, c.X, m, m, m, n, T(1), T(1));
  if (e > 0) {
    // XA = Q^{-1} A^T (:383), S11 = A XA (:384), V = G XA (:385)
    gemm_tiled<T, true>(c.XA, e, c.Qi, c.ldQi, c.A, n, n, e, n, T(1), T(0));
    gemm_tiled<T, false>(c.S11, e, c.A, n, c.XA, e, e, e, n, T(1), T(0));
    gemm_tiled<T, false>(c.Vm, e, c.G, c.ldG, c.XA, e, m, e, n, T(1), T(0));
    invert_inplace(c.S11, e, e, flag, c.v.scratch);            // :387  (A Q^{-1} A^T)^{-1}
    // W = S11^{-1} V^T (:395, the reference reuses (G Q^{-1} A^T)^T here) ; R -= V W (:403)
    gemm_tiled<T, true>(c.W, m, c.S11, e, c.Vm, e, e, m, e, T(1), T(0));
    gemm_tiled<T, false>(c.R, m, c.Vm, e, c.W, m, m, m, e, T(-1), T(1));
  }
  return true;
}

// ------------------------------------------------------------------ factor_kkt (pdipm.py:414-454)
template <typename T>
__device__ void factor_kkt(SceneCtx<T>& c, const T* d) {
  const int m = c.m;
  for (int t = threadIdx.x; t < m * m; t += blockDim.x) {
    const int i = t / m, j = t - i * m;
    T val = c.R[t];
    if (i == j) val += T(1) / d[i];                              // :427-429
    c.Tm[(size_t)i * c.ldT + j] = val;
  }
  __syncthreads();
  lu_blocked(c.Tm, c.ldT, m, c.v.perm, c.lu_flag);                // :431 (block-local pivoting)
}

// ------------------------------------------------------------------ solve_kkt (pdipm.py:325-354)
// rx / rz / ry may be nullptr (== zero vector). Outputs may not alias the inputs.
template <typename T>
__device__ void solve_kkt(SceneCtx<T>& c, const T* d, const T* rx, const T* rs, const T* rz, const T* ry,
                          T* dx, T* ds, T* dz, T* dy) {
  const int n = c.n, m = c.m, e = c.e, tid = threadIdx.x, NT = blockDim.x;
  Vecs<T>& v = c.v;
  T* t = v.tn;                                                   // Q^{-1} rx      :333
  if (rx) {
    gemv_rows(c.Qi, c.ldQi, n, n, rx, [&](int i, T a) { t[i] = a; });
    // hz = G t + rs/d - rz ; hy = A t - ry                       :337-340
    gemv_rows(c.G, c.ldG, m, n, t, [&](int i, T a) { v.hz[i] = a + rs[i] / d[i] - (rz ? rz[i] : T(0)); });
    if (e > 0) gemv_rows(c.A, n, e, n, t, [&](int i, T a) { v.hy[i] = a - (ry ? ry[i] : T(0)); });
  } else {
    for (int i = tid; i < m; i += NT) v.hz[i] = rs[i] / d[i] - (rz ? rz[i] : T(0));
    for (int i = tid; i < e; i += NT) v.hy[i] = -(ry ? ry[i] : T(0));
    __syncthreads();
  }
  // w = -S^{-1} h by block elimination of the e x e block          :342
  if (e > 0) {
    gemv_rows(c.S11, e, e, e, v.hy, [&](int i, T a) { v.te[i] = a; });
    gemv_rows(c.Vm, e, m, e, v.te, [&](int i, T a) { v.hz[i] -= a; });
  }
  lu_solve_vec(c.Tm, c.ldT, m, v.perm, v.hz, v.scratch);         // hz <- T^{-1}(..) = -w_z
  if (e > 0) {
    gemv_rows(c.W, m, e, m, v.hz, [&](int i, T a) { dy[i] = -(v.te[i] - a); });   // w_y
  }
  for (int i = tid; i < m; i += NT) {
    const T wz = -v.hz[i];
    dz[i] = wz;                                                  // :351
    ds[i] = (-rs[i] - wz) / d[i];                                // :347,350
  }
  __syncthreads();
  // g1 = -rx - G^T w_z - A^T w_y ; dx = Q^{-1} g1                 :344-349
  T* g1 = v.tn2;
  gemv_cols(c.G, c.ldG, m, n, dz, v.scratch, [&](int j, T a) { g1[j] = -(rx ? rx[j] : T(0)) - a; });
  if (e > 0) gemv_cols(c.A, n, e, n, dy, v.scratch, [&](int j, T a) { g1[j] -= a; });
  gemv_rows(c.Qi, c.ldQi, n, n, g1, [&](int i, T a) { dx[i] = a; });
}

// ------------------------------------------------------------------ get_step (pdipm.py:182-186), per scene
// Returns (step(z,dz), step(s,ds)) with torch's NaN semantics.
template <typename T>
__device__ void get_steps(const T* z, const T* dz, const T* s, const T* ds, int m, T* red, T& step_z, T& step_s) {
  const T NEG_INF = -INFINITY, POS_INF = INFINITY;
  T mx[2] = {NEG_INF, NEG_INF};
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    mx[0] = nan_max(mx[0], -z[i] / dz[i]);
    mx[1] = nan_max(mx[1], -s[i] / ds[i]);
  }
  block_reduce<T, 2>(mx, OpMax(), NEG_INF, red);
  const T fz = (mx[0] > T(1)) ? mx[0] : T(1);                   // python max(1.0, a.max()): NaN -> 1.0
  const T fs = (mx[1] > T(1)) ? mx[1] : T(1);
  T mn[2] = {POS_INF, POS_INF};
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const T az = (dz[i] > T(0)) ? fz : (-z[i] / dz[i]);
    const T as = (ds[i] > T(0)) ? fs : (-s[i] / ds[i]);
    mn[0] = nan_min(mn[0], az);
    mn[1] = nan_min(mn[1], as);
  }
  block_reduce<T, 2>(mn, OpMin(), POS_INF, red);
  step_z = mn[0];
  step_s = mn[1];
}

template <typename T>
struct FwdArgs {
  Plan P;
  int B;
  const T *Q, *p, *G, *h, *A, *b, *F;
  T *zhat, *nu, *lam, *slack, *resid;
  int *status, *iters;
  T eps;
  int not_improved_lim, max_iter;
  T* ws;
};

template <typename T>
__device__ void setup_ctx(SceneCtx<T>& c, const Plan& P, T* sm, T* ws) {
  c.n = P.n; c.m = P.m; c.e = P.e;
  c.Qi = P.Qi_smem ? sm + P.off_Qi : ws + P.w_Qi; c.ldQi = P.ldQi;
  c.Tm = P.T_smem ? sm + P.off_T : ws + P.w_T;    c.ldT = P.ldT;
  c.R = ws + P.w_R; c.X = ws + P.w_X; c.XA = ws + P.w_XA; c.S11 = ws + P.w_S11;
  c.Vm = ws + P.w_V; c.W = ws + P.w_W;
  c.v.carve(sm + P.off_vec, P.n, P.m, P.e, P.nt);
}

template <typename T>
__device__ void bind_scene(SceneCtx<T>& c, const Plan& P, T* sm, const T* Q, const T* G, const T* A, const T* F) {
  const int n = P.n, m = P.m;
  c.Q = Q; c.A = A; c.F = F;
  if (P.G_smem) {
    T* Gs = sm + P.off_G;
    for (int t = threadIdx.x; t < m * n; t += blockDim.x) { int i = t / n, j = t - i * n; Gs[(size_t)i * P.ldG + j] = G[t]; }
    c.G = Gs; c.ldG = P.ldG;
  } else {
    c.G = G; c.ldG = n;
  }
  __syncthreads();
}

// ------------------------------------------------------------------ forward (pdipm.py:49-179)
template <typename T>
__global__ void __launch_bounds__(512) lcp_forward_kernel(const FwdArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  __shared__ int flag;
  __shared__ int lu_flag_s;
  const Plan& P = a.P;
  const int n = P.n, m = P.m, e = P.e, tid = threadIdx.x, NT = blockDim.x;
  SceneCtx<T> c;
  setup_ctx(c, P, sm, a.ws + (size_t)blockIdx.x * P.ws_per_cta);
  c.lu_flag = &lu_flag_s;
  Vecs<T>& v = c.v;
  const T NANV = nan("");

  for (int sc = blockIdx.x; sc < a.B; sc += gridDim.x) {
    const T* p = a.p + (size_t)sc * n;
    const T* h = a.h + (size_t)sc * m;
    const T* b = e > 0 ? a.b + (size_t)sc * e : nullptr;
    T* o_x = a.zhat + (size_t)sc * n;
    T* o_z = a.lam + (size_t)sc * m;
    T* o_s = a.slack + (size_t)sc * m;
    T* o_y = e > 0 ? a.nu + (size_t)sc * e : nullptr;
    bind_scene(c, P, sm, a.Q + (size_t)sc * n * n, a.G + (size_t)sc * m * n,
               e > 0 ? a.A + (size_t)sc * e * n : nullptr, a.F + (size_t)sc * m * m);

    if (!prefactor(c, &flag)) {
      for (int i = tid; i < n; i += NT) o_x[i] = NANV;
      for (int i = tid; i < m; i += NT) { o_z[i] = NANV; o_s[i] = NANV; }
      for (int i = tid; i < e; i += NT) o_y[i] = NANV;
      if (tid == 0) { a.status[sc] = -1; a.iters[sc] = 0; if (a.resid) a.resid[sc] = NANV; }
      __syncthreads();
      continue;
    }

    // ---- initial point: d = 1, rhs (p, 0, -h, -b)                 :58-63
    for (int i = tid; i < m; i += NT) { v.d[i] = T(1); v.rs2[i] = T(0); v.rz[i] = -h[i]; }
    for (int i = tid; i < n; i += NT) v.rx[i] = p[i];
    for (int i = tid; i < e; i += NT) v.ry[i] = -b[i];
    __syncthreads();
    factor_kkt(c, v.d);
    solve_kkt(c, v.d, v.rx, v.rs2, v.rz, e > 0 ? v.ry : nullptr, v.x, v.s, v.z, v.y);
    {   // shift s and z to >= 1 where the row minimum is <= 0       :65-75
      T mn[2] = {INFINITY, INFINITY};
      for (int i = tid; i < m; i += NT) { mn[0] = nan_min(mn[0], v.s[i]); mn[1] = nan_min(mn[1], v.z[i]); }
      block_reduce<T, 2>(mn, OpMin(), (T)INFINITY, v.red);
      for (int i = tid; i < m; i += NT) {
        if (mn[0] <= T(0)) v.s[i] -= mn[0] - T(1);
        if (mn[1] <= T(0)) v.z[i] -= mn[1] - T(1);
      }
      __syncthreads();
    }

    T best = NANV;
    bool have_best = false;
    int not_improved = 0, status = 0, it = 0;
    for (it = 0; it < a.max_iter; ++it) {
      // ---- residuals                                              :82-96
      // rx = A^T y + G^T z + Q x + p
      gemv_cols(c.G, c.ldG, m, n, v.z, v.scratch, [&](int j, T acc) { v.rx[j] = acc; });
      if (e > 0) gemv_cols(c.A, n, e, n, v.y, v.scratch, [&](int j, T acc) { v.rx[j] = acc + v.rx[j]; });
      gemv_rows(c.Q, n, n, n, v.x, [&](int i, T acc) { v.rx[i] = v.rx[i] + acc + p[i]; });
      // rz = G x + s - h - F z
      gemv_rows(c.G, c.ldG, m, n, v.x, [&](int i, T acc) { v.rz[i] = acc + v.s[i] - h[i]; });
      gemv_rows(c.F, m, m, m, v.z, [&](int i, T acc) { v.rz[i] -= acc; });
      if (e > 0) gemv_rows(c.A, n, e, n, v.x, [&](int i, T acc) { v.ry[i] = acc - b[i]; });
      T q[4] = {0, 0, 0, 0};                                      // s.z, |rz|^2, |ry|^2, |rx|^2
      for (int i = tid; i < m; i += NT) { q[0] += v.s[i] * v.z[i]; q[1] += v.rz[i] * v.rz[i]; }
      for (int i = tid; i < e; i += NT) q[2] += v.ry[i] * v.ry[i];
      for (int i = tid; i < n; i += NT) q[3] += v.rx[i] * v.rx[i];
      block_reduce<T, 4>(q, OpSum(), T(0), v.red);
      const T sz = q[0];
      const T mu = fabs(sz / T(m));                               // :91
      const T resid = (e > 0 ? sqrt(q[2]) : T(0)) + sqrt(q[1]) + sqrt(q[3]) + T(m) * mu;   // :92-96

      // ---- d = z/s, factor                                        :98-100
      for (int i = tid; i < m; i += NT) v.d[i] = v.z[i] / v.s[i];
      __syncthreads();
      factor_kkt(c, v.d);

      // ---- best iterate / termination (per scene)                 :107-136
      bool improved;
      if (!have_best) { improved = true; have_best = true; not_improved = 0; }
      else { improved = resid < best; not_improved = improved ? 0 : not_improved + 1; }
      if (improved) {
        best = resid;
        for (int i = tid; i < n; i += NT) o_x[i] = v.x[i];
        for (int i = tid; i < m; i += NT) { o_z[i] = v.z[i]; o_s[i] = v.s[i]; }
        for (int i = tid; i < e; i += NT) o_y[i] = v.y[i];
      }
      if (not_improved == a.not_improved_lim) { status = 1; ++it; break; }
      if (best < a.eps) { status = 2; ++it; break; }
      if (mu > T(1e100)) { status = 3; ++it; break; }

      // ---- affine direction                                       :138-139   (rs = z)
      solve_kkt(c, v.d, v.rx, v.z, v.rz, e > 0 ? v.ry : nullptr, v.dxa, v.dsa, v.dza, v.dya);
      T stz, sts;
      get_steps(v.z, v.dza, v.s, v.dsa, m, v.red, stz, sts);
      const T alpha_aff = nan_min(nan_min(stz, sts), T(1));       // :142-144
      T t3[1] = {0};
      for (int i = tid; i < m; i += NT) t3[0] += (v.s[i] + alpha_aff * v.dsa[i]) * (v.z[i] + alpha_aff * v.dza[i]);
      block_reduce<T, 1>(t3, OpSum(), T(0), v.red);
      const T ratio = t3[0] / sz;                                 // :146-150
      const T sig = ratio * ratio * ratio;
      // ---- corrector                                              :152-158
      const T musig = -mu * sig;
      for (int i = tid; i < m; i += NT) v.rs2[i] = (musig + v.dsa[i] * v.dza[i]) / v.s[i];
      __syncthreads();
      solve_kkt(c, v.d, (const T*)nullptr, v.rs2, (const T*)nullptr, (const T*)nullptr, v.dxc, v.dsc, v.dzc, v.dyc);
      for (int i = tid; i < n; i += NT) v.dxa[i] += v.dxc[i];    // :160-163
      for (int i = tid; i < m; i += NT) { v.dsa[i] += v.dsc[i]; v.dza[i] += v.dzc[i]; }
      for (int i = tid; i < e; i += NT) v.dya[i] += v.dyc[i];
      __syncthreads();
      get_steps(v.z, v.dza, v.s, v.dsa, m, v.red, stz, sts);
      const T alpha = nan_min(T(0.999) * nan_min(stz, sts), T(1));   // :164-166
      for (int i = tid; i < n; i += NT) v.x[i] += alpha * v.dxa[i];  // :171-174
      for (int i = tid; i < m; i += NT) { v.s[i] += alpha * v.dsa[i]; v.z[i] += alpha * v.dza[i]; }
      for (int i = tid; i < e; i += NT) v.y[i] += alpha * v.dya[i];
      __syncthreads();
    }
    if (tid == 0) { a.status[sc] = status; a.iters[sc] = it; if (a.resid) a.resid[sc] = best; }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ backward (lcp.py:37-64)
template <typename T>
struct BwdArgs {
  Plan P;
  int B;
  const T *Q, *G, *A, *F;
  const T *zhat, *nu, *lam, *slack, *g;
  T *dQ, *dp, *dG, *dh, *dA, *db, *dF;
  unsigned flags;
  T* ws;
};

template <typename T>
__global__ void __launch_bounds__(512) lcp_backward_kernel(const BwdArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  __shared__ int flag;
  __shared__ int lu_flag_s;
  const Plan& P = a.P;
  const int n = P.n, m = P.m, e = P.e, tid = threadIdx.x, NT = blockDim.x;
  SceneCtx<T> c;
  setup_ctx(c, P, sm, a.ws + (size_t)blockIdx.x * P.ws_per_cta);
  c.lu_flag = &lu_flag_s;
  Vecs<T>& v = c.v;

  for (int sc = blockIdx.x; sc < a.B; sc += gridDim.x) {
    bind_scene(c, P, sm, a.Q + (size_t)sc * n * n, a.G + (size_t)sc * m * n,
               e > 0 ? a.A + (size_t)sc * e * n : nullptr, a.F + (size_t)sc * m * m);
    const T* zh = a.zhat + (size_t)sc * n;
    const T* lam = a.lam + (size_t)sc * m;
    const T* slk = a.slack + (size_t)sc * m;
    const T* nu = e > 0 ? a.nu + (size_t)sc * e : nullptr;
    prefactor(c, &flag);     // singular Q was already reported by the forward pass
    // stage saved vectors in shared memory
    for (int i = tid; i < n; i += NT) { v.x[i] = zh[i]; v.rx[i] = a.g[(size_t)sc * n + i]; }
    for (int i = tid; i < m; i += NT) { v.z[i] = lam[i]; v.s[i] = slk[i]; v.d[i] = lam[i] / slk[i]; v.rs2[i] = T(0); }   // :44
    for (int i = tid; i < e; i += NT) v.y[i] = nu[i];
    __syncthreads();
    factor_kkt(c, v.d);                                            // :46
    // :47-50  solve_kkt(rx = dl_dzhat, rs = 0, rz = 0, ry = 0)
    solve_kkt(c, v.d, v.rx, v.rs2, (const T*)nullptr, (const T*)nullptr, v.dxa, v.dsa, v.dza, v.dya);
    const T* dx = v.dxa; const T* dlam = v.dza; const T* dnu = v.dya;
    if (a.dp) for (int i = tid; i < n; i += NT) a.dp[(size_t)sc * n + i] = dx[i];                       // :52
    if (a.dh) for (int i = tid; i < m; i += NT) a.dh[(size_t)sc * m + i] = -dlam[i];                    // :55
    if (a.db && e > 0) for (int i = tid; i < e; i += NT) a.db[(size_t)sc * e + i] = -dnu[i];            // :58
    if (a.dG) {                                                    // :53  dlam (x) zhat + lam (x) dx
      T* o = a.dG + (size_t)sc * m * n;
      for (int t = tid; t < m * n; t += NT) { int i = t / n, j = t - i * n; o[t] = dlam[i] * v.x[j] + v.z[i] * dx[j]; }
    }
    if (a.dF) {                                                    // :54  -dlam (x) lam
      T* o = a.dF + (size_t)sc * m * m;
      for (int t = tid; t < m * m; t += NT) { int i = t / m, j = t - i * m; o[t] = -(dlam[i] * v.z[j]); }
    }
    if (a.dA && e > 0) {                                           // :57
      T* o = a.dA + (size_t)sc * e * n;
      for (int t = tid; t < e * n; t += NT) { int i = t / n, j = t - i * n; o[t] = dnu[i] * v.x[j] + v.y[i] * dx[j]; }
    }
    if (a.dQ) {                                                    // :61
      T* o = a.dQ + (size_t)sc * n * n;
      for (int t = tid; t < n * n; t += NT) { int i = t / n, j = t - i * n; o[t] = T(0.5) * (dx[i] * v.x[j] + v.x[i] * dx[j]); }
    }
    __syncthreads();
  }
}

}  // namespace lcpb200
