// lcp_condensed.cuh -- "condensed KKT" PDIPM kernels: the structured fast path of the solver.
//
// Same algorithm as lcp_solver.cuh (pdipm.py:49-186, :325-454; lcp.py:22-64) but the Newton system
// of every iteration is reduced to the PRIMAL side instead of the dual side. The reference
// eliminates dx and factors the m x m matrix T = G Q^-1 G^T + F + diag(s/z) (pdipm.py:414-454); every
// LCP the engine builds (engines.py:50-76, :80-116) has two properties this file exploits:
//   (1) F + diag(s/z) =: M is block diagonal after a row permutation -- the connected components
//       of F's sparsity graph are the per-contact blocks {normal, friction dirs, gamma}
//       (engines.py:66-72: E, mu, -E^T), size 2 + fd; for post-stabilisation F = 0 (size 1);
//   (2) a component's rows of G touch two bodies only (world.py:172-211): <= 6 columns.
// Eliminating dz = M^-1 (G dx + rz - rs/d) instead gives the n x n (+ e equality rows) system
//       [ Q + G^T M^-1 G   A^T ] [dx]   [ -rx - G^T M^-1 (rz - rs/d) ]
//       [ A                0   ] [dy] = [ -ry                        ]
// n = 96 instead of m = 256 at BASELINE config 3: 19x fewer factorisation flops, and the matrix is
// small enough for two scenes per SM. The same linear system is solved, so the iterates are the
// reference's up to round-off. Round-off is the catch: M^-1 has entries z/s up to 1e13 (fp32 runs)
// on active constraints and K = Q + G^T M^-1 G inherits them, so K is formed, factored and solved
// in FP64 whatever the I/O dtype (B200: FP64 FMA at half the FP32 rate). Measured on seeded
// config-3 scenes (DESIGN.md "Parity"): fp32 I/O + fp64 K is CLOSER to the fp64 reference than the
// reference's own fp32 run; an fp32 K is 20x worse than it; explicit inverses of 16x16 (even 4x4)
// diagonal blocks of the factors lose 4-6 digits, so the triangular solves are exact
// substitutions.
//
// Scenes whose inputs do not have this structure (dense F or G, non-diagonal Q, n + e > 128) are
// flagged and solved by the dual-form kernels of lcp_solver.cuh (the structure test is per scene,
// on the device, on the dense tensors the API receives).
//
// Execution: one CTA of 256 threads per scene, persistent grid, two CTAs per SM at n + e <= 96.
//   * structure (once per scene): ballot scans compact F and G rows, label propagation finds F's
//     components, rows are regrouped per component: Gd (dense cs x <= 8 block + its column list),
//     Fd (cs x cs), per-column lists of (component, slot) for G^T w and the assembly of K;
//   * per iteration: W_c = (Fd_c + diag(s/z))^-1 per component (fp64, Gauss-Jordan with partial
//     pivoting in registers), K assembled in shared memory (column major), moved to REGISTERS in a
//     2-D cyclic layout (thread (ti,tj) owns K[16r+ti][16c+tj]), LU without pivoting, two pivots
//     per CTA barrier (pivot rows/columns broadcast through shared memory, every thread redoes the
//     2x2 pivot arithmetic), factors written back transposed with U's columns pre-scaled by 1/u_kk;
//   * solves: ONE warp, exact forward/back substitution, 4 pivots per shuffle round (the 4x4
//     diagonal piece is redone by every lane, so the dependent chain per pivot is ~14 cycles).
#pragma once
#include "lcp_device.cuh"

namespace lcpb200 {
namespace cnd {

constexpr int NT = 256;          // threads per CTA (16 x 16 grid over K)
constexpr int UC = 8;            // max distinct columns of G over the rows of one component
constexpr int CSMAX = 6;         // max rows per component of F
constexpr int KS = 8;            // scan slots per row of F / G
constexpr int LMAX = 16;         // max (component, slot) entries per column of G
constexpr int STATUS_UNSUPPORTED = -100;   // internal: scene left to the dual-form kernel

__device__ __forceinline__ double rcp64(double x) {
  // reciprocal off the slow IEEE-division path: fp32 seed + 3 Newton steps (~1 ulp); exact
  // division outside the seed's range (also keeps inf / nan semantics)
  float xf = (float)x, rf;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(xf));
  if (!(fabsf(rf) < INFINITY) || rf == 0.0f) return 1.0 / x;
  double r = (double)rf;
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}

// ------------------------------------------------------------------ launch plan / shared layout
struct CPlan {
  int ok;                     // 0: the condensed path is not available for this (dtype, n, m, e)
  int n, m, e, N, NP, NS;
  int pcap;                   // capacity of component-sorted row positions
  int wcap;                   // capacity (elements) of W and Fd
  int smem_bytes;
  int ctas_per_sm;
};

template <typename T>
struct CSmem {
  double* K;                  // NP*NP (column major) | structure scratch | LU broadcast buffers
  double* W;                  // wcap: M_c^-1, component-major, stride cs*cs
  double* scr;                // pcap
  double* bx;                 // NP: right-hand side / solution of the condensed system
  double* rdiag;              // NP: 1 / u_kk
  T *Fd, *Gd, *As;            // wcap | UC*pcap (slot-major) | e*n
  T *x, *rx, *dx, *qd, *y, *ry, *dy;
  T *s, *z, *d, *rz, *ds, *dz, *rs2;
  T* red;                     // 128
  unsigned short *rows, *posof, *clist;       // pcap | m | LMAX*n
  unsigned char *ccols, *ncols, *clcnt;       // UC*pcap (slot-major, per component) | pcap | n
  int* misc;                  // 16 ints
  __host__ __device__ static size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
  __host__ __device__ static size_t scratch_bytes(int n, int m) {
    // structure-build scratch: Fi,Gi (u16 KS*m each), Fv,Gv (T KS*m each), 5 int arrays of m
    return al16((size_t)2 * KS * m * 2) + al16((size_t)2 * KS * m * sizeof(T)) + al16((size_t)5 * m * 4) + 64;
  }
  __host__ __device__ size_t carve(char* base, const CPlan& P) {
    size_t o = 0;
    const int n = P.n, m = P.m, e = P.e, NP = P.NP, pcap = P.pcap, wcap = P.wcap;
#define CND_TAKE(ptr, type, cnt) do { ptr = reinterpret_cast<type*>(base + o); o += al16((size_t)(cnt) * sizeof(type)); } while (0)
    size_t kb = (size_t)NP * NP * 8;
    const size_t sb = scratch_bytes(n, m);
    const size_t lb = (size_t)8 * NP * 8;                       // LU broadcast buffers (2 x 4 x NP doubles)
    if (sb > kb) kb = sb;
    if (lb > kb) kb = lb;
    K = reinterpret_cast<double*>(base + o); o += al16(kb);
    CND_TAKE(W, double, wcap);
    CND_TAKE(scr, double, pcap);
    CND_TAKE(bx, double, NP);
    CND_TAKE(rdiag, double, NP);
    CND_TAKE(Fd, T, wcap);
    CND_TAKE(Gd, T, UC * pcap);
    CND_TAKE(As, T, e * n);
    CND_TAKE(x, T, n); CND_TAKE(rx, T, n); CND_TAKE(dx, T, n); CND_TAKE(qd, T, n);
    CND_TAKE(y, T, e); CND_TAKE(ry, T, e); CND_TAKE(dy, T, e);
    CND_TAKE(s, T, m); CND_TAKE(z, T, m); CND_TAKE(d, T, m); CND_TAKE(rz, T, m);
    CND_TAKE(ds, T, m); CND_TAKE(dz, T, m); CND_TAKE(rs2, T, m);
    CND_TAKE(red, T, 128);
    CND_TAKE(rows, unsigned short, pcap);
    CND_TAKE(posof, unsigned short, m);
    CND_TAKE(clist, unsigned short, LMAX * n);
    CND_TAKE(ccols, unsigned char, UC * pcap);
    CND_TAKE(ncols, unsigned char, pcap);
    CND_TAKE(clcnt, unsigned char, n);
    CND_TAKE(misc, int, 16);
#undef CND_TAKE
    return o;
  }
};

// per-scene structure facts (uniform across the CTA, kept in registers)
struct Struct {
  int ncomp, cs;              // components, rows per component (uniform stride; short ones padded)
};

extern __shared__ __align__(16) unsigned char cnd_smem[];

// ------------------------------------------------------------------ structure detection
// Scans one dense row-major matrix (rows x cols) into KS-slot row lists (values + column indices),
// one warp per row, ordered (deterministic). Returns non-zero (CTA-wide) if a row has > KS entries.
template <typename T>
__device__ __forceinline__ int scan_rows(const T* __restrict__ A, int rows, int cols, T* vals, unsigned short* idx,
                                         int* cnt) {
  constexpr int CH = 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = NT >> 5;
  int bad = 0;
  for (int r = warp; r < rows; r += nw) {
    const T* row = A + (size_t)r * cols;
    int c = 0;
    for (int j0 = 0; j0 < cols; j0 += 32 * CH) {
      T v[CH];
#pragma unroll
      for (int q = 0; q < CH; ++q) { const int j = j0 + q * 32 + lane; v[q] = j < cols ? row[j] : T(0); }
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const bool nz = v[q] != T(0);
        const unsigned mask = __ballot_sync(FULL, nz);
        const int pos = c + __popc(mask & ((1u << lane) - 1u));
        if (nz && pos < KS) { vals[pos * rows + r] = v[q]; idx[pos * rows + r] = (unsigned short)(j0 + q * 32 + lane); }
        c += __popc(mask);
      }
    }
    if (c > KS) bad = 1;
    if (lane == 0) cnt[r] = c;
  }
  return bad;
}

// Builds the per-scene structure in shared memory. Returns false (uniformly) when the scene does
// not have the structure this path needs. *singular is set when Q has a zero / non-finite diagonal
// entry (the reference fails its LU of Q, pdipm.py:361-368).
template <typename T>
__device__ __noinline__ bool build_structure(const CPlan& P, CSmem<T>& S, Struct& st, const T* __restrict__ Q,
                                             const T* __restrict__ G, const T* __restrict__ A,
                                             const T* __restrict__ F, int* singular) {
  const int n = P.n, m = P.m, e = P.e, tid = threadIdx.x;
  // scratch carved from the K region
  char* sb = reinterpret_cast<char*>(S.K);
  unsigned short* Fi = reinterpret_cast<unsigned short*>(sb);
  unsigned short* Gi = Fi + (size_t)KS * m;
  size_t o = CSmem<T>::al16((size_t)2 * KS * m * 2);
  T* Fv = reinterpret_cast<T*>(sb + o);
  T* Gv = Fv + (size_t)KS * m;
  o += CSmem<T>::al16((size_t)2 * KS * m * sizeof(T));
  int* Fcnt = reinterpret_cast<int*>(sb + o);
  int* Gcnt = Fcnt + m;
  int* label = Gcnt + m;
  int* cidx = label + m;       // component index of a root row
  int* csize = cidx + m;       // rows per component (indexed by component)
  int* flags = S.misc;         // [0] bad, [1] changed, [2] ncomp, [3] csmax, [4] singular

  if (tid < 8) flags[tid] = 0;
  // ---- Q must be diagonal (every mass matrix world.py:57-61 builds)
  int bad = 0;
  for (int t = tid; t < n * n; t += NT) {
    const int i = t / n, j = t - i * n;
    const T q = Q[t];
    if (i == j) {
      S.qd[i] = q;
      if (!(q != T(0) && isfinite((double)q))) bad |= 2;
    } else if (q != T(0)) bad |= 1;
  }
  bad |= scan_rows<T>(F, m, m, Fv, Fi, Fcnt) ? 1 : 0;
  bad |= scan_rows<T>(G, m, n, Gv, Gi, Gcnt) ? 1 : 0;
  for (int t = tid; t < e * n; t += NT) S.As[t] = A[t];
  for (int i = tid; i < m; i += NT) label[i] = i;
  const int anybad = __syncthreads_or(bad);
  if (anybad & 2) { *singular = 1; return false; }
  if (anybad & 1) return false;

  // ---- components of F's sparsity graph: label propagation to the minimum row index
  for (int pass = 0; pass < 64; ++pass) {
    int changed = 0;
    for (int i = tid; i < m; i += NT) {
      const int c = Fcnt[i];
      int li = ((volatile int*)label)[i];
      for (int k = 0; k < c; ++k) {
        const int j = Fi[k * m + i];
        const int lj = ((volatile int*)label)[j];
        if (lj < li) { li = lj; changed = 1; }
        else if (li < lj) { atomicMin(&label[j], li); changed = 1; }
      }
      atomicMin(&label[i], li);
    }
    if (!__syncthreads_or(changed)) break;
    if (pass == 63) return false;
  }
  // ---- component index = rank of the root among roots (ordered by row), slot = rank inside it
  for (int i = tid; i < m; i += NT) csize[i] = 0;
  __syncthreads();
  {
    // exclusive count of roots before row i (m is a few hundred: a direct count per thread)
    for (int i = tid; i < m; i += NT) {
      if (label[i] == i) {
        int c = 0;
        for (int j = 0; j < i; ++j) c += (label[j] == j);
        cidx[i] = c;
        atomicMax(&flags[2], c + 1);
      }
    }
  }
  __syncthreads();
  const int ncomp = flags[2];
  int myslot[4];                      // up to 4 rows per thread (m <= 1024)
  if (m > 4 * NT) return false;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = tid + q * NT;
    myslot[q] = 0;
    if (i < m) {
      const int l = label[i];
      int r = 0;
      for (int j = l; j < i; ++j) r += (label[j] == l);
      myslot[q] = r;
      atomicMax(&csize[cidx[l]], r + 1);
    }
  }
  __syncthreads();
  {
    int mx = 0;
    for (int c = tid; c < ncomp; c += NT) mx = max(mx, csize[c]);
    atomicMax(&flags[3], mx);
  }
  __syncthreads();
  const int cs = flags[3];
  if (cs > CSMAX || ncomp * cs > P.pcap || ncomp * cs * cs > P.wcap || ncomp > 8191) return false;
  st.ncomp = ncomp; st.cs = cs;
  const int npos = ncomp * cs;
  // ---- rows <-> positions, Fd, union columns, Gd
  for (int p = tid; p < npos; p += NT) { S.rows[p] = 0xFFFF; S.ncols[p] = 0; }
  for (int t = tid; t < ncomp * cs * cs; t += NT) S.Fd[t] = T(0);
  for (int t = tid; t < UC * P.pcap; t += NT) { S.Gd[t] = T(0); S.ccols[t] = 0; }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = tid + q * NT;
    if (i < m) {
      const int pos = cidx[label[i]] * cs + myslot[q];
      S.rows[pos] = (unsigned short)i;
      S.posof[i] = (unsigned short)pos;
    }
  }
  __syncthreads();
  for (int i = tid; i < m; i += NT) {
    const int pos = S.posof[i], c = pos / cs, r = pos - c * cs;
    const int fc = Fcnt[i];
    for (int k = 0; k < fc; ++k) {
      const int pj = S.posof[Fi[k * m + i]];
      S.Fd[(size_t)c * cs * cs + r * cs + (pj - c * cs)] = Fv[k * m + i];
    }
  }
  int bad2 = 0;
  for (int c = tid; c < ncomp; c += NT) {           // sorted union of the columns of the component's rows
    int cnt = 0;
    for (int r = 0; r < cs; ++r) {
      const int i = S.rows[c * cs + r];
      if (i == 0xFFFF) continue;
      const int gc = Gcnt[i];
      for (int k = 0; k < gc; ++k) {
        const int col = Gi[k * m + i];
        int p = 0;
        while (p < cnt && S.ccols[p * P.pcap + c] < col) ++p;
        if (p < cnt && S.ccols[p * P.pcap + c] == col) continue;
        if (cnt == UC) { bad2 = 1; break; }
        for (int t = cnt; t > p; --t) S.ccols[t * P.pcap + c] = S.ccols[(t - 1) * P.pcap + c];
        S.ccols[p * P.pcap + c] = (unsigned char)col;
        ++cnt;
      }
    }
    S.ncols[c] = (unsigned char)cnt;
  }
  if (__syncthreads_or(bad2)) return false;
  for (int i = tid; i < m; i += NT) {
    const int pos = S.posof[i], c = pos / cs;
    const int gc = Gcnt[i], nc = S.ncols[c];
    for (int k = 0; k < gc; ++k) {
      const int col = Gi[k * m + i];
      int p = 0;
      while (p < nc && S.ccols[p * P.pcap + c] != col) ++p;
      S.Gd[(size_t)p * P.pcap + pos] = Gv[k * m + i];
    }
  }
  // ---- per-column lists of (component, slot): G^T w and the assembly of K gather through them
  int bad3 = 0;
  for (int a = tid; a < n; a += NT) {
    int cnt = 0;
    for (int c = 0; c < ncomp; ++c) {
      const int nc = S.ncols[c];
      for (int p = 0; p < nc; ++p)
        if (S.ccols[p * P.pcap + c] == a) {
          if (cnt < LMAX) S.clist[cnt * n + a] = (unsigned short)(c * 8 + p);
          ++cnt;
        }
    }
    if (cnt > LMAX) bad3 = 1;
    S.clcnt[a] = (unsigned char)min(cnt, LMAX);
  }
  if (__syncthreads_or(bad3)) return false;
  return true;
}

// ------------------------------------------------------------------ W_c = (Fd_c + diag(1/d))^-1
// One thread per component, Gauss-Jordan with partial pivoting on [M | I] held in registers
// (static indices only: row interchanges are conditional swaps).
template <typename T, int CS>
__device__ __forceinline__ void comp_inverse(const CPlan& P, CSmem<T>& S, const Struct& st) {
  for (int c = threadIdx.x; c < st.ncomp; c += NT) {
    double M[CS][CS], V[CS][CS];
#pragma unroll
    for (int r = 0; r < CS; ++r) {
      const int i = S.rows[c * CS + r];
#pragma unroll
      for (int q = 0; q < CS; ++q) {
        M[r][q] = (double)S.Fd[(size_t)c * CS * CS + r * CS + q];
        V[r][q] = (r == q) ? 1.0 : 0.0;
      }
      M[r][r] += (i == 0xFFFF) ? 1.0 : 1.0 / (double)S.d[i];
    }
#pragma unroll
    for (int k = 0; k < CS; ++k) {
#pragma unroll
      for (int i = k + 1; i < CS; ++i) {           // bring the largest |M[i][k]|, i >= k, to row k
        const bool sw = fabs(M[i][k]) > fabs(M[k][k]);
#pragma unroll
        for (int q = 0; q < CS; ++q) {
          const double a = M[k][q], b = M[i][q]; M[k][q] = sw ? b : a; M[i][q] = sw ? a : b;
          const double u = V[k][q], w = V[i][q]; V[k][q] = sw ? w : u; V[i][q] = sw ? u : w;
        }
      }
      const double r = 1.0 / M[k][k];
#pragma unroll
      for (int q = 0; q < CS; ++q) { M[k][q] *= r; V[k][q] *= r; }
#pragma unroll
      for (int i = 0; i < CS; ++i) {
        if (i == k) continue;
        const double f = M[i][k];
#pragma unroll
        for (int q = 0; q < CS; ++q) { M[i][q] = fma(-f, M[k][q], M[i][q]); V[i][q] = fma(-f, V[k][q], V[i][q]); }
      }
    }
#pragma unroll
    for (int r = 0; r < CS; ++r)
#pragma unroll
      for (int q = 0; q < CS; ++q) S.W[(size_t)c * CS * CS + r * CS + q] = V[r][q];
  }
}

// out_c = W_c * in_c for every component, in place in S.scr (one thread per component)
template <typename T, int CS>
__device__ __forceinline__ void comp_apply(CSmem<T>& S, const Struct& st) {
  for (int c = threadIdx.x; c < st.ncomp; c += NT) {
    double v[CS], w[CS];
#pragma unroll
    for (int r = 0; r < CS; ++r) v[r] = S.scr[c * CS + r];
#pragma unroll
    for (int r = 0; r < CS; ++r) {
      double a = 0;
#pragma unroll
      for (int q = 0; q < CS; ++q) a = fma(S.W[(size_t)c * CS * CS + r * CS + q], v[q], a);
      w[r] = a;
    }
#pragma unroll
    for (int r = 0; r < CS; ++r) S.scr[c * CS + r] = w[r];
  }
}

// ------------------------------------------------------------------ K (column major, fp64)
// Kbar = [[Q + G^T W G, A^T], [A, 0]] padded with an identity block to NP. Row owner a gathers the
// contributions of the components that touch column a (deterministic order).
template <typename T, int CS>
__device__ __forceinline__ void assemble_K(const CPlan& P, CSmem<T>& S, const Struct& st) {
  const int n = P.n, e = P.e, N = P.N, NP = P.NP, tid = threadIdx.x, pcap = P.pcap;
  double* K = S.K;
  for (int t = tid; t < NP * NP; t += NT) {
    const int j = t / NP, i = t - j * NP;           // K[i][j] at K[j*NP + i]
    double v = 0.0;
    if (i < n && j < n) v = (i == j) ? (double)S.qd[i] : 0.0;
    else if (i < n && j < N) v = (double)S.As[(j - n) * n + i];
    else if (j < n && i < N) v = (double)S.As[(i - n) * n + j];
    else if (i >= N && i == j) v = 1.0;
    K[t] = v;
  }
  __syncthreads();
  for (int a = tid; a < n; a += NT) {
    const int cnt = S.clcnt[a];
    for (int l = 0; l < cnt; ++l) {
      const int cp = S.clist[l * n + a], c = cp >> 3, p = cp & 7;
      double u[CS];
#pragma unroll
      for (int q = 0; q < CS; ++q) u[q] = 0.0;
#pragma unroll
      for (int r = 0; r < CS; ++r) {
        const double g = (double)S.Gd[(size_t)p * pcap + c * CS + r];
#pragma unroll
        for (int q = 0; q < CS; ++q) u[q] = fma(g, S.W[(size_t)c * CS * CS + r * CS + q], u[q]);
      }
      const int nc = S.ncols[c];
      for (int q2 = 0; q2 < nc; ++q2) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < CS; ++q) acc = fma(u[q], (double)S.Gd[(size_t)q2 * pcap + c * CS + q], acc);
        K[(size_t)S.ccols[q2 * pcap + c] * NP + a] += acc;
      }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------ LU in registers
// Thread (ti = tid & 15, tj = tid >> 4) owns a[r][c] = K[16r + ti][16c + tj]. One phase = the 16
// pivots of block B0, two per barrier. Pivot rows / columns travel through `buf` (double buffered:
// [2][4][NP]); every thread redoes the 2x2 pivot arithmetic. No pivoting: K + its border is
// quasi-definite (Q > 0, symmetric part of W_c is what the reference's own pivot-free GPU path
// relies on); zero pivots produce inf/nan exactly like the reference's LU would.
template <int NS, int B0>
__device__ __forceinline__ void lu_phase(double (&a)[NS][NS], double* buf, double* rdiag, int NP, int& step) {
  const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
#pragma unroll 1
  for (int kk = 0; kk < 16; kk += 2) {
    const int k = 16 * B0 + kk;
    double* u0 = buf + (size_t)(step & 1) * 4 * NP;
    double* u1 = u0 + NP;
    double* c0 = u1 + NP;
    double* c1 = c0 + NP;
    ++step;
    if (ti == kk) {
#pragma unroll
      for (int c = B0; c < NS; ++c) u0[16 * c + tj] = a[B0][c];
    }
    if (ti == kk + 1) {
#pragma unroll
      for (int c = B0; c < NS; ++c) u1[16 * c + tj] = a[B0][c];
    }
    if (tj == kk) {
#pragma unroll
      for (int r = B0; r < NS; ++r) c0[16 * r + ti] = a[r][B0];
    }
    if (tj == kk + 1) {
#pragma unroll
      for (int r = B0; r < NS; ++r) c1[16 * r + ti] = a[r][B0];
    }
    __syncthreads();
    const double p00 = u0[k], p01 = u0[k + 1], p10 = u1[k], p11 = u1[k + 1];
    const double r0 = rcp64(p00);
    const double l10 = p10 * r0;
    const double r1 = rcp64(fma(-l10, p01, p11));
    if (threadIdx.x == 0) { rdiag[k] = r0; rdiag[k + 1] = r1; }
    double w0[NS], w1[NS];                          // pivot rows restricted to my columns
#pragma unroll
    for (int c = B0; c < NS; ++c) {
      const int j = 16 * c + tj;
      const double x0 = u0[j];
      const double x1 = fma(-l10, x0, u1[j]);
      if (ti == kk + 1 && j > k) a[B0][c] = x1;     // row k+1 of U (its entry in column k is L, below)
      const bool live = (c > B0) || (tj > kk + 1);
      w0[c] = live ? x0 : 0.0;
      w1[c] = live ? x1 : 0.0;
    }
#pragma unroll
    for (int r = B0; r < NS; ++r) {
      const int i = 16 * r + ti;
      const double l0 = c0[i] * r0;
      const double l1 = fma(-l0, p01, c1[i]) * r1;
      const bool live = (r > B0) || (ti > kk + 1);
      if (tj == kk && i > k) a[r][B0] = l0;
      if (tj == kk + 1 && live) a[r][B0] = l1;
      const double m0 = live ? l0 : 0.0, m1 = live ? l1 : 0.0;
#pragma unroll
      for (int c = B0; c < NS; ++c) a[r][c] = fma(-m1, w1[c], fma(-m0, w0[c], a[r][c]));
    }
  }
}

template <int NS, int B0>
struct LuPhases {
  static __device__ __forceinline__ void run(double (&a)[NS][NS], double* buf, double* rdiag, int NP, int& step) {
    lu_phase<NS, B0>(a, buf, rdiag, NP, step);
    LuPhases<NS, B0 + 1>::run(a, buf, rdiag, NP, step);
  }
};
template <int NS>
struct LuPhases<NS, NS> {
  static __device__ __forceinline__ void run(double (&)[NS][NS], double*, double*, int, int&) {}
};

// K (shared, column major) -> registers -> LU -> factors back to K, transposed layout kept:
// S[j*NP + i] = L[i][j] (i > j), U[i][j] / u_jj (i < j); rdiag[j] = 1/u_jj.
template <int NS>
__device__ __noinline__ void factor_K(double* K, double* rdiag, int NP) {
  const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  double a[NS][NS];
#pragma unroll
  for (int c = 0; c < NS; ++c)
#pragma unroll
    for (int r = 0; r < NS; ++r) a[r][c] = K[(size_t)(16 * c + tj) * NP + 16 * r + ti];
  __syncthreads();                                   // K region becomes the broadcast buffer
  int step = 0;
  LuPhases<NS, 0>::run(a, K, rdiag, NP, step);
  __syncthreads();                                   // rdiag complete, broadcast buffers dead
#pragma unroll
  for (int c = 0; c < NS; ++c) {
    const int j = 16 * c + tj;
    const double rj = rdiag[j];
#pragma unroll
    for (int r = 0; r < NS; ++r) {
      const int i = 16 * r + ti;
      K[(size_t)j * NP + i] = (i < j) ? a[r][c] * rj : a[r][c];
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------ triangular solves (ONE warp)
// bx <- Kbar^-1 bx with the factors left by factor_K. Lane owns rows lane + 32q. Exact
// substitution order; 4 pivots per round: their values are broadcast by shuffles, the 4x4
// diagonal piece is solved redundantly by every lane, then each lane updates its rows.
template <int QN>
__device__ __noinline__ void solve_warp(const double* __restrict__ S, const double* __restrict__ rdiag,
                                           double* bx, int NP) {
  const int lane = threadIdx.x & 31;
  double y[QN];
#pragma unroll
  for (int q = 0; q < QN; ++q) y[q] = (lane + 32 * q < NP) ? bx[lane + 32 * q] : 0.0;
  // ---- forward: L y = b (unit lower)
#pragma unroll
  for (int q0 = 0; q0 < QN; ++q0) {
#pragma unroll 1
    for (int kk = 0; kk < 8; ++kk) {
      const int k = 32 * q0 + 4 * kk;
      if (k >= NP) break;
      const double* c0 = S + (size_t)k * NP;
      const double* c1 = c0 + NP;
      const double* c2 = c1 + NP;
      const double* c3 = c2 + NP;
      const double v0 = __shfl_sync(FULL, y[q0], 4 * kk), v1 = __shfl_sync(FULL, y[q0], 4 * kk + 1);
      const double v2 = __shfl_sync(FULL, y[q0], 4 * kk + 2), v3 = __shfl_sync(FULL, y[q0], 4 * kk + 3);
      const double y1 = fma(-c0[k + 1], v0, v1);
      const double y2 = fma(-c1[k + 2], y1, fma(-c0[k + 2], v0, v2));
      const double y3 = fma(-c2[k + 3], y2, fma(-c1[k + 3], y1, fma(-c0[k + 3], v0, v3)));
#pragma unroll
      for (int q = q0; q < QN; ++q) {
        const int i = lane + 32 * q;
        if (i > k + 3 && i < NP)
          y[q] = fma(-c3[i], y3, fma(-c2[i], y2, fma(-c1[i], y1, fma(-c0[i], v0, y[q]))));
      }
      const int w = lane - 4 * kk;
      if (w == 1) y[q0] = y1;
      if (w == 2) y[q0] = y2;
      if (w == 3) y[q0] = y3;
    }
  }
  // ---- backward: U x = y with columns of U pre-scaled by 1/u_kk (z_k = u_kk x_k)
#pragma unroll
  for (int q0 = QN - 1; q0 >= 0; --q0) {
#pragma unroll 1
    for (int kk = 7; kk >= 0; --kk) {
      const int k = 32 * q0 + 4 * kk;
      if (k >= NP) continue;
      const double* c0 = S + (size_t)k * NP;
      const double* c1 = c0 + NP;
      const double* c2 = c1 + NP;
      const double* c3 = c2 + NP;
      const double v0 = __shfl_sync(FULL, y[q0], 4 * kk), v1 = __shfl_sync(FULL, y[q0], 4 * kk + 1);
      const double v2 = __shfl_sync(FULL, y[q0], 4 * kk + 2), v3 = __shfl_sync(FULL, y[q0], 4 * kk + 3);
      const double z2 = fma(-c3[k + 2], v3, v2);
      const double z1 = fma(-c2[k + 1], z2, fma(-c3[k + 1], v3, v1));
      const double z0 = fma(-c1[k], z1, fma(-c2[k], z2, fma(-c3[k], v3, v0)));
#pragma unroll
      for (int q = 0; q <= q0; ++q) {
        const int i = lane + 32 * q;
        if (i < k) y[q] = fma(-c0[i], z0, fma(-c1[i], z1, fma(-c2[i], z2, fma(-c3[i], v3, y[q]))));
      }
      const int w = lane - 4 * kk;
      if (w == 0) y[q0] = z0;
      if (w == 1) y[q0] = z1;
      if (w == 2) y[q0] = z2;
    }
  }
#pragma unroll
  for (int q = 0; q < QN; ++q) {
    const int i = lane + 32 * q;
    if (i < NP) bx[i] = y[q] * rdiag[i];
  }
}

// ------------------------------------------------------------------ solve_kkt (pdipm.py:325-354)
// Inputs (nullptr == zero vector): rx[n], rs[m], rz[m], ry[e]; outputs dx[n], ds[m], dz[m], dy[e].
template <typename T, int CS, int QN>
__device__ __forceinline__ void solve_kkt(const CPlan& P, CSmem<T>& S, const Struct& st, const T* rx, const T* rs,
                                          const T* rz, const T* ry, T* dx, T* ds, T* dz, T* dy) {
  const int n = P.n, e = P.e, N = P.N, NP = P.NP, tid = threadIdx.x, pcap = P.pcap;
  const int npos = st.ncomp * CS;
  // t = rz - rs/d  -> v = W t
  for (int p = tid; p < npos; p += NT) {
    const int i = S.rows[p];
    double t = 0.0;
    if (i != 0xFFFF) t = (double)((rz ? rz[i] : T(0)) - rs[i] / S.d[i]);
    S.scr[p] = t;
  }
  __syncthreads();
  comp_apply<T, CS>(S, st);
  __syncthreads();
  for (int a = tid; a < NP; a += NT) {
    double acc = 0.0;
    if (a < n) {
      const int cnt = S.clcnt[a];
      for (int l = 0; l < cnt; ++l) {
        const int cp = S.clist[l * n + a], c = cp >> 3, p = cp & 7;
#pragma unroll
        for (int r = 0; r < CS; ++r) acc = fma((double)S.Gd[(size_t)p * pcap + c * CS + r], S.scr[c * CS + r], acc);
      }
      acc = -(double)(rx ? rx[a] : T(0)) - acc;
    } else if (a < N) {
      acc = -(double)(ry ? ry[a - n] : T(0));
    }
    S.bx[a] = acc;
  }
  __syncthreads();
  if (tid < 32) solve_warp<QN>(S.K, S.rdiag, S.bx, NP);
  __syncthreads();
  // dz = W (G dx + t)
  for (int p = tid; p < npos; p += NT) {
    const int i = S.rows[p];
    double t = 0.0;
    if (i != 0xFFFF) {
      const int c = p / CS, nc = S.ncols[c];
      t = (double)((rz ? rz[i] : T(0)) - rs[i] / S.d[i]);
      for (int q = 0; q < nc; ++q) t = fma((double)S.Gd[(size_t)q * pcap + p], S.bx[S.ccols[q * pcap + c]], t);
    }
    S.scr[p] = t;
  }
  __syncthreads();
  comp_apply<T, CS>(S, st);
  __syncthreads();
  for (int p = tid; p < npos; p += NT) {
    const int i = S.rows[p];
    if (i == 0xFFFF) continue;
    const T wz = (T)S.scr[p];
    const T rsi = rs[i];                                         // (dz may alias rs)
    dz[i] = wz;                                                  // :351
    ds[i] = (-rsi - wz) / S.d[i];                                // :347,350
  }
  for (int a = tid; a < n; a += NT) dx[a] = (T)S.bx[a];
  for (int k = tid; k < e; k += NT) dy[k] = (T)S.bx[n + k];
  __syncthreads();
}

// d is in S.d: W, K, LU
template <typename T, int NS, int CS>
__device__ __forceinline__ void factor_kkt(const CPlan& P, CSmem<T>& S, const Struct& st) {
  comp_inverse<T, CS>(P, S, st);
  __syncthreads();
  assemble_K<T, CS>(P, S, st);
  factor_K<NS>(S.K, S.rdiag, P.NP);
}

// ------------------------------------------------------------------ get_step (pdipm.py:182-186), per scene
template <typename T>
__device__ __forceinline__ void get_steps(const T* z, const T* dz, const T* s, const T* ds, int m, T* red, T& step_z,
                                          T& step_s) {
  const T NEG_INF = -INFINITY, POS_INF = INFINITY;
  T mx[2] = {NEG_INF, NEG_INF};
  for (int i = threadIdx.x; i < m; i += NT) {
    mx[0] = nan_max(mx[0], -z[i] / dz[i]);
    mx[1] = nan_max(mx[1], -s[i] / ds[i]);
  }
  block_reduce<T, 2>(mx, OpMax(), NEG_INF, red);
  const T fz = (mx[0] > T(1)) ? mx[0] : T(1);                   // python max(1.0, a.max()): NaN -> 1.0
  const T fs = (mx[1] > T(1)) ? mx[1] : T(1);
  T mn[2] = {POS_INF, POS_INF};
  for (int i = threadIdx.x; i < m; i += NT) {
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
struct CFwdArgs {
  CPlan P;
  int B;
  const T *Q, *p, *G, *h, *A, *b, *F;
  T *zhat, *nu, *lam, *slack, *resid;
  int *status, *iters;
  T eps;
  int not_improved_lim, max_iter;
};

template <typename T>
struct CBwdArgs {
  CPlan P;
  int B;
  const T *Q, *G, *A, *F;
  const T *zhat, *nu, *lam, *slack, *g;
  T *dQ, *dp, *dG, *dh, *dA, *db, *dF;
  int* done;                  // [B]: 1 = gradients written here, 0 = scene left to the dual-form kernel
};

// ------------------------------------------------------------------ forward (pdipm.py:49-179), one scene
template <typename T, int NS, int CS>
__device__ __forceinline__ void forward_scene(const CFwdArgs<T>& a, CSmem<T>& S, const Struct& st, int sc) {
  constexpr int QN = (NS + 1) / 2;
  const CPlan& P = a.P;
  const int n = P.n, m = P.m, e = P.e, tid = threadIdx.x, pcap = P.pcap;
  const T* p = a.p + (size_t)sc * n;
  const T* h = a.h + (size_t)sc * m;
  const T* b = e > 0 ? a.b + (size_t)sc * e : nullptr;
  T* o_x = a.zhat + (size_t)sc * n;
  T* o_z = a.lam + (size_t)sc * m;
  T* o_s = a.slack + (size_t)sc * m;
  T* o_y = e > 0 ? a.nu + (size_t)sc * e : nullptr;
  const T NANV = nan("");

  // ---- initial point: d = 1, rhs (p, 0, -h, -b)                 :58-63
  for (int i = tid; i < m; i += NT) { S.d[i] = T(1); S.rs2[i] = T(0); S.rz[i] = -h[i]; }
  for (int i = tid; i < n; i += NT) S.rx[i] = p[i];
  for (int i = tid; i < e; i += NT) S.ry[i] = -b[i];
  __syncthreads();
  factor_kkt<T, NS, CS>(P, S, st);
  solve_kkt<T, CS, QN>(P, S, st, S.rx, S.rs2, S.rz, e > 0 ? S.ry : nullptr, S.x, S.s, S.z, S.y);
  {   // shift s and z to >= 1 where the row minimum is <= 0       :65-75
    T mn[2] = {INFINITY, INFINITY};
    for (int i = tid; i < m; i += NT) { mn[0] = nan_min(mn[0], S.s[i]); mn[1] = nan_min(mn[1], S.z[i]); }
    block_reduce<T, 2>(mn, OpMin(), (T)INFINITY, S.red);
    for (int i = tid; i < m; i += NT) {
      if (mn[0] <= T(0)) S.s[i] -= mn[0] - T(1);
      if (mn[1] <= T(0)) S.z[i] -= mn[1] - T(1);
    }
    __syncthreads();
  }

  T best = NANV;
  bool have_best = false;
  int not_improved = 0, status = 0, it = 0;
  const int npos = st.ncomp * CS;
  for (it = 0; it < a.max_iter; ++it) {
    // ---- residuals                                              :82-96
    for (int c = tid; c < n; c += NT) {                            // rx = G^T z + Q x + p (+ A^T y)
      T acc = 0;
      const int cnt = S.clcnt[c];
      for (int l = 0; l < cnt; ++l) {
        const int cp = S.clist[l * n + c], cc = cp >> 3, pp = cp & 7;
#pragma unroll
        for (int r = 0; r < CS; ++r) {
          const int i = S.rows[cc * CS + r];
          if (i != 0xFFFF) acc = fma(S.Gd[(size_t)pp * pcap + cc * CS + r], S.z[i], acc);
        }
      }
      for (int k = 0; k < e; ++k) acc = fma(S.As[k * n + c], S.y[k], acc);
      S.rx[c] = acc + S.qd[c] * S.x[c] + p[c];
    }
    for (int pz = tid; pz < npos; pz += NT) {                       // rz = G x + s - h - F z
      const int i = S.rows[pz];
      if (i == 0xFFFF) continue;
      const int c = pz / CS, r = pz - c * CS, nc = S.ncols[c];
      T acc = 0;
      for (int q = 0; q < nc; ++q) acc = fma(S.Gd[(size_t)q * pcap + pz], S.x[S.ccols[q * pcap + c]], acc);
      T fz = 0;
#pragma unroll
      for (int q = 0; q < CS; ++q) {
        const int j = S.rows[c * CS + q];
        if (j != 0xFFFF) fz = fma(S.Fd[(size_t)c * CS * CS + r * CS + q], S.z[j], fz);
      }
      S.rz[i] = acc + S.s[i] - h[i] - fz;
    }
    for (int k = tid; k < e; k += NT) {                            // ry = A x - b
      T acc = 0;
      for (int j = 0; j < n; ++j) acc = fma(S.As[k * n + j], S.x[j], acc);
      S.ry[k] = acc - b[k];
    }
    __syncthreads();
    T q4[4] = {0, 0, 0, 0};                                        // s.z, |rz|^2, |ry|^2, |rx|^2
    for (int i = tid; i < m; i += NT) { q4[0] += S.s[i] * S.z[i]; q4[1] += S.rz[i] * S.rz[i]; }
    for (int i = tid; i < e; i += NT) q4[2] += S.ry[i] * S.ry[i];
    for (int i = tid; i < n; i += NT) q4[3] += S.rx[i] * S.rx[i];
    block_reduce<T, 4>(q4, OpSum(), T(0), S.red);
    const T sz = q4[0];
    const T mu = fabs(sz / T(m));                                  // :91
    const T resid = (e > 0 ? sqrt(q4[2]) : T(0)) + sqrt(q4[1]) + sqrt(q4[3]) + T(m) * mu;   // :92-96

    // ---- best iterate / termination (per scene)                 :107-136
    // (the reference refactors before this test, :98-102; the factors of a terminating iteration
    // are never used, so the test comes first here)
    bool improved;
    if (!have_best) { improved = true; have_best = true; not_improved = 0; }
    else { improved = resid < best; not_improved = improved ? 0 : not_improved + 1; }
    if (improved) {
      best = resid;
      for (int i = tid; i < n; i += NT) o_x[i] = S.x[i];
      for (int i = tid; i < m; i += NT) { o_z[i] = S.z[i]; o_s[i] = S.s[i]; }
      for (int i = tid; i < e; i += NT) o_y[i] = S.y[i];
    }
    if (not_improved == a.not_improved_lim) { status = 1; ++it; break; }
    if (best < a.eps) { status = 2; ++it; break; }
    if (mu > T(1e100)) { status = 3; ++it; break; }

    for (int i = tid; i < m; i += NT) S.d[i] = S.z[i] / S.s[i];     // :98
    __syncthreads();
    factor_kkt<T, NS, CS>(P, S, st);                               // :100

    // ---- affine direction                                       :138-139   (rs = z)
    solve_kkt<T, CS, QN>(P, S, st, S.rx, S.z, S.rz, e > 0 ? S.ry : nullptr, S.dx, S.ds, S.dz, S.dy);
    T stz, sts;
    get_steps(S.z, S.dz, S.s, S.ds, m, S.red, stz, sts);
    const T alpha_aff = nan_min(nan_min(stz, sts), T(1));          // :142-144
    T t3[1] = {0};
    for (int i = tid; i < m; i += NT) t3[0] += (S.s[i] + alpha_aff * S.ds[i]) * (S.z[i] + alpha_aff * S.dz[i]);
    block_reduce<T, 1>(t3, OpSum(), T(0), S.red);
    const T ratio = t3[0] / sz;                                    // :146-150
    const T sig = ratio * ratio * ratio;
    const T musig = -mu * sig;                                     // :152-158
    for (int i = tid; i < m; i += NT) S.rs2[i] = (musig + S.ds[i] * S.dz[i]) / S.s[i];
    __syncthreads();
    // corrector: outputs land in rx / rz / ry (dead until the next residual phase)
    solve_kkt<T, CS, QN>(P, S, st, nullptr, S.rs2, nullptr, nullptr, S.rx, S.rz, S.rs2, S.ry);
    // NOTE: ds_c -> S.rz, dz_c -> S.rs2 (solve_kkt reads rs before it writes dz/ds of the same row)
    for (int i = tid; i < n; i += NT) S.dx[i] += S.rx[i];          // :160-163
    for (int i = tid; i < m; i += NT) { S.ds[i] += S.rz[i]; S.dz[i] += S.rs2[i]; }
    for (int i = tid; i < e; i += NT) S.dy[i] += S.ry[i];
    __syncthreads();
    get_steps(S.z, S.dz, S.s, S.ds, m, S.red, stz, sts);
    const T alpha = nan_min(T(0.999) * nan_min(stz, sts), T(1));   // :164-166
    for (int i = tid; i < n; i += NT) S.x[i] += alpha * S.dx[i];   // :171-174
    for (int i = tid; i < m; i += NT) { S.s[i] += alpha * S.ds[i]; S.z[i] += alpha * S.dz[i]; }
    for (int i = tid; i < e; i += NT) S.y[i] += alpha * S.dy[i];
    __syncthreads();
  }
  if (tid == 0) { a.status[sc] = status; a.iters[sc] = it; if (a.resid) a.resid[sc] = best; }
}

template <typename T, int NS>
__global__ void __launch_bounds__(NT, (NS <= 3) ? 4 : ((NS <= 6) ? 2 : 1)) cond_forward_kernel(const CFwdArgs<T> a) {
  const CPlan& P = a.P;
  CSmem<T> S;
  S.carve(reinterpret_cast<char*>(cnd_smem), P);
  const int n = P.n, m = P.m, e = P.e, tid = threadIdx.x;
  __shared__ int singular_s;
  const T NANV = nan("");
  for (int sc = blockIdx.x; sc < a.B; sc += gridDim.x) {
    if (tid == 0) singular_s = 0;
    __syncthreads();
    Struct st;
    const bool ok = build_structure<T>(P, S, st, a.Q + (size_t)sc * n * n, a.G + (size_t)sc * m * n,
                                       e > 0 ? a.A + (size_t)sc * e * n : nullptr, a.F + (size_t)sc * m * m,
                                       &singular_s);
    __syncthreads();
    if (!ok) {
      if (singular_s) {                 // pdipm.py:361-368: the caller raises
        for (int i = tid; i < n; i += NT) a.zhat[(size_t)sc * n + i] = NANV;
        for (int i = tid; i < m; i += NT) { a.lam[(size_t)sc * m + i] = NANV; a.slack[(size_t)sc * m + i] = NANV; }
        for (int i = tid; i < e; i += NT) a.nu[(size_t)sc * e + i] = NANV;
        if (tid == 0) { a.status[sc] = -1; a.iters[sc] = 0; if (a.resid) a.resid[sc] = NANV; }
      } else if (tid == 0) {
        a.status[sc] = STATUS_UNSUPPORTED;
      }
      __syncthreads();
      continue;
    }
    switch (st.cs) {
      case 1: forward_scene<T, NS, 1>(a, S, st, sc); break;
      case 2: forward_scene<T, NS, 2>(a, S, st, sc); break;
      case 3: forward_scene<T, NS, 3>(a, S, st, sc); break;
      case 4: forward_scene<T, NS, 4>(a, S, st, sc); break;
      case 5: forward_scene<T, NS, 5>(a, S, st, sc); break;
      default: forward_scene<T, NS, 6>(a, S, st, sc); break;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ backward (lcp.py:37-64), one scene
template <typename T, int NS, int CS>
__device__ __forceinline__ void backward_scene(const CBwdArgs<T>& a, CSmem<T>& S, const Struct& st, int sc) {
  constexpr int QN = (NS + 1) / 2;
  const CPlan& P = a.P;
  const int n = P.n, m = P.m, e = P.e, tid = threadIdx.x;
  const T* zh = a.zhat + (size_t)sc * n;
  const T* lam = a.lam + (size_t)sc * m;
  const T* slk = a.slack + (size_t)sc * m;
  const T* nu = e > 0 ? a.nu + (size_t)sc * e : nullptr;
  for (int i = tid; i < n; i += NT) { S.x[i] = zh[i]; S.rx[i] = a.g[(size_t)sc * n + i]; }
  for (int i = tid; i < m; i += NT) { S.z[i] = lam[i]; S.s[i] = slk[i]; S.d[i] = lam[i] / slk[i]; S.rs2[i] = T(0); }   // :44
  for (int i = tid; i < e; i += NT) S.y[i] = nu[i];
  __syncthreads();
  factor_kkt<T, NS, CS>(P, S, st);                                                      // :46
  solve_kkt<T, CS, QN>(P, S, st, S.rx, S.rs2, nullptr, nullptr, S.dx, S.ds, S.dz, S.dy);   // :47-50
  const T* dx = S.dx; const T* dlam = S.dz; const T* dnu = S.dy;
  if (a.dp) for (int i = tid; i < n; i += NT) a.dp[(size_t)sc * n + i] = dx[i];                       // :52
  if (a.dh) for (int i = tid; i < m; i += NT) a.dh[(size_t)sc * m + i] = -dlam[i];                    // :55
  if (a.db && e > 0) for (int i = tid; i < e; i += NT) a.db[(size_t)sc * e + i] = -dnu[i];            // :58
  if (a.dG) {                                                    // :53  dlam (x) zhat + lam (x) dx
    T* o = a.dG + (size_t)sc * m * n;
    for (int t = tid; t < m * n; t += NT) { const int i = t / n, j = t - i * n; o[t] = dlam[i] * S.x[j] + S.z[i] * dx[j]; }
  }
  if (a.dF) {                                                    // :54  -dlam (x) lam
    T* o = a.dF + (size_t)sc * m * m;
    for (int t = tid; t < m * m; t += NT) { const int i = t / m, j = t - i * m; o[t] = -(dlam[i] * S.z[j]); }
  }
  if (a.dA && e > 0) {                                           // :57
    T* o = a.dA + (size_t)sc * e * n;
    for (int t = tid; t < e * n; t += NT) { const int i = t / n, j = t - i * n; o[t] = dnu[i] * S.x[j] + S.y[i] * dx[j]; }
  }
  if (a.dQ) {                                                    // :61
    T* o = a.dQ + (size_t)sc * n * n;
    for (int t = tid; t < n * n; t += NT) { const int i = t / n, j = t - i * n; o[t] = T(0.5) * (dx[i] * S.x[j] + S.x[i] * dx[j]); }
  }
  if (tid == 0) a.done[sc] = 1;
}

template <typename T, int NS>
__global__ void __launch_bounds__(NT, (NS <= 3) ? 4 : ((NS <= 6) ? 2 : 1)) cond_backward_kernel(const CBwdArgs<T> a) {
  const CPlan& P = a.P;
  CSmem<T> S;
  S.carve(reinterpret_cast<char*>(cnd_smem), P);
  const int n = P.n, m = P.m, e = P.e, tid = threadIdx.x;
  __shared__ int singular_s;
  for (int sc = blockIdx.x; sc < a.B; sc += gridDim.x) {
    if (tid == 0) singular_s = 0;
    __syncthreads();
    Struct st;
    const bool ok = build_structure<T>(P, S, st, a.Q + (size_t)sc * n * n, a.G + (size_t)sc * m * n,
                                       e > 0 ? a.A + (size_t)sc * e * n : nullptr, a.F + (size_t)sc * m * m,
                                       &singular_s);
    __syncthreads();
    if (!ok) {
      if (tid == 0) a.done[sc] = 0;
      __syncthreads();
      continue;
    }
    switch (st.cs) {
      case 1: backward_scene<T, NS, 1>(a, S, st, sc); break;
      case 2: backward_scene<T, NS, 2>(a, S, st, sc); break;
      case 3: backward_scene<T, NS, 3>(a, S, st, sc); break;
      case 4: backward_scene<T, NS, 4>(a, S, st, sc); break;
      case 5: backward_scene<T, NS, 5>(a, S, st, sc); break;
      default: backward_scene<T, NS, 6>(a, S, st, sc); break;
    }
    __syncthreads();
  }
}

}  // namespace cnd
}  // namespace lcpb200
