// lcp_assemble.cuh -- contact list -> dense mixed-LCP assembly and its adjoint.
//
// Restates world.py:144-234 (M, Jc, Jf, E, mu, restitutions: Python loops over
// contacts filling mostly-zero dense matrices) and engines.py:50-74 (G, F, h, p)
// as one pure function per OUTPUT element: every thread computes the value of the
// element it writes, so the dense Q/G/F are written exactly once, fully coalesced,
// with no memset and no atomics. fd = 2 (world.py:191-192 hard-codes dir2 = -dir1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

namespace lcpb200 {

template <typename T>
struct ContactView {
  const T *normal, *p1, *p2;   // [nc,2] for this scene
  const int32_t *b1, *b2;      // [nc]
};

// value of the Jacobian row of contact c along `dir` at column j (body = j/3)
template <typename T>
__device__ __forceinline__ T jrow_value(const ContactView<T>& cv, int c, T dx, T dy, int j) {
  const int body = j / 3, comp = j - 3 * body;
  if (body == cv.b1[c]) {
    if (comp == 0) return cv.p1[2 * c] * dy - cv.p1[2 * c + 1] * dx;       // cross_2d(p1, dir)
    return comp == 1 ? dx : dy;
  }
  if (body == cv.b2[c]) {
    if (comp == 0) return -(cv.p2[2 * c] * dy - cv.p2[2 * c + 1] * dx);
    return comp == 1 ? -dx : -dy;
  }
  return T(0);
}

template <typename T>
__global__ void assemble_kernel(int B, int nb, int nc, T dt, const T* __restrict__ mass,
                                const T* __restrict__ inertia, const T* __restrict__ v,
                                const T* __restrict__ fext, const T* __restrict__ normal,
                                const T* __restrict__ p1, const T* __restrict__ p2,
                                const int32_t* __restrict__ b1, const int32_t* __restrict__ b2,
                                const T* __restrict__ mu, const T* __restrict__ rest, T* __restrict__ Q,
                                T* __restrict__ p, T* __restrict__ G, T* __restrict__ h, T* __restrict__ F) {
  const int n = 3 * nb, m = 4 * nc, nf = 2 * nc;
  const long long per = (long long)n * n + n + (long long)m * n + m + (long long)m * m;
  const long long total = per * B;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int sc = (int)(idx / per);
    long long r = idx - (long long)sc * per;
    ContactView<T> cv{normal + (size_t)sc * nc * 2, p1 + (size_t)sc * nc * 2, p2 + (size_t)sc * nc * 2, b1, b2};
    if (r < (long long)n * n) {                                   // Q = blockdiag(I, m, m)  (world.py:57-61)
      const int i = (int)(r / n), j = (int)(r - (long long)i * n);
      T val = 0;
      if (i == j) { const int body = i / 3; val = (i - 3 * body == 0) ? inertia[(size_t)sc * nb + body] : mass[(size_t)sc * nb + body]; }
      Q[(size_t)sc * n * n + r] = val;
      continue;
    }
    r -= (long long)n * n;
    if (r < n) {                                                  // p = M v + dt f  (engines.py:32)
      const int j = (int)r, body = j / 3;
      const T md = (j - 3 * body == 0) ? inertia[(size_t)sc * nb + body] : mass[(size_t)sc * nb + body];
      p[(size_t)sc * n + j] = md * v[(size_t)sc * n + j] + dt * fext[(size_t)sc * n + j];
      continue;
    }
    r -= n;
    if (r < (long long)m * n) {                                   // G = [Jc; Jf; 0]  (engines.py:67-68)
      const int i = (int)(r / n), j = (int)(r - (long long)i * n);
      T val = 0;
      if (i < nc) {
        val = jrow_value(cv, i, cv.normal[2 * i], cv.normal[2 * i + 1], j);           // world.py:172-184
      } else if (i < nc + nf) {
        const int c = (i - nc) >> 1, k = (i - nc) & 1;
        const T sx = cv.normal[2 * c + 1], sy = -cv.normal[2 * c];                    // left_orthogonal(n)
        val = jrow_value(cv, c, k ? -sx : sx, k ? -sy : sy, j);                       // world.py:186-211
      }
      G[(size_t)sc * m * n + r] = val;
      continue;
    }
    r -= (long long)m * n;
    if (r < m) {                                                  // h = [(Jc v) rest, 0, 0]  (engines.py:53,74)
      const int i = (int)r;
      T val = 0;
      if (i < nc) {
        const T nx = cv.normal[2 * i], ny = cv.normal[2 * i + 1];
        const T* vs = v + (size_t)sc * n;
        const int j1 = 3 * b1[i], j2 = 3 * b2[i];
        T acc = 0;
        for (int q = 0; q < 3; ++q) acc += jrow_value(cv, i, nx, ny, j1 + q) * vs[j1 + q];
        for (int q = 0; q < 3; ++q) acc += jrow_value(cv, i, nx, ny, j2 + q) * vs[j2 + q];
        val = acc * rest[(size_t)sc * nc + i];
      }
      h[(size_t)sc * m + i] = val;
      continue;
    }
    r -= m;
    {                                                             // F  (engines.py:69-73)
      const int i = (int)(r / m), j = (int)(r - (long long)i * m);
      T val = 0;
      if (i >= nc && i < nc + nf) {                               // [0 0 E]
        if (j >= nc + nf && ((i - nc) >> 1) == j - nc - nf) val = T(1);
      } else if (i >= nc + nf) {                                  // [mu -E^T 0]
        const int c = i - nc - nf;
        if (j < nc) { if (j == c) val = mu[(size_t)sc * nc + c]; }
        else if (j < nc + nf) { if (((j - nc) >> 1) == c) val = T(-1); }
      }
      F[(size_t)sc * m * m + r] = val;
    }
  }
}

// adjoint, part 1: one thread per (scene, contact): geometry, mu, restitution
template <typename T>
__global__ void assemble_bwd_contacts_kernel(int B, int nb, int nc, const T* __restrict__ v,
                                             const T* __restrict__ normal, const T* __restrict__ p1,
                                             const T* __restrict__ p2, const int32_t* __restrict__ b1,
                                             const int32_t* __restrict__ b2, const T* __restrict__ rest,
                                             const T* __restrict__ dG, const T* __restrict__ dh,
                                             const T* __restrict__ dF, T* __restrict__ dnormal,
                                             T* __restrict__ dp1, T* __restrict__ dp2, T* __restrict__ dmu,
                                             T* __restrict__ drest) {
  const int n = 3 * nb, m = 4 * nc, nf = 2 * nc;
  const long long total = (long long)B * nc;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int sc = (int)(idx / nc), c = (int)(idx - (long long)sc * nc);
    const T nx = normal[idx * 2], ny = normal[idx * 2 + 1];
    const T p1x = p1[idx * 2], p1y = p1[idx * 2 + 1], p2x = p2[idx * 2], p2y = p2[idx * 2 + 1];
    const int j1 = 3 * b1[c], j2 = 3 * b2[c];
    const T* vs = v + (size_t)sc * n;
    const T* dGs = dG + (size_t)sc * m * n;
    const T rc = rest[idx], dhc = dh[(size_t)sc * m + c];
    T gnx = 0, gny = 0, g1x = 0, g1y = 0, g2x = 0, g2y = 0;
    // rows: 0 = normal, 1 = dir1 = (ny,-nx), 2 = dir2 = -dir1
    const int rows[3] = {c, nc + 2 * c, nc + 2 * c + 1};
    const T dxs[3] = {nx, ny, -ny}, dys[3] = {ny, -nx, nx};
    T jcv = 0;
    for (int q = 0; q < 3; ++q) {
      const T dx = dxs[q], dy = dys[q];
      T g[6];
      for (int t = 0; t < 3; ++t) { g[t] = dGs[(size_t)rows[q] * n + j1 + t]; g[3 + t] = dGs[(size_t)rows[q] * n + j2 + t]; }
      if (q == 0) {
        // h_c = rest_c * (row . v)
        const T row[6] = {p1x * dy - p1y * dx, dx, dy, -(p2x * dy - p2y * dx), -dx, -dy};
        for (int t = 0; t < 3; ++t) { jcv += row[t] * vs[j1 + t] + row[3 + t] * vs[j2 + t]; }
        for (int t = 0; t < 3; ++t) { g[t] += dhc * rc * vs[j1 + t]; g[3 + t] += dhc * rc * vs[j2 + t]; }
      }
      const T ddx = -p1y * g[0] + g[1] + p2y * g[3] - g[4];
      const T ddy = p1x * g[0] + g[2] - p2x * g[3] - g[5];
      g1x += dy * g[0]; g1y += -dx * g[0];
      g2x += -dy * g[3]; g2y += dx * g[3];
      if (q == 0) { gnx += ddx; gny += ddy; }
      else if (q == 1) { gny += ddx; gnx += -ddy; }               // dir1 = (ny, -nx)
      else { gny += -ddx; gnx += ddy; }                           // dir2 = (-ny, nx)
    }
    if (dnormal) { dnormal[idx * 2] = gnx; dnormal[idx * 2 + 1] = gny; }
    if (dp1) { dp1[idx * 2] = g1x; dp1[idx * 2 + 1] = g1y; }
    if (dp2) { dp2[idx * 2] = g2x; dp2[idx * 2 + 1] = g2y; }
    if (drest) drest[idx] = dhc * jcv;
    if (dmu) dmu[idx] = dF[(size_t)sc * m * m + (size_t)(nc + nf + c) * m + c];
  }
}

// adjoint, part 2: one thread per (scene, dof): mass, inertia, v, fext (deterministic, no atomics)
template <typename T>
__global__ void assemble_bwd_bodies_kernel(int B, int nb, int nc, T dt, const T* __restrict__ mass,
                                           const T* __restrict__ inertia, const T* __restrict__ v,
                                           const T* __restrict__ normal, const T* __restrict__ p1,
                                           const T* __restrict__ p2, const int32_t* __restrict__ b1,
                                           const int32_t* __restrict__ b2, const T* __restrict__ rest,
                                           const T* __restrict__ dQ, const T* __restrict__ dp,
                                           const T* __restrict__ dh, T* __restrict__ dmass,
                                           T* __restrict__ dinertia, T* __restrict__ dv, T* __restrict__ dfext) {
  const int n = 3 * nb, m = 4 * nc;
  const long long total = (long long)B * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int sc = (int)(idx / n), j = (int)(idx - (long long)sc * n);
    const int body = j / 3, comp = j - 3 * body;
    const T md = comp == 0 ? inertia[(size_t)sc * nb + body] : mass[(size_t)sc * nb + body];
    const T dpj = dp[idx];
    if (dfext) dfext[idx] = dt * dpj;
    if (dv) {
      ContactView<T> cv{normal + (size_t)sc * nc * 2, p1 + (size_t)sc * nc * 2, p2 + (size_t)sc * nc * 2, b1, b2};
      T acc = md * dpj;
      for (int c = 0; c < nc; ++c) {
        if (b1[c] != body && b2[c] != body) continue;
        acc += dh[(size_t)sc * m + c] * rest[(size_t)sc * nc + c] *
               jrow_value(cv, c, cv.normal[2 * c], cv.normal[2 * c + 1], j);
      }
      dv[idx] = acc;
    }
    // d(Mdiag_j) = dQ_jj + dp_j v_j ; inertia <- comp 0, mass <- comps 1 and 2 (summed by the comp-1 thread)
    const T* dQs = dQ + (size_t)sc * n * n;
    if (comp == 0 && dinertia) dinertia[(size_t)sc * nb + body] = dQs[(size_t)j * n + j] + dpj * v[idx];
    if (comp == 1 && dmass)
      dmass[(size_t)sc * nb + body] = dQs[(size_t)j * n + j] + dpj * v[idx] +
                                      dQs[(size_t)(j + 1) * n + j + 1] + dp[idx + 1] * v[idx + 1];
  }
}

template <typename T>
static void launch_assemble(int B, int nb, int nc, T dt, const T* mass, const T* inertia, const T* v,
                            const T* fext, const T* normal, const T* p1, const T* p2, const int32_t* b1,
                            const int32_t* b2, const T* mu, const T* rest, T* Q, T* p, T* G, T* h, T* F,
                            cudaStream_t st) {
  const int n = 3 * nb, m = 4 * nc;
  const long long total = ((long long)n * n + n + (long long)m * n + m + (long long)m * m) * B;
  const int threads = 256;
  const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148LL * 16);
  assemble_kernel<T><<<blocks, threads, 0, st>>>(B, nb, nc, dt, mass, inertia, v, fext, normal, p1, p2, b1, b2, mu,
                                                 rest, Q, p, G, h, F);
}

template <typename T>
static void launch_assemble_backward(int B, int nb, int nc, T dt, const T* mass, const T* inertia, const T* v,
                                     const T* normal, const T* p1, const T* p2, const int32_t* b1,
                                     const int32_t* b2, const T* mu, const T* rest, const T* dQ, const T* dp,
                                     const T* dG, const T* dh, const T* dF, T* dmass, T* dinertia, T* dv,
                                     T* dfext, T* dnormal, T* dp1, T* dp2, T* dmu, T* drest, cudaStream_t st) {
  (void)mu;
  const int threads = 128;
  {
    const long long total = (long long)B * nc;
    const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148LL * 16);
    assemble_bwd_contacts_kernel<T><<<blocks, threads, 0, st>>>(B, nb, nc, v, normal, p1, p2, b1, b2, rest, dG, dh,
                                                                dF, dnormal, dp1, dp2, dmu, drest);
  }
  {
    const long long total = (long long)B * 3 * nb;
    const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148LL * 16);
    assemble_bwd_bodies_kernel<T><<<blocks, threads, 0, st>>>(B, nb, nc, dt, mass, inertia, v, normal, p1, p2, b1,
                                                              b2, rest, dQ, dp, dh, dmass, dinertia, dv, dfext);
  }
}

}  // namespace lcpb200
