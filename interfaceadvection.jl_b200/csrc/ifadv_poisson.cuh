// ifadv_poisson.cuh -- variable-coefficient pressure projection: the reference's Jacobi-preconditioned
// conjugate-gradient solver psolver! (src/flow.jl:300-326), inproject! (:343-347) and myproject! (:328-341) on WaterLily's Poisson
// arrays (L ≡ Flow.μ₀, x ≡ Flow.p, z ≡ Flow.σ; D, iD, ϵ, r).
//
// The reference runs, per iteration, 2·|perdir| ghost-plane copies + 5 field passes + 3 dot products, each dot product a device ->
// host synchronisation (alpha, beta and the loop condition are host scalars).  Here an iteration is three kernels and NO host round
// trip: every scalar of the recurrence (rho, z·ϵ, beta, r₂, the iteration count, the loop condition) lives in a control block in
// device memory, written by the LAST CTA of the kernel that completes the corresponding reduction:
//   pois_mult_kernel    z = A ϵ on inside(x);                            Σ z·ϵ                                   (:312-313)   6 s B/cell
//   pois_update_kernel  alpha = rho / (z·ϵ); x += alpha ϵ; r -= alpha z; z = r·iD;   Σ r·z, Σ r·r               (:313-317,321) 8 s B/cell
//                       last CTA: beta = rho2/rho, rho = rho2, r₂, n += 1, done = !(r₂ > tol && n < itmx)        (:309,318,320)
//   pois_dir_kernel     ϵ = beta ϵ + z                                                                           (:319)       3 s B/cell
// (+ one launch for perBC!(ϵ) when a direction is periodic) = 17 s B per cell and iteration.  Kernels of iterations past convergence
// return at once (they read `done`), so the host enqueues iterations in batches and polls the control block one batch behind.
// Reductions: Float64 partial sums per CTA in a fixed order, summed by the last CTA in a fixed order (deterministic for a grid
// size), rounded to T -- the reference's `⋅` is BLAS / CUBLAS dot with an unspecified order, so comparisons carry a tolerance.
// Per-cell arithmetic follows the reference expression by expression (-fmad=false, IEEE division).
#pragma once
#include <cooperative_groups.h>

#include "ifadv_math.cuh"
#include "ifadv_sweep.cuh"

namespace ifadv {

#define IFADV_POIS_MAXB 2048  // upper bound of the reduction grids

struct PoisCtl {
  double rho, zeps, beta, r2, mean, tol, r2_0;
  int n, itmx, done, sub_mean;
  int slab, pad_;   // slab != 0 (z-slab context): the last CTA only stores its rank's sums in acc[]; the host all-reduces them and
  double acc[4];    // pois_fin_kernel applies the scalar step -- identical values, hence identical decisions, on every rank
  unsigned ticket[4];
  double part[3][IFADV_POIS_MAXB];
  // pcg! smoother of the multigrid levels (ifadv_mlpoisson.cuh): step length, "still iterating", "r2 belongs to the current r"
  double alpha;
  int live, r2_valid;
};

template <class T> struct teps;
template <> struct teps<float> { static constexpr float v = 1.1920928955078125e-07f; };
template <> struct teps<double> { static constexpr double v = 2.220446049250313e-16; };

// CTA partial sums -> part[q][blockIdx.x]; returns true in every thread of the last CTA to arrive, with tot[q] the grid totals
// (valid in thread 0).  NQ <= 2.
template <int NQ> IFADV_DI bool grid_reduce(double (&v)[NQ], PoisCtl* ctl, int tk, double (&tot)[NQ]) {
  __shared__ double ws[8][NQ];
  __shared__ int last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    double t = v[q];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if (lane == 0) ws[wid][q] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      double t = 0.0;
      for (int w = 0; w < wpb; ++w) t += ws[w][q];
      ctl->part[q][blockIdx.x] = t;
    }
    __threadfence();
    last = (atomicAdd(&ctl->ticket[tk], 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += __ldcg(&ctl->part[q][b]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    __syncthreads();
    if (lane == 0) ws[wid][q] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      double t = 0.0;
      for (int w = 0; w < wpb; ++w) t += ws[w][q];
      tot[q] = t;
    }
    ctl->ticket[tk] = 0u;
  }
  return true;
}

// mult(I,L,D,x) = x[I]·D[I] + Σᵢ L[I,i]·x[I-δᵢ] + Σᵢ L[I+δᵢ,i]·x[I+δᵢ]   (WaterLily Poisson.jl, restated)
template <class T, int D> IFADV_DI T pois_mult(const T* __restrict__ L, const T* __restrict__ Dg, const T* __restrict__ x, const Geo& g, long long l) {
  T lo = T(0), up = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const long long st = (i == 0) ? 1 : ((i == 1) ? g.s1 : g.s2);
    lo = lo + __ldg(L + (long long)i * g.S + l) * __ldg(x + l - st);
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const long long st = (i == 0) ? 1 : ((i == 1) ? g.s1 : g.s2);
    up = up + __ldg(L + (long long)i * g.S + l + st) * __ldg(x + l + st);
  }
  return __ldg(x + l) * __ldg(Dg + l) + lo + up;
}

// rows of inside(x) -- planes [kz0, kz1) of a z-slab -- walked by warps, lanes along the contiguous dimension
#define IFADV_POIS_ROWS(...)                                                                                   \
  const int ny = g.n[1] - 2, nz = (D == 3) ? kz1 - kz0 : 1;                                                    \
  const long long rows = (long long)ny * nz;                                                                   \
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wid = threadIdx.x >> 5;                            \
  for (long long rw = (long long)blockIdx.x * wpb + wid; rw < rows; rw += (long long)gridDim.x * wpb) {        \
    const int y = 2 + (int)(rw % ny), zc = (D == 3) ? kz0 + (int)(rw / ny) : 1;                                \
    const long long l0 = lin3(g, 0, y, zc);                                                                    \
    for (int xc = 2 + lane; xc <= g.n[0] - 1; xc += 32) {                                                      \
      const long long l = l0 + xc;                                                                             \
      __VA_ARGS__                                                                                                  \
    }                                                                                                          \
  }

// The same walk with U cells per lane and trip: the kernels below load all U cells first, then compute, then store, so that a thread
// keeps U times as many loads in flight (the solver is a pure streaming workload: latency x bandwidth needs ~40 KB in flight per SM).
#define IFADV_POIS_ROWS_U(U, ...)                                                                              \
  const int ny = g.n[1] - 2, nz = (D == 3) ? kz1 - kz0 : 1;                                                    \
  const long long rows = (long long)ny * nz;                                                                   \
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wid = threadIdx.x >> 5;                            \
  const int xlast = g.n[0] - 1;                                                                                \
  for (long long rw = (long long)blockIdx.x * wpb + wid; rw < rows; rw += (long long)gridDim.x * wpb) {        \
    const int y = 2 + (int)(rw % ny), zc = (D == 3) ? kz0 + (int)(rw / ny) : 1;                                \
    const long long l0 = lin3(g, 0, y, zc);                                                                    \
    for (int xb = 2 + lane; xb <= xlast; xb += 32 * (U)) {                                                     \
      __VA_ARGS__                                                                                              \
    }                                                                                                          \
  }

// set_diag!(D,iD,L) = update!(p::Poisson): D = -Σᵢ (L[I,i] + L[I+δᵢ,i]); iD = D² < 2eps ? 0 : 1/D on inside
template <class T, int D> __global__ void __launch_bounds__(256) pois_diag_kernel(T* __restrict__ Dg, T* __restrict__ iD, const T* __restrict__ L, const Geo g, int kz0, int kz1) {
  IFADV_POIS_ROWS({
    T s = T(0);
_Pragma("unroll")
    for (int i = 0; i < D; ++i) {
      const long long st = (i == 0) ? 1 : ((i == 1) ? g.s1 : g.s2);
      s = s - (__ldg(L + (long long)i * g.S + l) + __ldg(L + (long long)i * g.S + l + st));
    }
    Dg[l] = s;
    iD[l] = (s * s < T(2) * teps<T>::v) ? T(0) : T(1) / s;
  })
}

// perBC!(a,perdir): every cell with a ghost coordinate in a periodic direction takes the value of the cell with those coordinates
// wrapped (closed form of the sequential plane copies; coordinates of non-periodic directions stay as they are)
template <class T, int D> __global__ void perbc_kernel(T* a, const Geo g) {
  const long long n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const long long c0 = (g.per & 1u) ? 2 * n1 * n2 : 0, c1 = (g.per & 2u) ? 2 * n0 * n2 : 0, c2 = (D == 3 && (g.per & 4u)) ? 2 * n0 * n1 : 0;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= c0 + c1 + c2) return;
  int x, y, z;
  if (t < c0) { const long long r = t >> 1; x = (t & 1) ? (int)n0 : 1; y = (int)(r % n1) + 1; z = (int)(r / n1) + 1; }
  else if (t < c0 + c1) { const long long q = t - c0, r = q >> 1; y = (q & 1) ? (int)n1 : 1; x = (int)(r % n0) + 1; z = (int)(r / n0) + 1; }
  else { const long long q = t - c0 - c1, r = q >> 1; z = (q & 1) ? (int)n2 : 1; x = (int)(r % n0) + 1; y = (int)(r / n0) + 1; }
  const int mx = (g.per & 1u) ? wrapc(x, g.n[0]) : x, my = (g.per & 2u) ? wrapc(y, g.n[1]) : y, mz = (D == 3 && (g.per & 4u)) ? wrapc(z, g.n[2]) : z;
  a[lin3(g, x, y, z)] = a[lin3(g, mx, my, mz)];
}

// inproject!: z = div(I,u) on inside(z), 0 elsewhere; ϵ = r = 0; x *= dt over all entries        (flow.jl:344-345)
template <class T, int D> __global__ void pois_setup_kernel(T* __restrict__ x, T* __restrict__ eps, T* __restrict__ r, T* __restrict__ z,
                                                            const T* __restrict__ u, const Geo g, T dt) {
  const int xc = 1 + blockIdx.x * blockDim.x + threadIdx.x, y = 1 + blockIdx.y, zc = 1 + blockIdx.z;
  if (xc > g.n[0]) return;
  const long long l = lin3(g, xc, y, zc);
  const bool in = xc >= 2 && xc <= g.n[0] - 1 && y >= 2 && y <= g.n[1] - 1 && (D == 2 || (zc >= 2 && zc <= g.n[2] - 1));
  T s = T(0);
  if (in) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const long long st = (i == 0) ? 1 : ((i == 1) ? g.s1 : g.s2);
      s = s + (__ldg(u + (long long)i * g.S + l + st) - __ldg(u + (long long)i * g.S + l));
    }
  }
  z[l] = s;
  eps[l] = T(0);
  r[l] = T(0);
  x[l] = x[l] * dt;
}

// ---- the scalar steps behind the four reductions (last CTA of the kernel, or pois_fin_kernel after the all-reduce of a z-slab run) ----
template <class T> IFADV_DI void fin_residual(PoisCtl* ctl, double sum, double cnt, double tol, int itmx) {
  const T s = (T)sum / (T)cnt;                                     // sum(p.r)/length(inside(p.r))
  ctl->mean = (double)s;
  ctl->sub_mean = (t_abs(s) <= T(2) * teps<T>::v) ? 0 : 1;
  ctl->tol = tol;
  ctl->itmx = itmx;
  ctl->n = 0;
  ctl->done = 0;
}
template <class T> IFADV_DI void fin_start(PoisCtl* ctl, double rr, double rz) {
  const T r2 = (T)rr, tol = (T)ctl->tol;
  ctl->r2 = (double)r2;
  ctl->r2_0 = (double)r2;
  ctl->rho = (double)(T)rz;
  ctl->done = ((r2 > tol || r2 > tol / T(4)) && 0 < ctl->itmx) ? 0 : 1;  // flow.jl:309 with nᵖ == 0
}
template <class T> IFADV_DI void fin_mult(PoisCtl* ctl, double ze) { ctl->zeps = (double)(T)ze; }
template <class T> IFADV_DI void fin_update(PoisCtl* ctl, double rz, double rr) {
  const T rho2 = (T)rz, r2 = (T)rr;
  ctl->beta = (double)(rho2 / (T)ctl->rho);                        // :318
  ctl->rho = (double)rho2;                                         // :320
  ctl->r2 = (double)r2;                                            // :321
  const int n = ctl->n + 1;
  ctl->n = n;
  ctl->done = (r2 > (T)ctl->tol && n < ctl->itmx) ? 0 : 1;         // :309 with nᵖ >= 1
}
// z-slab runs: acc[] holds the all-reduced sums.  `which`: 0 residual, 1 start, 2 mult, 3 update -- skipped like the kernel it follows
template <class T> __global__ void pois_fin_kernel(PoisCtl* ctl, int which, double tol, int itmx) {
  if (which == 0) fin_residual<T>(ctl, ctl->acc[0], ctl->acc[1], tol, itmx);
  else if (which == 1) fin_start<T>(ctl, ctl->acc[0], ctl->acc[1]);
  else if (ctl->done) return;
  else if (which == 2) fin_mult<T>(ctl, ctl->acc[0]);
  else fin_update<T>(ctl, ctl->acc[0], ctl->acc[1]);
}

// residual!(p), first half: r = iD == 0 ? 0 : z - A x on inside;  Σ r -> mean s = Σr / |inside|, dropped when |s| <= 2eps
template <class T, int D> __global__ void __launch_bounds__(256) pois_residual_kernel(T* __restrict__ r, const T* __restrict__ z, const T* __restrict__ x,
                                                                                      const T* __restrict__ L, const T* __restrict__ Dg,
                                                                                      const T* __restrict__ iD, const Geo g, PoisCtl* ctl,
                                                                                      double tol, int itmx, int kz0, int kz1) {
  double acc[1] = {0.0};
  IFADV_POIS_ROWS({
    const T v = (__ldg(iD + l) == T(0)) ? T(0) : __ldg(z + l) - pois_mult<T, D>(L, Dg, x, g, l);
    r[l] = v;
    acc[0] += (double)v;
  })
  double tot[1];
  if (grid_reduce<1>(acc, ctl, 0, tot) && threadIdx.x == 0) {
    const double cnt = (double)(rows * (long long)(g.n[0] - 2));
    if (ctl->slab) { ctl->acc[0] = tot[0]; ctl->acc[1] = cnt; }
    else fin_residual<T>(ctl, tot[0], cnt, tol, itmx);
  }
}

// residual!, second half (r -= s) fused with psolver!'s start: z = ϵ = r·iD; r₂ = r·r, rho = r·z          (flow.jl:302-307)
template <class T, int D> __global__ void __launch_bounds__(256) pois_start_kernel(T* __restrict__ r, T* __restrict__ z, T* __restrict__ eps,
                                                                                   const T* __restrict__ iD, const Geo g, PoisCtl* ctl, int kz0, int kz1) {
  const bool sub = ctl->sub_mean != 0;
  const T s = (T)ctl->mean;
  double acc[2] = {0.0, 0.0};
  IFADV_POIS_ROWS({
    T rv = r[l];
    if (sub) { rv = rv - s; r[l] = rv; }
    const T zv = rv * __ldg(iD + l);
    z[l] = zv;
    eps[l] = zv;
    acc[0] += (double)rv * (double)rv;
    acc[1] += (double)rv * (double)zv;
  })
  double tot[2];
  if (grid_reduce<2>(acc, ctl, 1, tot) && threadIdx.x == 0) {
    if (ctl->slab) { ctl->acc[0] = tot[0]; ctl->acc[1] = tot[1]; }
    else fin_start<T>(ctl, tot[0], tot[1]);
  }
}

// z = A ϵ on inside;  Σ z·ϵ                                                                              (flow.jl:312-313)
template <class T, int D> __global__ void __launch_bounds__(256) pois_mult_kernel(T* __restrict__ z, const T* __restrict__ eps, const T* __restrict__ L,
                                                                                  const T* __restrict__ Dg, const Geo g, PoisCtl* ctl, int kz0, int kz1) {
  if (ctl->done) return;
  constexpr int U = (sizeof(T) == 4) ? 4 : 2;
  double acc[1] = {0.0};
  IFADV_POIS_ROWS_U(U, {
    T ec[U], dg[U], ll[U][D], lu[U][D], em[U][D], ep[U][D];
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) {
        const long long l = l0 + xc;
        ec[k] = __ldg(eps + l);
        dg[k] = __ldg(Dg + l);
_Pragma("unroll")
        for (int i = 0; i < D; ++i) {
          const long long st = (i == 0) ? 1 : ((i == 1) ? g.s1 : g.s2);
          ll[k][i] = __ldg(L + (long long)i * g.S + l);
          lu[k][i] = __ldg(L + (long long)i * g.S + l + st);
          em[k][i] = __ldg(eps + l - st);
          ep[k][i] = __ldg(eps + l + st);
        }
      }
    }
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) {
        T lo = T(0), up = T(0);
_Pragma("unroll")
        for (int i = 0; i < D; ++i) lo = lo + ll[k][i] * em[k][i];
_Pragma("unroll")
        for (int i = 0; i < D; ++i) up = up + lu[k][i] * ep[k][i];
        const T v = ec[k] * dg[k] + lo + up;
        z[l0 + xc] = v;
        acc[0] += (double)v * (double)ec[k];
      }
    }
  })
  double tot[1];
  if (grid_reduce<1>(acc, ctl, 2, tot) && threadIdx.x == 0) {
    if (ctl->slab) ctl->acc[0] = tot[0];
    else fin_mult<T>(ctl, tot[0]);
  }
}

// The same product, marching: a thread owns a column (x, y) and walks a chunk of z planes with ϵ(z-1), ϵ(z), ϵ(z+1) and the two z-face
// coefficients in registers -- 11 loads per cell instead of 14, the three dropped ones being L2 hits in the row form (3-D only).  Lanes run
// along x, the 8 warps of a CTA along y (their y-neighbours meet in L1).  Identical per-cell arithmetic; MODE 0: psolver! (gate `done`,
// scalar step fin_mult), MODE 1: the multigrid smoother (gate `live`, alpha and its range check, ifadv_mlpoisson.cuh).
template <class T, int MODE> IFADV_DI void fin_mult_mode(PoisCtl* ctl, double ze) {
  if (MODE == 0) {
    if (ctl->slab) ctl->acc[0] = ze;
    else fin_mult<T>(ctl, ze);
  } else {
    const T alpha = (T)ctl->rho / (T)ze;
    ctl->alpha = (double)alpha;
    const double aa = fabs((double)alpha);
    if (aa < 1e-2 || aa > 1e3) ctl->live = 0;
  }
}
template <class T, int MODE> __global__ void __launch_bounds__(256) pois_mult_march_kernel(T* __restrict__ z, const T* __restrict__ eps,
                                                                                           const T* __restrict__ L, const T* __restrict__ Dg,
                                                                                           const Geo g, PoisCtl* ctl, int kz0, int kz1, int chunk) {
  if (MODE == 0 ? (ctl->done != 0) : (ctl->live == 0)) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nx = g.n[0] - 2, ny = g.n[1] - 2;
  const int xseg = (nx + 31) >> 5, ygrp = (ny + 7) >> 3, nzc = (kz1 - kz0 + chunk - 1) / chunk;
  const long long items = (long long)xseg * ygrp * nzc;
  const long long s1 = g.s1, s2 = g.s2;
  const T* __restrict__ Lx = L;
  const T* __restrict__ Ly = L + g.S;
  const T* __restrict__ Lz = L + 2 * g.S;
  double acc[1] = {0.0};
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const int xs = (int)(it % xseg);
    const long long q = it / xseg;
    const int yg = (int)(q % ygrp), zc = (int)(q / ygrp);
    const int x = 2 + (xs << 5) + lane, y = 2 + (yg << 3) + wid;
    if (x > g.n[0] - 1 || y > g.n[1] - 1) continue;
    const int za = kz0 + zc * chunk, zb = min(kz1, za + chunk);
    long long l = lin3(g, x, y, za);
    T em = __ldg(eps + l - s2), ec = __ldg(eps + l), lzc = __ldg(Lz + l);
    // the 11 loads of plane zz+1 are issued before the arithmetic of plane zz (two planes of loads in flight per thread)
    T ep = __ldg(eps + l + s2), lzp = __ldg(Lz + l + s2);
    T exm = __ldg(eps + l - 1), exp_ = __ldg(eps + l + 1), eym = __ldg(eps + l - s1), eyp = __ldg(eps + l + s1);
    T lx = __ldg(Lx + l), lxp = __ldg(Lx + l + 1), ly = __ldg(Ly + l), lyp = __ldg(Ly + l + s1), dg = __ldg(Dg + l);
    for (int zz = za; zz < zb; ++zz, l += s2) {
      T n_ep = T(0), n_lzp = T(0), n_exm = T(0), n_exp = T(0), n_eym = T(0), n_eyp = T(0), n_lx = T(0), n_lxp = T(0), n_ly = T(0), n_lyp = T(0),
        n_dg = T(0);
      if (zz + 1 < zb) {
        const long long ln = l + s2;
        n_ep = __ldg(eps + ln + s2); n_lzp = __ldg(Lz + ln + s2);
        n_exm = __ldg(eps + ln - 1); n_exp = __ldg(eps + ln + 1); n_eym = __ldg(eps + ln - s1); n_eyp = __ldg(eps + ln + s1);
        n_lx = __ldg(Lx + ln); n_lxp = __ldg(Lx + ln + 1); n_ly = __ldg(Ly + ln); n_lyp = __ldg(Ly + ln + s1); n_dg = __ldg(Dg + ln);
      }
      T lo = T(0), up = T(0);
      lo = lo + lx * exm; lo = lo + ly * eym; lo = lo + lzc * em;
      up = up + lxp * exp_; up = up + lyp * eyp; up = up + lzp * ep;
      const T v = ec * dg + lo + up;
      z[l] = v;
      acc[0] += (double)v * (double)ec;
      em = ec; ec = ep; lzc = lzp;
      ep = n_ep; lzp = n_lzp; exm = n_exm; exp_ = n_exp; eym = n_eym; eyp = n_eyp; lx = n_lx; lxp = n_lxp; ly = n_ly; lyp = n_lyp; dg = n_dg;
    }
  }
  double tot[1];
  if (grid_reduce<1>(acc, ctl, 2, tot) && threadIdx.x == 0) fin_mult_mode<T, MODE>(ctl, tot[0]);
}

// x += alpha ϵ; r -= alpha z; z = r·iD;  Σ r·z, Σ r·r; last CTA: the scalar recurrence and the loop condition   (flow.jl:313-321)
template <class T, int D> __global__ void __launch_bounds__(256) pois_update_kernel(T* __restrict__ x, T* __restrict__ r, T* __restrict__ z,
                                                                                    const T* __restrict__ eps, const T* __restrict__ iD, const Geo g,
                                                                                    PoisCtl* ctl, int kz0, int kz1) {
  if (ctl->done) return;
  const T alpha = (T)ctl->rho / (T)ctl->zeps;
  constexpr int U = (sizeof(T) == 4) ? 8 : 4;
  double acc[2] = {0.0, 0.0};
  IFADV_POIS_ROWS_U(U, {
    T xv[U], ev[U], rv[U], zv[U], dv[U];
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) {
        const long long l = l0 + xc;
        xv[k] = x[l]; ev[k] = __ldg(eps + l); rv[k] = r[l]; zv[k] = z[l]; dv[k] = __ldg(iD + l);
      }
    }
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) {
        const long long l = l0 + xc;
        x[l] = xv[k] + alpha * ev[k];
        const T rn = rv[k] - alpha * zv[k];
        r[l] = rn;
        const T zn = rn * dv[k];
        z[l] = zn;
        acc[0] += (double)rn * (double)zn;
        acc[1] += (double)rn * (double)rn;
      }
    }
  })
  double tot[2];
  if (grid_reduce<2>(acc, ctl, 3, tot) && threadIdx.x == 0) {
    if (ctl->slab) { ctl->acc[0] = tot[0]; ctl->acc[1] = tot[1]; }
    else fin_update<T>(ctl, tot[0], tot[1]);
  }
}

// ϵ = beta ϵ + z on inside; runs iff the update kernel of iteration `it` ran (n == it + 1)                 (flow.jl:319)
template <class T, int D> __global__ void __launch_bounds__(256) pois_dir_kernel(T* __restrict__ eps, const T* __restrict__ z, const Geo g,
                                                                                 const PoisCtl* ctl, int it, int kz0, int kz1) {
  if (ctl->n != it + 1) return;
  const T beta = (T)ctl->beta;
  constexpr int U = (sizeof(T) == 4) ? 8 : 4;
  IFADV_POIS_ROWS_U(U, {
    T ev[U], zv[U];
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) { ev[k] = eps[l0 + xc]; zv[k] = __ldg(z + l0 + xc); }
    }
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) eps[l0 + xc] = beta * ev[k] + zv[k];
    }
  })
}

// ---- very small grids: a batch of iterations in ONE cooperative launch -------------------------------------------------------------
// On small grids an iteration of the three-kernel form is bound by launch latency (32³: 28 µs per iteration, 3 launches).  Here the
// CTAs stay resident and meet at grid-wide barriers instead (32³: 17 µs; no gain from 64³ on, see ifadv_poisson.cu): perBC!(ϵ) | barrier | z = Aϵ, Σ z·ϵ |
// barrier | x, r, z update, Σ r·z, Σ r·r, scalar step | barrier | ϵ = beta ϵ + z | barrier.  Same per-cell arithmetic and the same
// last-CTA reductions (partial sums grouped by this launch's grid, so the dot products agree with the three-kernel form to round-off).
// Arrays written inside the launch are read with plain loads (never through the read-only path).
template <class T, int D> IFADV_DI T pois_mult_plain(const T* __restrict__ L, const T* __restrict__ Dg, const T* x, const Geo& g, long long l) {
  T lo = T(0), up = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const long long st = (i == 0) ? 1 : ((i == 1) ? g.s1 : g.s2);
    lo = lo + __ldg(L + (long long)i * g.S + l) * x[l - st];
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const long long st = (i == 0) ? 1 : ((i == 1) ? g.s1 : g.s2);
    up = up + __ldg(L + (long long)i * g.S + l + st) * x[l + st];
  }
  return x[l] * __ldg(Dg + l) + lo + up;
}
template <class T, int D> __global__ void __launch_bounds__(256) pois_pcg_coop_kernel(T* x, T* r, T* z, T* eps, const T* __restrict__ L,
                                                                                      const T* __restrict__ Dg, const T* __restrict__ iD, const Geo g,
                                                                                      PoisCtl* ctl, int it0, int it1, int kz0, int kz1) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  volatile PoisCtl* vc = ctl;
  const int ny = g.n[1] - 2, nz = (D == 3) ? kz1 - kz0 : 1;
  const long long rows = (long long)ny * nz;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wid = threadIdx.x >> 5;
  const long long n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const long long c0 = (g.per & 1u) ? 2 * n1 * n2 : 0, c1 = (g.per & 2u) ? 2 * n0 * n2 : 0, c2 = (D == 3 && (g.per & 4u)) ? 2 * n0 * n1 : 0;
  for (int it = it0; it < it1; ++it) {
    if (vc->done) break;  // uniform: written before the last barrier of the previous iteration (or by an earlier launch)
    if (g.per) {          // perBC!(ϵ), flow.jl:311
      for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < c0 + c1 + c2; t += (long long)gridDim.x * blockDim.x) {
        int xx, yy, zz;
        if (t < c0) { const long long q = t >> 1; xx = (t & 1) ? (int)n0 : 1; yy = (int)(q % n1) + 1; zz = (int)(q / n1) + 1; }
        else if (t < c0 + c1) { const long long u_ = t - c0, q = u_ >> 1; yy = (u_ & 1) ? (int)n1 : 1; xx = (int)(q % n0) + 1; zz = (int)(q / n0) + 1; }
        else { const long long u_ = t - c0 - c1, q = u_ >> 1; zz = (u_ & 1) ? (int)n2 : 1; xx = (int)(q % n0) + 1; yy = (int)(q / n0) + 1; }
        const int mx = (g.per & 1u) ? wrapc(xx, g.n[0]) : xx, my = (g.per & 2u) ? wrapc(yy, g.n[1]) : yy,
                  mz = (D == 3 && (g.per & 4u)) ? wrapc(zz, g.n[2]) : zz;
        eps[lin3(g, xx, yy, zz)] = eps[lin3(g, mx, my, mz)];
      }
      grid.sync();
    }
    {  // z = A ϵ, Σ z·ϵ   (:312-313)
      double acc[1] = {0.0};
      for (long long rw = (long long)blockIdx.x * wpb + wid; rw < rows; rw += (long long)gridDim.x * wpb) {
        const long long l0 = lin3(g, 0, 2 + (int)(rw % ny), (D == 3) ? kz0 + (int)(rw / ny) : 1);
        for (int xc = 2 + lane; xc <= g.n[0] - 1; xc += 32) {
          const long long l = l0 + xc;
          const T v = pois_mult_plain<T, D>(L, Dg, eps, g, l);
          z[l] = v;
          acc[0] += (double)v * (double)eps[l];
        }
      }
      double tot[1];
      if (grid_reduce<1>(acc, ctl, 2, tot) && threadIdx.x == 0) fin_mult<T>(ctl, tot[0]);
    }
    grid.sync();
    {  // x, r, z update, Σ r·z, Σ r·r, scalar step   (:313-321)
      const T alpha = (T)vc->rho / (T)vc->zeps;
      double acc[2] = {0.0, 0.0};
      for (long long rw = (long long)blockIdx.x * wpb + wid; rw < rows; rw += (long long)gridDim.x * wpb) {
        const long long l0 = lin3(g, 0, 2 + (int)(rw % ny), (D == 3) ? kz0 + (int)(rw / ny) : 1);
        for (int xc = 2 + lane; xc <= g.n[0] - 1; xc += 32) {
          const long long l = l0 + xc;
          x[l] = x[l] + alpha * eps[l];
          const T rn = r[l] - alpha * z[l];
          r[l] = rn;
          const T zn = rn * __ldg(iD + l);
          z[l] = zn;
          acc[0] += (double)rn * (double)zn;
          acc[1] += (double)rn * (double)rn;
        }
      }
      double tot[2];
      if (grid_reduce<2>(acc, ctl, 3, tot) && threadIdx.x == 0) fin_update<T>(ctl, tot[0], tot[1]);
    }
    grid.sync();
    {  // ϵ = beta ϵ + z   (:319)
      const T beta = (T)vc->beta;
      for (long long rw = (long long)blockIdx.x * wpb + wid; rw < rows; rw += (long long)gridDim.x * wpb) {
        const long long l0 = lin3(g, 0, 2 + (int)(rw % ny), (D == 3) ? kz0 + (int)(rw / ny) : 1);
        for (int xc = 2 + lane; xc <= g.n[0] - 1; xc += 32) {
          const long long l = l0 + xc;
          eps[l] = beta * eps[l] + z[l];
        }
      }
    }
    grid.sync();
  }
}

// myproject!: u[I,i] -= L[I,i]·∂(i,I,x) on inside(x)                                                      (flow.jl:331-333)
template <class T, int D> __global__ void __launch_bounds__(256) pois_apply_kernel(T* __restrict__ u, const T* __restrict__ L, const T* __restrict__ x,
                                                                                   const Geo g, int kz0, int kz1) {
  IFADV_POIS_ROWS({
    const T xc_ = __ldg(x + l);
_Pragma("unroll")
    for (int i = 0; i < D; ++i) {
      const long long st = (i == 0) ? 1 : ((i == 1) ? g.s1 : g.s2);
      const long long li = (long long)i * g.S + l;
      u[li] = u[li] - __ldg(L + li) * (xc_ - __ldg(x + l - st));
    }
  })
}

template <class T> __global__ void scale_kernel(T* a, T s, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = a[i] * s;
}

}  // namespace ifadv
