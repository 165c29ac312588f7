// ifadv_mlpoisson.cuh -- kernels of WaterLily's geometric multigrid MultiLevelPoisson, the solver behind the second method of the
// reference's inproject! (src/flow.jl:343-347: z <- div u; x <- x dt; solver!(b;tol=1e-4,itmx=200)) and WaterLily's default `psolver`.
// WaterLily is not under /root/reference; restrictL! / restrict! / prolongate! / Jacobi! / increment! / pcg! / Vcycle! / solver! are
// restated from its published 1.x sources (MultiLevelPoisson.jl, Poisson.jl), like the primitives of ifadv_poisson.cuh.
//
// A V-cycle visits every level twice and smooths with six pcg! iterations per level; on all but the finest levels that is launch
// latency, not bandwidth.  So, as in ifadv_poisson.cuh, NO scalar of the smoother ever goes to the host: rho, alpha, beta and pcg!'s
// early exits (|rho| < 10eps, alpha outside [1e-2, 1e3]) live in the level's control block, written by the last CTA of the kernel that
// finishes the reduction, and the kernels behind an early exit return at once (`live`).  The launch sequence of one
// Vcycle! + smooth! + L2 is therefore data-independent and is replayed as ONE CUDA graph per solver cycle (ifadv_mlpoisson.cu).
// Per-cell arithmetic follows the reference expression by expression (-fmad=false, IEEE division); dot products are Float64 partial
// sums in a fixed order, rounded to T.
#pragma once
#include "ifadv_poisson.cuh"

namespace ifadv {

// restrictL!(a,b): a[I,i] = 0.5 Σ_{J ∈ up(I,i)} b[J,i] on inside(a), up(I,i) = (2I-2):(2I-1-δᵢ) -- the fine faces that tile the coarse face
template <class T, int D> __global__ void __launch_bounds__(256) ml_restrictL_kernel(T* __restrict__ a, const Geo g, const T* __restrict__ b,
                                                                                      const Geo gf, int kz0, int kz1) {
  IFADV_POIS_ROWS({
_Pragma("unroll")
    for (int i = 0; i < D; ++i) {
      T s = T(0);
      const int h2 = (D == 3 && i != 2) ? 1 : 0, h1 = (i != 1) ? 1 : 0, h0 = (i != 0) ? 1 : 0;
      for (int c2 = 0; c2 <= h2; ++c2)
        for (int c1 = 0; c1 <= h1; ++c1)
          for (int c0 = 0; c0 <= h0; ++c0)
            s = s + __ldg(b + (long long)i * gf.S + lin3(gf, 2 * xc - 2 + c0, 2 * y - 2 + c1, (D == 3) ? 2 * zc - 2 + c2 : 1));
      a[(long long)i * g.S + l] = T(0.5) * s;
    }
  })
}

// restrict!(coarse.r, fine.r): Σ over the 2^D children; fill!(coarse.x, 0) (its ghost entries are never written: they stay 0)
template <class T, int D> __global__ void __launch_bounds__(256) ml_restrict_kernel(T* __restrict__ rc, T* __restrict__ xcoarse, const Geo g,
                                                                                     const T* __restrict__ rf, const Geo gf, int kz0, int kz1) {
  IFADV_POIS_ROWS({
    T s = T(0);
    for (int c2 = 0; c2 <= ((D == 3) ? 1 : 0); ++c2)
      for (int c1 = 0; c1 <= 1; ++c1)
        for (int c0 = 0; c0 <= 1; ++c0) s = s + __ldg(rf + lin3(gf, 2 * xc - 2 + c0, 2 * y - 2 + c1, (D == 3) ? 2 * zc - 2 + c2 : 1));
    rc[l] = s;
    xcoarse[l] = T(0);
  })
}

// prolongate!(fine.ϵ, coarse.x): ϵ[I] = x[down(I)], down(I) = (I+2)÷2, on inside(fine)
template <class T, int D> __global__ void __launch_bounds__(256) ml_prolongate_kernel(T* __restrict__ eps, const Geo g, const T* __restrict__ xcoarse,
                                                                                       const Geo gc, int kz0, int kz1) {
  IFADV_POIS_ROWS({ eps[l] = __ldg(xcoarse + lin3(gc, (xc + 2) >> 1, (y + 2) >> 1, (D == 3) ? (zc + 2) >> 1 : 1)); })
}

// Jacobi!, first half: ϵ = r·iD on inside
template <class T, int D> __global__ void __launch_bounds__(256) ml_jacobi_kernel(T* __restrict__ eps, const T* __restrict__ r, const T* __restrict__ iD,
                                                                                   const Geo g, int kz0, int kz1) {
  constexpr int U = (sizeof(T) == 4) ? 8 : 4;
  IFADV_POIS_ROWS_U(U, {
    T rv[U], dv[U];
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) { rv[k] = __ldg(r + l0 + xc); dv[k] = __ldg(iD + l0 + xc); }
    }
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) eps[l0 + xc] = rv[k] * dv[k];
    }
  })
}

// increment!(p) behind perBC!(ϵ): r -= A ϵ; x += ϵ on inside
template <class T, int D> __global__ void __launch_bounds__(256) ml_increment_kernel(T* __restrict__ x, T* __restrict__ r, const T* __restrict__ eps,
                                                                                      const T* __restrict__ L, const T* __restrict__ Dg, const Geo g,
                                                                                      int kz0, int kz1) {
  constexpr int U = (sizeof(T) == 4) ? 4 : 2;
  IFADV_POIS_ROWS_U(U, {
    T ec[U], dg[U], rv[U], xv[U], ll[U][D], lu[U][D], em[U][D], ep[U][D];
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) {
        const long long l = l0 + xc;
        ec[k] = __ldg(eps + l);
        dg[k] = __ldg(Dg + l);
        rv[k] = r[l];
        xv[k] = x[l];
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
        r[l0 + xc] = rv[k] - (ec[k] * dg[k] + lo + up);
        x[l0 + xc] = xv[k] + ec[k];
      }
    }
  })
}

// residual!, second half: r -= s on inside when |s| > 2eps (s from pois_residual_kernel); the r2 of the control block is now stale
template <class T, int D> __global__ void __launch_bounds__(256) ml_submean_kernel(T* __restrict__ r, const Geo g, PoisCtl* ctl, int kz0, int kz1) {
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->r2_valid = 0;
  if (!ctl->sub_mean) return;
  const T s = (T)ctl->mean;
  IFADV_POIS_ROWS({ r[l] = r[l] - s; })
}

// ---- pcg!(p;it=6), the smoother ----------------------------------------------------------------------------------------------------
// z = ϵ = r·iD on inside; rho = r·z; |rho| < 10eps: return
template <class T, int D> __global__ void __launch_bounds__(256) ml_pcg_start_kernel(T* __restrict__ z, T* __restrict__ eps, const T* __restrict__ r,
                                                                                      const T* __restrict__ iD, const Geo g, PoisCtl* ctl, int kz0,
                                                                                      int kz1) {
  constexpr int U = (sizeof(T) == 4) ? 8 : 4;
  double acc[1] = {0.0};
  IFADV_POIS_ROWS_U(U, {
    T rv[U], dv[U];
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) { rv[k] = __ldg(r + l0 + xc); dv[k] = __ldg(iD + l0 + xc); }
    }
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) {
        const T zv = rv[k] * dv[k];
        z[l0 + xc] = zv;
        eps[l0 + xc] = zv;
        acc[0] += (double)rv[k] * (double)zv;
      }
    }
  })
  double tot[1];
  if (grid_reduce<1>(acc, ctl, 0, tot) && threadIdx.x == 0) {
    const T rho = (T)tot[0];
    ctl->rho = (double)rho;
    ctl->live = (t_abs(rho) < T(10) * teps<T>::v) ? 0 : 1;
    ctl->r2_valid = 0;
  }
}

// z = A ϵ on inside; alpha = rho / (z·ϵ); alpha outside [1e-2, 1e3]: return ("alpha should be O(1)")
template <class T, int D> __global__ void __launch_bounds__(256) ml_pcg_mult_kernel(T* __restrict__ z, const T* __restrict__ eps, const T* __restrict__ L,
                                                                                     const T* __restrict__ Dg, const Geo g, PoisCtl* ctl, int kz0,
                                                                                     int kz1) {
  if (!ctl->live) return;
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
    const T alpha = (T)ctl->rho / (T)tot[0];
    ctl->alpha = (double)alpha;
    const double aa = fabs((double)alpha);  // the reference compares against the Float64 literals 1e-2 and 1e3
    if (aa < 1e-2 || aa > 1e3) ctl->live = 0;
  }
}

// x += alpha ϵ; r -= alpha z; r2 = r·r.  Not the last iteration: z = r·iD, rho2 = r·z; |rho2| < 10eps: return; beta = rho2/rho; rho = rho2
template <class T, int D> __global__ void __launch_bounds__(256) ml_pcg_update_kernel(T* __restrict__ x, T* __restrict__ r, T* __restrict__ z,
                                                                                       const T* __restrict__ eps, const T* __restrict__ iD,
                                                                                       const Geo g, PoisCtl* ctl, int last, int kz0, int kz1) {
  if (!ctl->live) return;
  const T alpha = (T)ctl->alpha;
  constexpr int U = (sizeof(T) == 4) ? 8 : 4;
  double acc[2] = {0.0, 0.0};
  IFADV_POIS_ROWS_U(U, {
    T xv[U], ev[U], rv[U], zv[U], dv[U];
_Pragma("unroll")
    for (int k = 0; k < U; ++k) {
      const int xc = xb + 32 * k;
      if (xc <= xlast) {
        const long long l = l0 + xc;
        xv[k] = x[l]; ev[k] = __ldg(eps + l); rv[k] = r[l]; zv[k] = z[l];
        dv[k] = last ? T(0) : __ldg(iD + l);
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
        acc[1] += (double)rn * (double)rn;
        if (!last) {
          const T zn = rn * dv[k];
          z[l] = zn;
          acc[0] += (double)rn * (double)zn;
        }
      }
    }
  })
  double tot[2];
  if (grid_reduce<2>(acc, ctl, 3, tot) && threadIdx.x == 0) {
    ctl->r2 = (double)(T)tot[1];
    ctl->r2_valid = 1;
    if (last) {
      ctl->live = 0;
    } else {
      const T rho2 = (T)tot[0];
      if (t_abs(rho2) < T(10) * teps<T>::v) {
        ctl->live = 0;
      } else {
        ctl->beta = (double)(rho2 / (T)ctl->rho);
        ctl->rho = (double)rho2;
      }
    }
  }
}

// ϵ = beta ϵ + z on inside
template <class T, int D> __global__ void __launch_bounds__(256) ml_pcg_dir_kernel(T* __restrict__ eps, const T* __restrict__ z, const Geo g,
                                                                                    const PoisCtl* ctl, int kz0, int kz1) {
  if (!ctl->live) return;
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

// L2(p) = r·r when no update kernel has left it behind (pcg! returned before its first update)
template <class T, int D> __global__ void __launch_bounds__(256) ml_r2_kernel(const T* __restrict__ r, const Geo g, PoisCtl* ctl, int kz0, int kz1) {
  if (ctl->r2_valid) return;
  double acc[1] = {0.0};
  IFADV_POIS_ROWS({
    const T rv = __ldg(r + l);
    acc[0] += (double)rv * (double)rv;
  })
  double tot[1];
  if (grid_reduce<1>(acc, ctl, 1, tot) && threadIdx.x == 0) {
    ctl->r2 = (double)(T)tot[0];
    ctl->r2_valid = 1;
  }
}

// ---- the bottom of the V-cycle in ONE CTA -------------------------------------------------------------------------------------------
// Levels of a few thousand cells are pure launch latency in the multi-kernel form (~35 launches of ~2.5 us per level and cycle, for
// microseconds of work).  ml_bottom_kernel runs "[Vcycle!(l = b)]; smooth!(level b)" for all levels from b down to the coarsest in a
// single CTA of 1024 threads: __syncthreads() stands where the kernel boundaries were, the smoother's scalars are block-uniform
// registers, dot products are block reductions in Float64 (fixed order).  Same per-cell arithmetic as the kernels above.  Arrays
// written inside the launch (x, ϵ, r, z) are read with plain loads; L, D, iD are constant here and go through the read-only path.
#define IFADV_ML_BOTTOM_MAX 8
template <class T> struct MLDev {
  Geo g;
  const T *L, *D, *iD;
  T *x, *eps, *r, *z;
};
template <class T> struct MLBottom {
  MLDev<T> lv[IFADV_ML_BOTTOM_MAX];
  int n;
  unsigned per;
};

template <int D, class F> IFADV_DI void bt_inside(const Geo& g, F&& fn) {
  const int nx = g.n[0] - 2, ny = g.n[1] - 2, nz = (D == 3) ? g.n[2] - 2 : 1;
  const int tot = nx * ny * nz;
  for (int t = threadIdx.x; t < tot; t += blockDim.x) {
    const int ix = t % nx, q = t / nx, iy = q % ny, iz = q / ny;
    fn(2 + ix, 2 + iy, (D == 3) ? 2 + iz : 1, lin3(g, 2 + ix, 2 + iy, (D == 3) ? 2 + iz : 1));
  }
}
// block sum, returned to every thread; two barriers inside, so everything written before the call is visible after it
IFADV_DI double bt_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = (lane < nw) ? sh[lane] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  const double r = sh[32];
  __syncthreads();
  return r;
}
// perBC!(a,perdir) by the whole CTA (closed form of the sequential plane copies, as perbc_kernel); barrier before and after
template <class T, int D> IFADV_DI void bt_perbc(T* a, const Geo& g, unsigned per) {
  if (!per) return;
  __syncthreads();
  const long long n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const long long c0 = (per & 1u) ? 2 * n1 * n2 : 0, c1 = (per & 2u) ? 2 * n0 * n2 : 0, c2 = (D == 3 && (per & 4u)) ? 2 * n0 * n1 : 0;
  for (long long t = threadIdx.x; t < c0 + c1 + c2; t += blockDim.x) {
    int xx, yy, zz;
    if (t < c0) { const long long q = t >> 1; xx = (t & 1) ? (int)n0 : 1; yy = (int)(q % n1) + 1; zz = (int)(q / n1) + 1; }
    else if (t < c0 + c1) { const long long u_ = t - c0, q = u_ >> 1; yy = (u_ & 1) ? (int)n1 : 1; xx = (int)(q % n0) + 1; zz = (int)(q / n0) + 1; }
    else { const long long u_ = t - c0 - c1, q = u_ >> 1; zz = (u_ & 1) ? (int)n2 : 1; xx = (int)(q % n0) + 1; yy = (int)(q / n0) + 1; }
    const int mx = (per & 1u) ? wrapc(xx, g.n[0]) : xx, my = (per & 2u) ? wrapc(yy, g.n[1]) : yy, mz = (D == 3 && (per & 4u)) ? wrapc(zz, g.n[2]) : zz;
    a[lin3(g, xx, yy, zz)] = a[lin3(g, mx, my, mz)];
  }
  __syncthreads();
}
// increment!(p): perBC!(ϵ); r -= Aϵ; x += ϵ
template <class T, int D> IFADV_DI void bt_increment(const MLDev<T>& v, unsigned per) {
  __syncthreads();
  bt_perbc<T, D>(v.eps, v.g, per);
  bt_inside<D>(v.g, [&](int, int, int, long long l) {
    v.r[l] = v.r[l] - pois_mult_plain<T, D>(v.L, v.D, v.eps, v.g, l);
    v.x[l] = v.x[l] + v.eps[l];
  });
  __syncthreads();
}
// smooth!(p) = pcg!(p;it=6); block-uniform control flow
template <class T, int D> IFADV_DI void bt_smooth(const MLDev<T>& v, unsigned per, double* sh) {
  double acc = 0.0;
  bt_inside<D>(v.g, [&](int, int, int, long long l) {
    const T rv = v.r[l], zv = rv * __ldg(v.iD + l);
    v.z[l] = zv;
    v.eps[l] = zv;
    acc += (double)rv * (double)zv;
  });
  T rho = (T)bt_sum(acc, sh);
  if (t_abs(rho) < T(10) * teps<T>::v) return;
  for (int i = 1; i <= 6; ++i) {
    bt_perbc<T, D>(v.eps, v.g, per);
    acc = 0.0;
    bt_inside<D>(v.g, [&](int, int, int, long long l) {
      const T w = pois_mult_plain<T, D>(v.L, v.D, v.eps, v.g, l);
      v.z[l] = w;
      acc += (double)w * (double)v.eps[l];
    });
    const T alpha = rho / (T)bt_sum(acc, sh);
    const double aa = fabs((double)alpha);
    if (aa < 1e-2 || aa > 1e3) return;
    acc = 0.0;
    const bool last = (i == 6);
    bt_inside<D>(v.g, [&](int, int, int, long long l) {
      v.x[l] = v.x[l] + alpha * v.eps[l];
      const T rn = v.r[l] - alpha * v.z[l];
      v.r[l] = rn;
      if (!last) {
        const T zn = rn * __ldg(v.iD + l);
        v.z[l] = zn;
        acc += (double)rn * (double)zn;
      }
    });
    if (last) { __syncthreads(); return; }
    const T rho2 = (T)bt_sum(acc, sh);
    if (t_abs(rho2) < T(10) * teps<T>::v) return;
    const T beta = rho2 / rho;
    bt_inside<D>(v.g, [&](int, int, int, long long l) { v.eps[l] = beta * v.eps[l] + v.z[l]; });
    rho = rho2;
    __syncthreads();
  }
}
template <class T, int D> __global__ void __launch_bounds__(1024) ml_bottom_kernel(const MLBottom<T> P) {
  __shared__ double sh[40];
  const int n = P.n;
  for (int k = 0; k + 1 < n; ++k) {  // down: Jacobi!(fine); restrict!(coarse.r, fine.r); fill!(coarse.x, 0)
    const MLDev<T>& f = P.lv[k];
    const MLDev<T>& c = P.lv[k + 1];
    bt_inside<D>(f.g, [&](int, int, int, long long l) { f.eps[l] = f.r[l] * __ldg(f.iD + l); });
    bt_increment<T, D>(f, P.per);
    bt_inside<D>(c.g, [&](int xc, int y, int zc, long long l) {
      T s = T(0);
      for (int c2 = 0; c2 <= ((D == 3) ? 1 : 0); ++c2)
        for (int c1 = 0; c1 <= 1; ++c1)
          for (int c0 = 0; c0 <= 1; ++c0) s = s + f.r[lin3(f.g, 2 * xc - 2 + c0, 2 * y - 2 + c1, (D == 3) ? 2 * zc - 2 + c2 : 1)];
      c.r[l] = s;
      c.x[l] = T(0);
    });
    __syncthreads();
  }
  bt_smooth<T, D>(P.lv[n - 1], P.per, sh);
  for (int k = n - 2; k >= 0; --k) {  // up: prolongate!(fine.ϵ, coarse.x); increment!(fine); then smooth! of that level (its parent's call)
    const MLDev<T>& f = P.lv[k];
    const MLDev<T>& c = P.lv[k + 1];
    __syncthreads();
    bt_inside<D>(f.g, [&](int xc, int y, int zc, long long l) { f.eps[l] = c.x[lin3(c.g, (xc + 2) >> 1, (y + 2) >> 1, (D == 3) ? (zc + 2) >> 1 : 1)]; });
    bt_increment<T, D>(f, P.per);
    bt_smooth<T, D>(f, P.per, sh);
  }
}

}  // namespace ifadv
