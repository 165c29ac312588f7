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

}  // namespace ifadv
