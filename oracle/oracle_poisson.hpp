// oracle/oracle_poisson.hpp -- TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).
// Variable-coefficient pressure projection: the reference's Jacobi-preconditioned conjugate-gradient solver
// psolver! (src/flow.jl:300-326), inproject! (:343-347) and myproject! (:328-341), on WaterLily's `Poisson` struct
// (fields L, D, iD, x, ϵ, r, z, perdir).  One loop per `@loop` / `@inside`, same order, same arrays as the reference.
//
// WaterLily.jl (Project.toml:15, compat "1.8", no Manifest) is NOT under /root/reference; the primitives psolver! calls --
// set_diag! / update!, diag, mult (multL, multU), perBC!, residual!, L₂ -- are restated from the published WaterLily 1.x
// src/Poisson.jl and are consistent with every call site in flow.jl:300-347.
// Parity status: PARITY UNPINNED -- the reference holds no known-answer test for the projection (test/maintests.jl exercises it only
// through the two integration tests, :193-217).  tests/test_oracle_poisson.py adds self-derived properties: A·x = z to the solver
// tolerance against a dense solve, symmetry / negative semi-definiteness of the operator, a divergence-free field after myproject!,
// the iteration-count rule of the loop condition, periodic wrap.
// Dot products: the reference's `⋅` is BLAS / CUBLAS dot (summation order unspecified); here a sequential Float64 accumulation over
// inside(x), rounded to T -- every comparison against it carries a tolerance.
#pragma once
#include <limits>

#include "oracle_fields.hpp"

namespace orc {

template <class T> struct Pois {  // WaterLily.Poisson
  VF<T> L;                        // lower-face coefficients (Ng..., D)  (Flow.μ₀)
  SF<T> D, iD, x, eps, r, z;      // diagonal, 1/diagonal, solution (Flow.p), increment, residual, source (Flow.σ)
  unsigned perdir;
};

// diag(I,L) = -Σᵢ (L[I,i] + L[I+δᵢ,i])
template <class T> inline T pois_diag(const Grid& g, const I3& I, const VF<T>& L) {
  T s = 0;
  for (int i = 0; i < g.D; ++i) s -= L(I, i) + L(sh(I, i, +1), i);
  return s;
}
// mult(I,L,D,x) = x[I]·D[I] + Σᵢ L[I,i]·x[I-δᵢ] + Σᵢ L[I+δᵢ,i]·x[I+δᵢ]
template <class T> inline T pois_mult(const Grid& g, const I3& I, const VF<T>& L, const SF<T>& D, const SF<T>& x) {
  T lo = 0, up = 0;
  for (int i = 0; i < g.D; ++i) lo += L(I, i) * x(sh(I, i, -1));
  for (int i = 0; i < g.D; ++i) up += L(sh(I, i, +1), i) * x(sh(I, i, +1));
  return x(I) * D(I) + lo + up;
}
// set_diag!(D,iD,L) = update!(p::Poisson)
template <class T> void pois_update(const Grid& g, const Pois<T>& p) {
  loop(r_inside(g), [&](I3 I) { p.D(I) = pois_diag(g, I, p.L); });
  loop(r_inside(g), [&](I3 I) { p.iD(I) = (p.D(I) * p.D(I) < 2 * std::numeric_limits<T>::epsilon()) ? T(0) : T(1) / p.D(I); });
}
// perBC!(a,perdir): periodic directions only, in order
template <class T> void perBC(const Grid& g, const SF<T>& a, unsigned perdir) {
  for (int j = 0; j < g.D; ++j) {
    if (!isper(perdir, j)) continue;
    const int64_t Nj = g.n[j];
    loop(r_slice(g, 1, j), [&](I3 I) { a(I) = a(CIj(j, I, Nj - 1)); });
    loop(r_slice(g, Nj, j), [&](I3 I) { a(I) = a(CIj(j, I, 2)); });
  }
}
// Σ over inside of a·b, accumulated in Float64, rounded to T
template <class T> T pois_dot(const Grid& g, const SF<T>& a, const SF<T>& b) {
  const Range r = r_inside(g);
  double s = 0;
#ifdef _OPENMP
#pragma omp parallel for collapse(2) reduction(+ : s) schedule(static)
#endif
  for (int64_t k = r.lo[2]; k <= r.hi[2]; ++k)
    for (int64_t j = r.lo[1]; j <= r.hi[1]; ++j)
      for (int64_t i = r.lo[0]; i <= r.hi[0]; ++i) {
        const I3 I{{i, j, k}};
        s += (double)a(I) * (double)b(I);
      }
  return (T)s;
}
template <class T> T pois_sum(const Grid& g, const SF<T>& a) {
  const Range r = r_inside(g);
  double s = 0;
#ifdef _OPENMP
#pragma omp parallel for collapse(2) reduction(+ : s) schedule(static)
#endif
  for (int64_t k = r.lo[2]; k <= r.hi[2]; ++k)
    for (int64_t j = r.lo[1]; j <= r.hi[1]; ++j)
      for (int64_t i = r.lo[0]; i <= r.hi[0]; ++i) s += (double)a(I3{{i, j, k}});
  return (T)s;
}
// residual!(p): r = z - A x (0 where iD == 0), mean removed when it is not round-off
template <class T> void pois_residual(const Grid& g, const Pois<T>& p) {
  perBC(g, p.x, p.perdir);
  loop(r_inside(g), [&](I3 I) { p.r(I) = (p.iD(I) == T(0)) ? T(0) : p.z(I) - pois_mult(g, I, p.L, p.D, p.x); });
  const Range ri = r_inside(g);
  int64_t cnt = 1;
  for (int d = 0; d < g.D; ++d) cnt *= ri.hi[d] - ri.lo[d] + 1;
  const T s = pois_sum(g, p.r) / (T)cnt;
  if (std::abs(s) <= 2 * std::numeric_limits<T>::epsilon()) return;
  loop(r_inside(g), [&](I3 I) { p.r(I) = p.r(I) - s; });
}
// psolver!(p;tol,itmx), src/flow.jl:300-326.  Returns nᵖ; *r2_out = the last r₂
template <class T> int psolver(const Grid& g, const Pois<T>& p, T tol, int itmx, double* r2_out) {
  perBC(g, p.x, p.perdir);                                                   // :301
  pois_residual(g, p);                                                       // :302
  T r2 = pois_dot(g, p.r, p.r);
  int np = 0;
  loop(r_inside(g), [&](I3 I) { p.z(I) = p.eps(I) = p.r(I) * p.iD(I); });     // :305
  T rho = pois_dot(g, p.r, p.z);                                             // :307
  while ((r2 > tol || (r2 > tol / 4 && np == 0)) && np < itmx) {             // :309
    perBC(g, p.eps, p.perdir);                                               // :311
    loop(r_inside(g), [&](I3 I) { p.z(I) = pois_mult(g, I, p.L, p.D, p.eps); });  // :312
    const T alpha = rho / pois_dot(g, p.z, p.eps);                           // :313
    loop(r_inside(g), [&](I3 I) {                                            // :314-315
      p.x(I) += alpha * p.eps(I);
      p.r(I) -= alpha * p.z(I);
    });
    loop(r_inside(g), [&](I3 I) { p.z(I) = p.r(I) * p.iD(I); });              // :316
    const T rho2 = pois_dot(g, p.r, p.z);                                    // :317
    const T beta = rho2 / rho;                                               // :318
    loop(r_inside(g), [&](I3 I) { p.eps(I) = beta * p.eps(I) + p.z(I); });    // :319
    rho = rho2;
    r2 = pois_dot(g, p.r, p.r);                                              // :321
    ++np;
  }
  perBC(g, p.x, p.perdir);                                                   // :325
  if (r2_out) *r2_out = (double)r2;
  return np;
}
// inproject!(a,b::Poisson,dt), src/flow.jl:343-347
template <class T> int inproject(const Grid& g, const VF<T>& u, const Pois<T>& p, T dt, double* r2_out) {
  for (int64_t k = 0; k < g.S; ++k) { p.z.p[k] = 0; p.eps.p[k] = 0; p.r.p[k] = 0; }
  loop(r_inside(g), [&](I3 I) {  // z = div(I,u)
    T s = 0;
    for (int i = 0; i < g.D; ++i) s += u(sh(I, i, +1), i) - u(I, i);
    p.z(I) = s;
  });
  for (int64_t k = 0; k < g.S; ++k) p.x.p[k] *= dt;
  return psolver(g, p, T(50) * std::numeric_limits<T>::epsilon(), 2000, r2_out);
}
// myproject!(a,b,w) with dt = T(w)·last(a.Δt) formed by the caller, src/flow.jl:328-341
template <class T> int myproject(const Grid& g, const VF<T>& u, const Pois<T>& p, T dt, double* r2_out) {
  const int np = inproject(g, u, p, dt, r2_out);
  for (int i = 0; i < g.D; ++i) loop(r_inside(g), [&](I3 I) { u(I, i) -= p.L(I, i) * (p.x(I) - p.x(sh(I, i, -1))); });
  const T idt = T(1) / dt;
  for (int64_t k = 0; k < g.S; ++k) p.x.p[k] *= idt;
  return np;
}

}  // namespace orc
