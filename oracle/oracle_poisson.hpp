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

// ================================================================================================================================
// WaterLily.MultiLevelPoisson (geometric multigrid: V-cycles with Jacobi pre-smoothing and pcg! smoothing), the solver behind
// inproject!(a,b::MultiLevelPoisson,dt) = solver!(b;tol=1e-4,itmx=200) (src/flow.jl:343-347) and WaterLily's default `psolver`.
// Restated from the published WaterLily 1.x src/MultiLevelPoisson.jl and src/Poisson.jl (restrictL / restrict / prolongate, up / down,
// divisible, restrictML, update!, Vcycle!, solver!, Jacobi!, increment!, pcg!).  PARITY UNPINNED like the rest of this header.
// ================================================================================================================================
#include <memory>
#include <vector>

namespace orc {

template <class T> struct MLLevel {
  Grid g;
  std::vector<T> oL, oD, oiD, ox, oeps, orr, oz;  // storage the level owns (level 1 borrows x, L, z from the caller)
  Pois<T> p;
};
template <class T> struct MLPois {
  std::vector<std::unique_ptr<MLLevel<T>>> lv;
  unsigned perdir;
  std::vector<int> n;
};

inline bool ml_divisible(const Grid& g) {  // divisible(N) = mod(N,2)==0 && N>4 on every extent of x
  for (int d = 0; d < g.D; ++d)
    if (g.n[d] % 2 != 0 || g.n[d] <= 4) return false;
  return true;
}
// restrictL!(a,b;perdir): a[I,i] = 0.5·Σ_{J ∈ up(I,i)} b[J,i] on inside(a), up(I,i) = (2I-2):(2I-1-δᵢ); BC!(a,0,false,perdir)
template <class T> void ml_restrictL(const Grid& gc, const VF<T>& a, const Grid& gf, const VF<T>& b, unsigned perdir) {
  for (int i = 0; i < gc.D; ++i)
    loop(r_inside(gc), [&](I3 I) {
      T s = 0;
      const int64_t h2 = (gc.D == 3) ? 1 : 0;
      for (int64_t c2 = 0; c2 <= ((i == 2) ? 0 : h2); ++c2)
        for (int64_t c1 = 0; c1 <= ((i == 1) ? 0 : 1); ++c1)
          for (int64_t c0 = 0; c0 <= ((i == 0) ? 0 : 1); ++c0) {
            I3 J{{2 * I.i[0] - 2 + c0, 2 * I.i[1] - 2 + c1, (gc.D == 3) ? 2 * I.i[2] - 2 + c2 : 1}};
            s += b(J, i);
          }
      a(I, i) = T(0.5) * s;
    });
  T Z[3] = {0, 0, 0};
  BC_vec<T>(gc, a, Z, false, perdir);
}
// restrict!(a,b): a[I] = Σ_{J ∈ up(I)} b[J] on inside(a)
template <class T> void ml_restrict(const Grid& gc, const SF<T>& a, const SF<T>& b) {
  loop(r_inside(gc), [&](I3 I) {
    T s = 0;
    for (int64_t c2 = 0; c2 <= ((gc.D == 3) ? 1 : 0); ++c2)
      for (int64_t c1 = 0; c1 <= 1; ++c1)
        for (int64_t c0 = 0; c0 <= 1; ++c0) s += b(I3{{2 * I.i[0] - 2 + c0, 2 * I.i[1] - 2 + c1, (gc.D == 3) ? 2 * I.i[2] - 2 + c2 : 1}});
    a(I) = s;
  });
}
// prolongate!(a,b): a[I] = b[down(I)] on inside(a), down(I) = (I+2)÷2
template <class T> void ml_prolongate(const Grid& gf, const SF<T>& a, const SF<T>& b) {
  loop(r_inside(gf), [&](I3 I) { a(I) = b(I3{{(I.i[0] + 2) / 2, (I.i[1] + 2) / 2, (gf.D == 3) ? (I.i[2] + 2) / 2 : 1}}); });
}
// increment!(p): perBC!(ϵ); r -= Aϵ; x += ϵ
template <class T> void pois_increment(const Grid& g, const Pois<T>& p) {
  perBC(g, p.eps, p.perdir);
  loop(r_inside(g), [&](I3 I) {
    p.r(I) = p.r(I) - pois_mult(g, I, p.L, p.D, p.eps);
    p.x(I) = p.x(I) + p.eps(I);
  });
}
// Jacobi!(p;it=1)
template <class T> void pois_jacobi(const Grid& g, const Pois<T>& p) {
  loop(r_inside(g), [&](I3 I) { p.eps(I) = p.r(I) * p.iD(I); });
  pois_increment(g, p);
}
// smooth!(p) = pcg!(p;it=6)
template <class T> void pois_pcg_smooth(const Grid& g, const Pois<T>& p, int it = 6) {
  const T eps10 = 10 * std::numeric_limits<T>::epsilon();
  loop(r_inside(g), [&](I3 I) { p.z(I) = p.eps(I) = p.r(I) * p.iD(I); });
  T rho = pois_dot(g, p.r, p.z);
  if (std::abs(rho) < eps10) return;
  for (int i = 1; i <= it; ++i) {
    perBC(g, p.eps, p.perdir);
    loop(r_inside(g), [&](I3 I) { p.z(I) = pois_mult(g, I, p.L, p.D, p.eps); });
    const T alpha = rho / pois_dot(g, p.z, p.eps);
    if (std::abs((double)alpha) < 1e-2 || std::abs((double)alpha) > 1e3) return;  // alpha should be O(1); Float64 literals in the reference
    loop(r_inside(g), [&](I3 I) {
      p.x(I) += alpha * p.eps(I);
      p.r(I) -= alpha * p.z(I);
    });
    if (i == it) return;
    loop(r_inside(g), [&](I3 I) { p.z(I) = p.r(I) * p.iD(I); });
    const T rho2 = pois_dot(g, p.r, p.z);
    if (std::abs(rho2) < eps10) return;
    const T beta = rho2 / rho;
    loop(r_inside(g), [&](I3 I) { p.eps(I) = beta * p.eps(I) + p.z(I); });
    rho = rho2;
  }
}
// MultiLevelPoisson(x,L,z;maxlevels=10,perdir): level 1 on the caller's arrays, coarser levels by restrictML while divisible
template <class T> MLPois<T>* ml_create(int D, const int64_t* Ng, T* x, T* L, T* z, unsigned perdir, int maxlevels) {
  auto* ml = new MLPois<T>();
  ml->perdir = perdir;
  auto mk = [&](const Grid& g, T* px, T* pL, T* pz) {
    auto lv = std::make_unique<MLLevel<T>>();
    lv->g = g;
    const size_t S = (size_t)g.S;
    if (!pL) { lv->oL.assign(S * g.D, T(0)); pL = lv->oL.data(); }
    if (!px) { lv->ox.assign(S, T(0)); px = lv->ox.data(); }
    if (!pz) { lv->oz.assign(S, T(0)); pz = lv->oz.data(); }
    lv->oD.assign(S, T(0)); lv->oiD.assign(S, T(0)); lv->oeps.assign(S, T(0)); lv->orr.assign(S, T(0));
    const Grid* gp = &lv->g;
    lv->p = Pois<T>{VF<T>{pL, gp}, SF<T>{lv->oD.data(), gp}, SF<T>{lv->oiD.data(), gp}, SF<T>{px, gp}, SF<T>{lv->oeps.data(), gp},
                    SF<T>{lv->orr.data(), gp}, SF<T>{pz, gp}, perdir};
    return lv;
  };
  ml->lv.push_back(mk(make_grid(D, Ng), x, L, z));
  pois_update(ml->lv[0]->g, ml->lv[0]->p);
  while (ml_divisible(ml->lv.back()->g) && (int)ml->lv.size() <= maxlevels) {  // restrictML
    const Grid& gf = ml->lv.back()->g;
    int64_t Na[3] = {1 + gf.n[0] / 2, 1 + gf.n[1] / 2, (D == 3) ? 1 + gf.n[2] / 2 : 1};
    auto c = mk(make_grid(D, Na), nullptr, nullptr, nullptr);
    ml_restrictL(c->g, c->p.L, gf, ml->lv.back()->p.L, perdir);
    pois_update(c->g, c->p);
    ml->lv.push_back(std::move(c));
  }
  return ml;
}
// update!(ml)
template <class T> void ml_update(MLPois<T>& ml) {
  pois_update(ml.lv[0]->g, ml.lv[0]->p);
  for (size_t l = 1; l < ml.lv.size(); ++l) {
    ml_restrictL(ml.lv[l]->g, ml.lv[l]->p.L, ml.lv[l - 1]->g, ml.lv[l - 1]->p.L, ml.perdir);
    pois_update(ml.lv[l]->g, ml.lv[l]->p);
  }
}
// Vcycle!(ml;l)   (l 0-based here)
template <class T> void ml_vcycle(MLPois<T>& ml, size_t l = 0) {
  MLLevel<T>&fine = *ml.lv[l], &coarse = *ml.lv[l + 1];
  pois_jacobi(fine.g, fine.p);
  ml_restrict(coarse.g, coarse.p.r, fine.p.r);
  for (int64_t k = 0; k < coarse.g.S; ++k) coarse.p.x.p[k] = 0;
  if (l + 2 < ml.lv.size()) ml_vcycle(ml, l + 1);
  pois_pcg_smooth(coarse.g, coarse.p);
  ml_prolongate(fine.g, fine.p.eps, coarse.p.x);
  pois_increment(fine.g, fine.p);
}
// solver!(ml;tol=1e-4,itmx=32)
template <class T> int ml_solver(MLPois<T>& ml, T tol, int itmx, double* r2_out) {
  MLLevel<T>& f = *ml.lv[0];
  pois_residual(f.g, f.p);
  T r2 = pois_dot(f.g, f.p.r, f.p.r);
  int np = 0;
  while (np < itmx) {
    ml_vcycle(ml);
    pois_pcg_smooth(f.g, f.p);
    r2 = pois_dot(f.g, f.p.r, f.p.r);
    ++np;
    if (r2 < tol) break;
  }
  perBC(f.g, f.p.x, f.p.perdir);
  ml.n.push_back(np);
  if (r2_out) *r2_out = (double)r2;
  return np;
}
// myproject!(a,b::MultiLevelPoisson,w): inproject! = z ← ∇·u, x ← x·dt, solver!(b;tol=1e-4,itmx=200)   (src/flow.jl:328-341,343-347)
template <class T> int ml_myproject(MLPois<T>& ml, const VF<T>& u, T dt, double* r2_out) {
  MLLevel<T>& f = *ml.lv[0];
  const Grid& g = f.g;
  loop(r_inside(g), [&](I3 I) {
    T s = 0;
    for (int i = 0; i < g.D; ++i) s += u(sh(I, i, +1), i) - u(I, i);
    f.p.z(I) = s;
  });
  for (int64_t k = 0; k < g.S; ++k) f.p.x.p[k] *= dt;
  const int np = ml_solver(ml, T(1e-4), 200, r2_out);
  for (int i = 0; i < g.D; ++i) loop(r_inside(g), [&](I3 I) { u(I, i) -= f.p.L(I, i) * (f.p.x(I) - f.p.x(sh(I, i, -1))); });
  const T idt = T(1) / dt;
  for (int64_t k = 0; k < g.S; ++k) f.p.x.p[k] *= idt;
  return np;
}

}  // namespace orc
