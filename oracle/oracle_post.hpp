// oracle/oracle_post.hpp -- TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).
// Post-processing: level-set redistancing (src/redistaning.jl:8-143) and the energy / momentum / enstrophy
// metrics (src/metrics.jl:15-50).  One loop per `@loop`, same order, same arrays as the reference.
//
// Parity status: the metrics' per-cell functions are PINNED by the reference's known answers (test/maintests.jl:321-345);
// computeL! and _redistaningStage! by its exactness tests (:264-292: L ≡ 0 on a signed-distance ramp for Neumann and periodic
// building blocks; a uniform-L field integrates like forward Euler) and redistaning! by the planar-interface test (:304-318)
// -- all in tests/test_oracle_post.py.
#pragma once
#include "oracle_fields.hpp"

namespace orc {

template <class T> inline T minmod2(T a, T b) { return (std::abs(a) <= std::abs(b)) ? a : b; }  // redistaning.jl:106
// 𝛁ϕᵢ²(a,b,c,d,e,s), redistaning.jl:119-143 (second-order ENO, Sussman et al. 1999)
template <class T> inline T gradphi2(T a, T b, T c, T d, T e, T s) {
  const T dp = d - c, dm = c - b;
  const T ddp = e + c - 2 * d, dd0 = d + b - 2 * c, ddm = c + a - 2 * b;
  const T dR = dp - minmod2(ddp, dd0) / 2;
  const T dL = dm + minmod2(dd0, ddm) / 2;
  const T wR = dR * s, wL = dL * s;
  if (wR < 0 && (wR + wL) < 0) return dR * dR;
  if (wL > 0 && (wR + wL) > 0) return dL * dL;
  return T(0);
}
// computeL!(L,ϕ,ϕini;perdir), redistaning.jl:67-87 with the boundary blocks :91-104
template <class T> void computeL(const Grid& g, const SF<T>& L, const SF<T>& phi, const SF<T>& pini, unsigned perdir) {
  for (int64_t k = 0; k < g.S; ++k) L.p[k] = 0;
  for (int i = 0; i < g.D; ++i) {
    const int64_t Ni = g.n[i];
    const bool per = isper(perdir, i);
    auto sg = [&](I3 I) { return sgn(pini(I)); };
    loop(r_slice(g, 2, i, 2), [&](I3 I) {  // lowerL!
      if (!inside_others(g, I, i)) return;
      const T a = per ? phi(CIj(i, I, Ni - 2)) : phi(sh(I, i, -1));
      L(I) += gradphi2(a, phi(sh(I, i, -1)), phi(I), phi(sh(I, i, +1)), phi(sh(I, i, +2)), sg(I));
    });
    Range r = r_inside(g);
    r.lo[i] = 3; r.hi[i] = Ni - 2;
    loop(r, [&](I3 I) { L(I) += gradphi2(phi(sh(I, i, -2)), phi(sh(I, i, -1)), phi(I), phi(sh(I, i, +1)), phi(sh(I, i, +2)), sg(I)); });
    loop(r_slice(g, Ni - 1, i, 2), [&](I3 I) {  // upperL!
      if (!inside_others(g, I, i)) return;
      const T e = per ? phi(CIj(i, I, 3)) : phi(sh(I, i, +1));
      L(I) += gradphi2(phi(sh(I, i, -2)), phi(sh(I, i, -1)), phi(I), phi(sh(I, i, +1)), e, sg(I));
    });
  }
  loop(r_inside(g), [&](I3 I) { L(I) = pini(I) * (1 - std::sqrt(L(I))); });
}
// _redistaningStage!(ϕ,ϕ⁰,ϕini,L,dτ,α;perdir), redistaning.jl:31-34
template <class T> void redistStage(const Grid& g, const SF<T>& phi, const SF<T>& phi0, const SF<T>& pini, const SF<T>& L, T dtau, T alpha, unsigned perdir) {
  computeL(g, L, phi, pini, perdir);
  loop(r_inside(g), [&](I3 I) { phi(I) = alpha * phi0(I) + (1 - alpha) * (phi(I) + dtau * L(I)); });
}
// redistaning!(ls; d, dτ, perdir), redistaning.jl:44-57 (third-order SSP Runge-Kutta in pseudo-time)
template <class T> void redistance(const Grid& g, const SF<T>& phi, const SF<T>& phi0, const SF<T>& pini, const SF<T>& L, double d, double dtau, unsigned perdir) {
  const int itmx = (int)std::nearbyint(d / dtau);
  for (int it = 0; it < itmx; ++it) {
    for (int64_t k = 0; k < g.S; ++k) phi0.p[k] = phi.p[k];
    redistStage(g, phi, phi0, pini, L, (T)dtau, T(0), perdir); BCf(g, phi, perdir);
    redistStage(g, phi, phi0, pini, L, (T)dtau, T(3.0 / 4), perdir); BCf(g, phi, perdir);
    redistStage(g, phi, phi0, pini, L, (T)dtau, T(1.0 / 3), perdir); BCf(g, phi, perdir);
  }
}

// ---- metrics, src/metrics.jl ---------------------------------------------------------------------------------------------------
template <class T> inline double rhokeI(const Grid& g, const I3& I, const VF<T>& u, const SF<T>& f, T lr, const T* U) {  // :15-17
  T s = 0;
  for (int i = 0; i < g.D; ++i) {
    const T a = u(I, i) - U[i], b = u(sh(I, i, +1), i) - U[i];
    s += (a * a + b * b) * linInterpProp(f(I), lr);
  }
  return 0.25 * (double)s;
}
template <class T> inline T rhogh(const Grid& g, const I3& I, const T* grav, const SF<T>& f, T lr, const T* statWL) {  // :25
  T s = 0;
  for (int i = 0; i < g.D; ++i) s += grav[i] * ((T(I.i[i]) - T(1.5)) - statWL[i]);
  return -linInterpProp(f(I), lr) * s;
}
template <class T> inline double rhouI(const Grid& g, int i, const I3& I, const VF<T>& u, const SF<T>& f, T lr, const T* U) {  // :49-51
  return 0.5 * (double)(u(I, i) + u(sh(I, i, +1), i) - 2 * U[i]) * (double)linInterpProp(f(I), lr);
}
// EnsI(I,ω), :34-41.  3-D: ω is a vector field; 2-D: a scalar field
template <class T> inline double EnsI(const Grid& g, const I3& I, const T* om) {
  auto at = [&](const I3& J, int c) { return om[lin(g, J) + (int64_t)c * g.S]; };
  if (g.D == 3) {
    T s = 0;
    for (int i = 0; i < 3; ++i) {
      const int ix = (i + 1) % 3, iy = (i + 2) % 3;  // shiftDir(i,3,1), shiftDir(i,3,2)
      const T a = at(I, i), b = at(sh(I, ix, +1), i), c = at(sh(I, iy, +1), i), d = at(sh(sh(I, ix, +1), iy, +1), i);
      s += a * a + b * b + c * c + d * d;
    }
    return 0.5 * 0.25 * (double)s;
  }
  const T a = at(I, 0), b = at(sh(I, 0, +1), 0), c = at(sh(I, 1, +1), 0), d = at(sh(sh(I, 0, +1), 1, +1), 0);
  return 0.5 * 0.25 * (double)(a * a + b * b + c * c + d * d);
}

}  // namespace orc
