// oracle/oracle_flow.hpp -- TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).
// CMOM momentum transport + SynDRoM + CFL, src/flow.jl:5-57,157-241,262-298.  "Parity unpinned":
// no reference test pins these; the reference source is the only authority.
#pragma once
#include "oracle_fields.hpp"

namespace orc {

// ϕq (SynDRoM), flow.jl:37-57.  Ii = (I, i); j sweep direction; fOld = face-centred old f (dρ).
template <class T>
inline T phiq(const Grid& g, int j, int i, const I3& I, const VF<T>& fOld, const VF<T>& rhouf, T uu, T cc, T dd, T dt, T lr, int lam) {
  T Psi = (rhouf(I, j) + rhouf(sh(I, i, -1), j)) / 2;  // ϕ(i,CI(I,j),ρuf)
  I3 ICell = (Psi > 0) ? sh(I, j, -1) : I;
  T vI = cc;
  T vd = limiter(lam, uu, cc, dd);
  T va = 2 * vI - vd;
  T mOut = std::abs(Psi) * dt;
  T mOld = linInterpProp(fOld(ICell, i), lr);  // getρ(IiCell,fOld,λρ)
  if (mOut > mOld) return Psi * vI;
  T l2 = std::abs(mOut) / mOld;
  T l1 = 1 - l2;
  T vb = l2 * va + l1 * vd;
  return Psi * (vb + vd) / 2;
}
// ϕu / ϕuP / ϕuL / ϕuR, flow.jl:20-35 (f = uStar component i)
template <class T>
inline T phiu(const Grid& g, int j, int i, const I3& I, T Psi, const SF<T>& f, const VF<T>& rhouf, const VF<T>& fOld, T dt, T lr, int lam) {
  return (Psi > 0) ? phiq(g, j, i, I, fOld, rhouf, f(sh(I, j, -2)), f(sh(I, j, -1)), f(I), dt, lr, lam)
                   : phiq(g, j, i, I, fOld, rhouf, f(sh(I, j, +1)), f(I), f(sh(I, j, -1)), dt, lr, lam);
}
template <class T>
inline T phiuP(const Grid& g, int j, int i, const I3& Ip, const I3& I, T Psi, const SF<T>& f, const VF<T>& rhouf, const VF<T>& fOld, T dt,
               T lr, int lam) {
  return (Psi > 0) ? phiq(g, j, i, I, fOld, rhouf, f(Ip), f(sh(I, j, -1)), f(I), dt, lr, lam)
                   : phiq(g, j, i, I, fOld, rhouf, f(sh(I, j, +1)), f(I), f(sh(I, j, -1)), dt, lr, lam);
}
template <class T>
inline T phiuL(const Grid& g, int j, int i, const I3& I, T Psi, const SF<T>& f, const VF<T>& rhouf, const VF<T>& fOld, T dt, T lr, int lam) {
  return (Psi > 0) ? phiq(g, j, i, I, fOld, rhouf, 2 * f(sh(I, j, -1)) - f(I), f(sh(I, j, -1)), f(I), dt, lr, lam)
                   : phiq(g, j, i, I, fOld, rhouf, f(sh(I, j, +1)), f(I), f(sh(I, j, -1)), dt, lr, lam);
}
template <class T>
inline T phiuR(const Grid& g, int j, int i, const I3& I, T Psi, const SF<T>& f, const VF<T>& rhouf, const VF<T>& fOld, T dt, T lr, int lam) {
  return (Psi < 0) ? phiq(g, j, i, I, fOld, rhouf, 2 * f(I) - f(sh(I, j, -1)), f(I), f(sh(I, j, -1)), dt, lr, lam)
                   : phiq(g, j, i, I, fOld, rhouf, f(sh(I, j, -2)), f(sh(I, j, -1)), f(I), dt, lr, lam);
}

// advectρuu1D!, flow.jl:212-241
template <class T>
void advectrhouu1D(const Grid& g, const VF<T>& rhou, const VF<T>& r, const SF<T>& Phi, const VF<T>& rhouf, const VF<T>& uStar,
                   const VF<T>& uOld, const VF<T>& fOld, const SF<T>& dil, const VF<T>& u, const VF<T>& u0, const int8_t* cbar, T lr, int lam,
                   int d, T dt, unsigned perdir) {
  const int D = g.D;
  std::memset(r.p, 0, sizeof(T) * g.S * D);
  const int j = d;
  loop(r_inside(g), [&](I3 I) { dil(I) = linInterpProp(T(cbar[lin(g, I)]), lr) * (d_vec(d, I, u) + d_vec(d, I, u0)) / 2; });
  BCf(g, dil, perdir);
  const int64_t Nj = g.n[j];
  for (int i = 0; i < D; ++i) {
    const bool tagper = isper(perdir, j);
    SF<T> us = uStar.comp(i);
    auto Psi_at = [&](const I3& I) { return (rhouf(I, j) + rhouf(sh(I, i, -1), j)) / 2; };
    // lower boundary (:235 / :239-240)
    if (!tagper)
      loop(r_slice(g, 2, j, 2), [&](I3 I) { r(I, i) += phiuL(g, j, i, I, Psi_at(I), us, rhouf, fOld, dt, lr, lam); });
    else
      loop(r_slice(g, 2, j, 2), [&](I3 I) {
        Phi(I) = phiuP(g, j, i, CIj(j, I, Nj - 2), I, Psi_at(I), us, rhouf, fOld, dt, lr, lam);
        r(I, i) += Phi(I);
      });
    // inner cells (:223-225)
    loop(r_inside_u(g, j), [&](I3 I) {
      Phi(I) = phiu(g, j, i, I, Psi_at(I), us, rhouf, fOld, dt, lr, lam);
      r(I, i) += Phi(I);
    });
    loop(r_inside_u(g, j), [&](I3 I) { r(sh(I, j, -1), i) -= Phi(I); });
    // upper boundary (:236 / :241)
    if (!tagper)
      loop(r_slice(g, Nj, j, 2), [&](I3 I) { r(sh(I, j, -1), i) += -phiuR(g, j, i, I, Psi_at(I), us, rhouf, fOld, dt, lr, lam); });
    else
      loop(r_slice(g, Nj, j, 2), [&](I3 I) { r(sh(I, j, -1), i) -= Phi(CIj(j, I, 2)); });
    // dilation source (:229)
    loop(r_inside(g), [&](I3 I) { r(I, i) += uOld(I, i) * phi_face(i, I, dil); });
  }
  // axpy!(δt, r, ρu)  (:231)
  const int64_t tot = g.S * D;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int64_t l = 0; l < tot; ++l) rhou.p[l] = rhou.p[l] + dt * r.p[l];
}

// advectVOFρuu!, flow.jl:165-210.  Array arguments keep the reference's aliasing freedom: the caller
// may pass uStar≡n̂ and dilaU≡α exactly as advectfq! does (flow.jl:157-160).
template <class T>
int advectVOFrhouu(const Grid& g, T* f_, T* ff_, T* al_, T* nh_, T* u_, T* u0_, T Dt, int8_t* cbar, T* rhou_, T* r_, T* Phi_, T* rhouf_,
                   T* uStar_, T* uOld_, T* dilaU_, T* drho_, T lr, int lam, int ns, const T* uBC, unsigned perdir, bool exitBC,
                   const int* dirO, FillReport* rep, int only_op = -1) {
  // only_op >= 0: run just the directional sweep number only_op of the call (c̄ is computed with sweep 0) -- lets a slab-decomposed
  // test driver exchange ghost planes between the sweeps exactly like the multi-GPU path does
  const int D = g.D;
  SF<T> f{f_, &g}, ff{ff_, &g}, al{al_, &g}, Phi{Phi_, &g}, dil{dilaU_, &g};
  VF<T> nh{nh_, &g}, u{u_, &g}, u0{u0_, &g}, rhou{rhou_, &g}, r{r_, &g}, rhouf{rhouf_, &g}, uStar{uStar_, &g}, uOld{uOld_, &g},
      drho{drho_, &g};
  const T tol = 10 * std::numeric_limits<T>::epsilon();
  if (only_op <= 0) compute_cbar(g, f, cbar);
  int status = 0;
  if (rep) { rep->status = 0; rep->dir = -1; }
  for (int iOp = 0; iOp < D; ++iOp) {
    if (only_op >= 0 && iOp != only_op) continue;
    const int d = dirO[iOp] - 1;
    const T dt = T(1) * Dt;
    rhou2u(g, r, rhou, f, lr);                 // :197
    BC_vec(g, r, uBC, exitBC, perdir);
    std::memcpy(Phi_, f_, sizeof(T) * g.S);   // :199
    std::memset(rhouf_, 0, sizeof(T) * g.S * D);  // :201
    int st = advectVOF1d(g, f, ff, al, nh, u, u0, dt, cbar, rhouf, lr, ns, d, perdir, tol, 10 * tol, rep);  // :202
    if (st < 0) { if (rep) rep->status = st; return st; }
    status |= st;
    if (rep) rep->status = status;
    f2face(g, drho, Phi, perdir);                      // :205
    std::memcpy(uStar_, r_, sizeof(T) * g.S * D);      // :206
    {                                                  // :207  rmul!(ρuf, inv(δt)); BC!
      const T idt = T(1) / dt;
      const int64_t tot = g.S * D;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
      for (int64_t l = 0; l < tot; ++l) rhouf_[l] *= idt;
    }
    BC_vec(g, rhouf, uBC, exitBC, perdir);
    advectrhouu1D(g, rhou, r, Phi, rhouf, uStar, uOld, drho, dil, u, u0, cbar, lr, lam, d, dt, perdir);  // :208
  }
  return status;
}

// MPCFL, flow.jl:262-298 (gravity limit enters as |g| evaluated by the caller; <=0 disables)
template <class T>
T MPCFL(const Grid& g, const VF<T>& u, const SF<T>& sigma, T nu, T mu, T lam_mu, T lam_rho, T eta, T gnorm, T dt_max, T safety) {
  std::memset(sigma.p, 0, sizeof(T) * g.S);
  auto maximum = [&]() {
    T m = -std::numeric_limits<T>::infinity();
    for (int64_t l = 0; l < g.S; ++l) m = std::max(m, sigma.p[l]);
    return m;
  };
  loop(r_inside(g), [&](I3 I) {  // WaterLily flux_out
    T s = 0;
    for (int i = 0; i < g.D; ++i) s += std::max(T(0), u(sh(I, i, +1), i)) + std::max(T(0), -u(I, i));
    sigma(I) = s;
  });
  T dtAdv = 1 / (maximum() + 5 * nu);
  loop(r_inside(g), [&](I3 I) {  // maxTotalFlux :283-289
    T s = 0;
    for (int i = 0; i < g.D; ++i) s += std::max(std::abs(u(I, i)), std::abs(u(sh(I, i, +1), i)));
    sigma(I) = s;
  });
  T dtVOF = 1 / (2 * maximum());
  T dtGrav = (gnorm > 0) ? 1 / (2 * gnorm) : dt_max;
  T dtVisc = (mu > 0) ? 3 / (14 * mu * std::max(T(1), lam_mu / lam_rho)) : dt_max;
  T dtSurf = (eta > 0) ? std::sqrt((1 + lam_rho) / (T(8 * M_PI) * eta)) : dt_max;
  return safety * std::min(std::min(std::min(dtVOF, dtAdv), std::min(dtGrav, dtVisc)), dtSurf);
}

}  // namespace orc
