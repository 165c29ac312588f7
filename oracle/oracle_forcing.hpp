// oracle/oracle_forcing.hpp -- TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).
// Explicit forcing between advection and projection: viscSurfTenρu! / visc! / getμ,
// surfTen! + height-function curvature, updateU!, updateL!.  src/flow.jl:113-153,244-259; src/VOFutil.jl:186-191;
// src/surfaceTension.jl:8-101.  One loop per `@loop`, same order, same arrays as the reference.
//
// Parity status: getμ (test/maintests.jl:67-70), getPopinetHeight and the 2-D getCurvature (:182-190) are PINNED by the
// reference's own known-answer tests (tests/test_oracle_forcing.py).  visc!, the 3-D getCurvature, surfTen!, updateU!,
// updateL! are "parity unpinned": no reference test holds values for them; the source text is the only authority.
// WaterLily's accelerate! is restated for a constant gravity vector only (g(i,x,t) = g[i]).
#pragma once
#include "oracle_fields.hpp"

namespace orc {

template <class T> inline bool containInterface(T f) { return T(0) < f && f < T(1); }  // VOFutil.jl:144

// getμ(i,j,I,fFace,λμ,μ,λρ), VOFutil.jl:186-191
template <class T> inline T getmu(int i, int j, const I3& I, const VF<T>& fF, T lmu, T mu, T lr) {
  const T f1 = fF(sh(I, j, -1), i), f2 = fF(I, i), f3 = fF(sh(I, i, -1), j), f4 = fF(I, j);
  const T s = (f1 + f2 + f3 + f4) / 4;
  const T frmin = (lr < 1) ? std::min(std::min(std::min(f1, f2), f3), f4) : std::max(std::max(std::max(f1, f2), f3), f4);
  const T w = ((double)s > 0.5) ? T(1) : lmu / lr;
  return mu * std::min(linInterpProp(s, lmu), w * linInterpProp(frmin, lr));
}
// scalar-form ∂(a,Ii,u) with Ii = (I,i) of rank D+1 (WaterLily): u[I,i] - u[I-δ(a),i]
template <class T> inline T d_sc(int a, const I3& I, int i, const VF<T>& u) { return u(I, i) - u(sh(I, a, -1), i); }
// viscF(i,j,I,u,fFace,λμ,μ,λρ), flow.jl:141
template <class T> inline T viscF(int i, int j, const I3& I, const VF<T>& u, const VF<T>& fF, T lmu, T mu, T lr) {
  return getmu(i, j, I, fF, lmu, mu, lr) * (d_sc(j, I, i, u) + d_sc(i, I, j, u));
}

// visc!(r,u,fFace,Φ,f,λμ,μ,λρ;perdir), flow.jl:120-138 with the boundary blocks :144-152
template <class T>
void visc(const Grid& g, const VF<T>& r, const VF<T>& u, const VF<T>& fF, const SF<T>& Phi, const SF<T>& f, T lmu, T mu, T lr,
          unsigned perdir) {
  f2face(g, fF, f, perdir);
  for (int i = 0; i < g.D; ++i)
    for (int j = 0; j < g.D; ++j) {
      const int64_t Nj = g.n[j];
      if (!isper(perdir, j)) {  // lowerBoundaryVisc!, Val{false}
        loop(r_slice(g, 2, j, 2), [&](I3 I) { r(I, i) += -viscF(i, j, I, u, fF, lmu, mu, lr); });
      } else {  // Val{true}
        loop(r_slice(g, 2, j, 2), [&](I3 I) {
          Phi(I) = -viscF(i, j, I, u, fF, lmu, mu, lr);
          r(I, i) += Phi(I);
        });
      }
      loop(r_inside_u(g, j), [&](I3 I) {
        Phi(I) = -viscF(i, j, I, u, fF, lmu, mu, lr);
        r(I, i) += Phi(I);
      });
      loop(r_inside_u(g, j), [&](I3 I) { r(sh(I, j, -1), i) -= Phi(I); });
      if (!isper(perdir, j)) {  // upperBoundaryVisc!
        loop(r_slice(g, Nj, j, 2), [&](I3 I) { r(sh(I, j, -1), i) += viscF(i, j, I, u, fF, lmu, mu, lr); });
      } else {
        loop(r_slice(g, Nj, j, 2), [&](I3 I) { r(sh(I, j, -1), i) -= Phi(CIj(j, I, 2)); });
      }
    }
}

// δd(i,I) with a signed 1-based direction i (util.jl:56): I + s*sign(i)*δ(|i|)
inline I3 shd(I3 a, int sdir, int64_t s) {
  const int d = (sdir < 0 ? -sdir : sdir) - 1;
  a.i[d] += (sdir < 0 ? -s : s);
  return a;
}
inline bool validCI(const Grid& g, const I3& I) {  // util.jl:81
  for (int d = 0; d < g.D; ++d)
    if (I.i[d] < 1 || I.i[d] > g.n[d]) return false;
  return true;
}
// getPopinetHeightAdaptive(I,f,i,monotonic=true), surfaceTension.jl:76-99 (i: signed 1-based direction)
template <class T> inline T getPopinetHeight(const Grid& g, const I3& I, const SF<T>& f, int i) {
  I3 Inow = I;
  T fnow = f(Inow);
  T H = fnow - T(0.5);
  bool fin = fnow < 1;
  while (!fin || containInterface(fnow)) {
    Inow = shd(Inow, i, +1);
    if (!validCI(g, Inow)) break;
    const T fi = f(Inow);
    fnow = (fi > fnow) ? T(0) : fi;
    H += fnow;
    fin = containInterface(fnow) ? true : fin;
  }
  Inow = I;
  fnow = f(Inow);
  fin = fnow > 0;
  while (!fin || containInterface(fnow)) {
    Inow = shd(Inow, i, -1);
    if (!validCI(g, Inow)) break;
    const T fi = f(Inow);
    fnow = (fi < fnow) ? T(1) : fi;
    H += fnow - 1;
    fin = containInterface(fnow) ? true : fin;
  }
  return H;
}
template <class T> inline T root1p5(T a) { return std::sqrt(a * a * a); }  // surfaceTension.jl:101
inline int isgn(int i) { return (i > 0) - (i < 0); }
// getCurvature(I,f,i), surfaceTension.jl:31-65 (i: signed 1-based major direction)
template <class T> inline T getCurvature(const Grid& g, const I3& I, const SF<T>& f, int i) {
  const int ai = i < 0 ? -i : i;
  if (g.D == 3) {
    const int ix = isgn(i) * (ai % 3 + 1), iy = (ai + 1) % 3 + 1;  // getXYdir, util.jl:65
    T H[3][3];
    for (int a = -1; a <= 1; ++a)
      for (int b = -1; b <= 1; ++b) H[a + 1][b + 1] = getPopinetHeight(g, shd(shd(I, ix, a), iy, b), f, i);
    const T filter = T(0.2);
    const T Hx = (H[2][1] - H[0][1]) / 2;
    const T Hy = (H[1][2] - H[1][0]) / 2;
    const T Hxx = ((H[2][1] + H[0][1] - 2 * H[1][1]) + (H[2][0] + H[0][0] - 2 * H[1][0]) * filter + (H[2][2] + H[0][2] - 2 * H[1][2]) * filter) /
                  (1 + 2 * filter);
    const T Hyy = ((H[1][2] + H[1][0] - 2 * H[1][1]) + (H[0][2] + H[0][0] - 2 * H[0][1]) * filter + (H[2][2] + H[2][0] - 2 * H[2][1]) * filter) /
                  (1 + 2 * filter);
    const T Hxy = (H[2][2] + H[0][0] - H[2][0] - H[0][2]) / 4;
    return (Hxx * (1 + Hy * Hy) + Hyy * (1 + Hx * Hx) - 2 * Hxy * Hx * Hy) / root1p5(1 + Hx * Hx + Hy * Hy);
  }
  const int ix = (ai == 1) ? -2 * isgn(i) : isgn(i);  // getXdir, util.jl:64
  T H[3];
  for (int a = -1; a <= 1; ++a) H[a + 1] = getPopinetHeight(g, shd(I, ix, a), f, i);
  const T Hx = (H[2] - H[0]) / 2;
  const T Hxx = H[2] + H[0] - 2 * H[1];
  return Hxx / root1p5(1 + Hx * Hx);
}
// majorDir(n̂,I), util.jl:72-75: signed 1-based direction of the largest |n̂| (first strict maximum)
template <class T> inline int majorDir(int D, const VF<T>& nh, const I3& I) {
  const int i = argabsmax_at(D, nh, I);
  return std::signbit(nh(I, i)) ? -(i + 1) : (i + 1);
}

// surfTen!(forcing,f,α,n̂,fbuffer,η;perdir), surfaceTension.jl:8-21
template <class T>
void surfTen(const Grid& g, const VF<T>& forcing, const SF<T>& f, const VF<T>& nh, const SF<T>& fb, T eta, unsigned perdir) {
  for (int d = 0; d < g.D; ++d) {
    loop(r_inside(g), [&](I3 I) { fb(I) = phi_face(d, I, f); });
    BCv1D(g, fb, d, perdir);  // BCf!(d,fbuffer;perdir), VOFutil.jl:76-89
    loop(r_inside(g), [&](I3 I) {
      if (containInterface(fb(I))) normal_WY(g.D, fb, nh, I);
    });
    loop(r_inside(g), [&](I3 I) {
      if (containInterface(fb(I))) forcing(I, d) += eta * getCurvature(g, I, fb, majorDir(g.D, nh, I)) * -(f(I) - f(sh(I, d, -1)));
    });
  }
}

// viscSurfTenρu!(r,u,Φ,f,α,n̂,fbuffer,λμ,μ,λρ,η;perdir), flow.jl:113-117.  has_mu / has_eta: μ, η !== nothing
template <class T>
void viscSurfTenRhou(const Grid& g, const VF<T>& r, const VF<T>& u, const SF<T>& Phi, const SF<T>& f, const VF<T>& nh, const SF<T>& fb, T lmu,
                     T mu, bool has_mu, T lr, T eta, bool has_eta, unsigned perdir) {
  const int64_t n = g.S * g.D;
  for (int64_t k = 0; k < n; ++k) r.p[k] = 0;
  if (has_mu) visc(g, r, u, nh, Phi, f, lmu, mu, lr, perdir);
  if (has_eta) surfTen(g, r, f, nh, fb, eta, perdir);
}

// updateU!(u,ρu,ρu⁰,forcing,dt,f,λρ,tNow,g,uBC,w), flow.jl:244-252; gravity: constant vector or null (accelerate! with g=nothing and a
// constant uBC tuple adds nothing)
template <class T>
void updateU(const Grid& g, const VF<T>& u, const VF<T>& ru, const VF<T>& ru0, const VF<T>& forcing, T dt, const SF<T>& f, T lr,
             const T* grav, T w) {
  const T a = 1 / w - 1;
  const int64_t n = g.S * g.D;
  for (int64_t k = 0; k < n; ++k) ru.p[k] = (a * ru0.p[k] + ru.p[k] + forcing.p[k] * dt) * w;
  rhou2u(g, u, ru, f, lr);
  for (int64_t k = 0; k < n; ++k) forcing.p[k] = 0;
  if (grav)
    for (int i = 0; i < g.D; ++i)
      for (int64_t k = 0; k < g.S; ++k) forcing.p[k + i * g.S] += grav[i];
  const T c = dt * w;
  for (int64_t k = 0; k < n; ++k) u.p[k] = c * forcing.p[k] + u.p[k];  // axpy!(dt*wT, forcing, u)
}

// updateL!(μ₀,f,λρ;perdir), flow.jl:254-259
template <class T> void updateL(const Grid& g, const VF<T>& mu0, const SF<T>& f, T lr, unsigned perdir) {
  for (int d = 0; d < g.D; ++d) loop(r_inside(g), [&](I3 I) { mu0(I, d) /= linInterpProp(phi_face(d, I, f), lr); });
  const T Z[3] = {0, 0, 0};
  BC_vec(g, mu0, Z, false, perdir);
}

}  // namespace orc
