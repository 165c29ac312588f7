// oracle/oracle_fields.hpp -- TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).
// Field passes of the reference, one function per reference function, one loop per `@loop`.
#pragma once
#include "oracle_core.hpp"

namespace orc {

enum NormalScheme { NS_WH = 0, NS_WY, NS_COLUMN, NS_PCD, NS_SLIC, NS_MYC, NS_YOUNGS, NS_CD, NS_XYLIC };

// ----------------------------------------------------------------------------------------------
// Interface normals, src/normalEstimation.jl.  All write n̂[I,:] in place like the reference.
// Float32 runs promote to Float64 wherever the reference multiplies/compares with a Float64 literal.
// ----------------------------------------------------------------------------------------------
template <class T> inline void normal_PCD(int D, const SF<T>& f, const VF<T>& nh, const I3& I) {  // :161-165
  for (int d = 0; d < D; ++d) nh(I, d) = f(sh(I, d, -1)) - f(sh(I, d, +1));
}
template <class T> inline int argabsmax_at(int D, const VF<T>& nh, const I3& I) {
  T v[3] = {0, 0, 0};
  for (int d = 0; d < D; ++d) v[d] = nh(I, d);
  return myArgAbsMax(D, v);
}
template <class T> inline void normal_Column(int D, const SF<T>& f, const VF<T>& nh, const I3& I) {  // :75-97
  normal_PCD(D, f, nh, I);
  int dom = argabsmax_at(D, nh, I);
  for (int d = 0; d < D; ++d) {
    if (d == dom) {
      T s = sgn(nh(I, d));
      nh(I, d) = (s == 0) ? T(1) : s;
      continue;
    }
    T hl = get3CellHeight(f, sh(I, d, -1), dom);
    T hr = get3CellHeight(f, sh(I, d, +1), dom);
    nh(I, d) = (hl - hr) / 2;
  }
}
template <class T> inline void normal_WY(int D, const SF<T>& f, const VF<T>& nh, const I3& I) {  // :35-68
  normal_PCD(D, f, nh, I);
  int dom = argabsmax_at(D, nh, I);
  for (int d = 0; d < D; ++d) {
    if (d == dom) {
      T s = sgn(nh(I, d));
      nh(I, d) = (s == 0) ? T(1) : s;
      continue;
    }
    T hl = get3CellHeight(f, sh(I, d, -1), dom);
    T hc = get3CellHeight(f, I, dom);
    T hr = get3CellHeight(f, sh(I, d, +1), dom);
    T n = (hl - hr) / 2;
    // `abs(n)>0.5`, `n*(hc-1.5) >= 0.0` : Float64 literals promote
    if (std::abs((double)n) > 0.5) n = ((double)n * ((double)hc - 1.5) >= 0.0) ? hc - hr : hl - hc;
    nh(I, d) = n;
  }
}
template <class T> inline void normal_WH(int D, const SF<T>& f, const VF<T>& nh, const I3& I) {  // :105-154
  normal_Column(D, f, nh, I);
  // majorDir (util.jl:72-75): signed dominant direction
  int adom = argabsmax_at(D, nh, I);
  int sdom = std::signbit(nh(I, adom)) ? -1 : +1;
  T an = std::abs(nh(I, adom));
  an = (an == 0) ? T(1) : an;
  for (int d = 0; d < D; ++d) nh(I, d) /= an;
  for (int d = 0; d < D; ++d) {
    if (d == adom) continue;
    T slope = std::abs(nh(I, d));
    int scur = std::signbit(nh(I, d)) ? -1 : +1;  // curDir = copysign(d, n̂[I,d])
    T hl = get3CellHeight(f, sh(I, d, -scur), adom);
    T hc = get3CellHeight(f, I, adom);
    T hr = get3CellHeight(f, sh(I, d, +scur), adom);
    T sumh = hl + hc + hr;
    double s45 = 4.5 * (double)slope;  // `4.5slope` is Float64 for either T
    if (s45 <= (double)sumh && (double)sumh <= 9 - s45) continue;
    // min(1/2slope, 4.5slope): first operand in T (Int literals), then promoted by min
    double thr = std::min((double)(T(1) / (2 * slope)), s45);
    if ((double)sumh < s45) {
      T wb = get3CellHeight(f, sh(I, adom, -sdom), d);
      if ((double)wb > thr) nh(I, d) = (T)std::copysign(((double)hl - 0.5) / ((double)wb - 0.5), (double)scur);
    }
    if ((double)sumh > 9 - s45) {
      T wt = get3CellHeight(f, sh(I, adom, +sdom), d);
      if ((double)wt < 3 - thr) nh(I, d) = (T)std::copysign((2.5 - (double)hr) / (2.5 - (double)wt), (double)scur);
    }
  }
}
template <class T> inline void normal_SLIC(int D, const SF<T>& f, const VF<T>& nh, const I3& I) {  // :172-176
  normal_PCD(D, f, nh, I);
  int d = argabsmax_at(D, nh, I);
  for (int i = 0; i < D; ++i) nh(I, i) = (i == d) ? sgn(nh(I, i)) : T(0);
}
template <class T> inline T YoungSum(int D, const SF<T>& f, const I3& I, int d) {  // :248-257
  // II ∈ I-δxy:I, III ∈ II:II+δxy, first dimension fastest, δxy = ones except in d
  int c[2], nc = 0;
  for (int k = 0; k < D; ++k)
    if (k != d) c[nc++] = k;
  T a = 0;
  if (nc == 1) {
    for (int b1 = -1; b1 <= 0; ++b1)
      for (int a1 = 0; a1 <= 1; ++a1) a += f(sh(I, c[0], b1 + a1));
  } else {
    for (int b2 = -1; b2 <= 0; ++b2)
      for (int b1 = -1; b1 <= 0; ++b1)
        for (int a2 = 0; a2 <= 1; ++a2)
          for (int a1 = 0; a1 <= 1; ++a1) a += f(sh(sh(I, c[0], b1 + a1), c[1], b2 + a2));
  }
  return a;
}
template <class T> inline void normal_Y(int D, const SF<T>& f, const VF<T>& nh, const I3& I) {  // :232-247
  T a = 0;
  for (int d = 0; d < D; ++d) {
    nh(I, d) = (T)(((double)(YoungSum(D, f, sh(I, d, -1), d) - YoungSum(D, f, sh(I, d, +1), d))) * 0.5);
    a += std::abs(nh(I, d));
  }
  if (a == 0) {
    for (int d = 0; d < D; ++d) nh(I, d) = (T)(1.0 / D);
  } else {
    for (int d = 0; d < D; ++d) nh(I, d) /= a;
  }
}
template <class T> inline void normal_CCi(int D, const SF<T>& f, const VF<T>& nh, const I3& I, int dc, T* out) {  // :210-224
  T s = 0;
  for (int d = 0; d < D; ++d) {
    if (d == dc) {
      T sg = sgn(nh(I, d));
      out[d] = (sg == 0) ? T(1) : sg;
    } else {
      T hu = get3CellHeight(f, sh(I, d, +1), dc);
      T hd = get3CellHeight(f, sh(I, d, -1), dc);
      out[d] = -(hu - hd) / 2;
    }
  }
  for (int d = 0; d < D; ++d) s += std::abs(out[d]);
  for (int d = 0; d < D; ++d) out[d] /= s;
}
template <class T> inline void normal_MYC(int D, const SF<T>& f, const VF<T>& nh, const I3& I) {  // :185-202
  normal_Y(D, f, nh, I);
  T maxN = 0;
  for (int i = 0; i < D; ++i) maxN = (std::abs(nh(I, i)) > maxN) ? std::abs(nh(I, i)) : maxN;
  T curm0 = 0;
  int CCiz = 0;
  T cur[3];
  for (int iz = 0; iz < D; ++iz) {
    normal_CCi(D, f, nh, I, iz, cur);
    if (std::abs(cur[iz]) > curm0) CCiz = iz;
    curm0 = std::abs(cur[iz]);  // unconditional, as in the reference (:195)
  }
  normal_CCi(D, f, nh, I, CCiz, cur);
  if (std::abs(cur[CCiz]) < maxN)
    for (int i = 0; i < D; ++i) nh(I, i) = cur[i];
}
template <class T> inline T crossSummation(int D, const SF<T>& f, const I3& I, int d) {  // :271-277 (γ=1)
  T a = f(I);
  for (int k = 0; k < D; ++k) a += (k != d) ? T(1) * (f(sh(I, k, -1)) + f(sh(I, k, +1))) : T(0);
  return a;
}
template <class T> inline void normal_CD(int D, const SF<T>& f, const VF<T>& nh, const I3& I) {  // :266-270
  for (int d = 0; d < D; ++d)
    nh(I, d) = (T)(((double)(crossSummation(D, f, sh(I, d, -1), d) - crossSummation(D, f, sh(I, d, +1), d))) * 0.5);
}
template <class T> inline void normal_XYLIC(int D, const VF<T>& nh, const I3& I) {  // :279-283 (d=2)
  for (int i = 0; i < D; ++i) nh(I, i) = (i == 1) ? T(1) : T(0);
}
template <class T> inline void normalScheme(int ns, int D, const SF<T>& f, const VF<T>& nh, const I3& I) {
  switch (ns) {
    case NS_WH: normal_WH(D, f, nh, I); break;
    case NS_WY: normal_WY(D, f, nh, I); break;
    case NS_COLUMN: normal_Column(D, f, nh, I); break;
    case NS_PCD: normal_PCD(D, f, nh, I); break;
    case NS_SLIC: normal_SLIC(D, f, nh, I); break;
    case NS_MYC: normal_MYC(D, f, nh, I); break;
    case NS_YOUNGS: normal_Y(D, f, nh, I); break;
    case NS_CD: normal_CD(D, f, nh, I); break;
    case NS_XYLIC: normal_XYLIC(D, nh, I); break;
  }
}

// ----------------------------------------------------------------------------------------------
// Boundary conditions, src/VOFutil.jl:44-119 and WaterLily BC!
// ----------------------------------------------------------------------------------------------
inline bool isper(unsigned perdir, int j) { return (perdir >> j) & 1u; }

template <class T> void BCf(const Grid& g, const SF<T>& f, unsigned perdir) {  // VOFutil.jl:64-75
  for (int j = 0; j < g.D; ++j) {
    const int64_t Nj = g.n[j];
    if (isper(perdir, j)) {
      loop(r_slice(g, 1, j), [&](I3 I) { f(I) = f(CIj(j, I, Nj - 1)); });
      loop(r_slice(g, Nj, j), [&](I3 I) { f(I) = f(CIj(j, I, 2)); });
    } else {
      loop(r_slice(g, 1, j), [&](I3 I) { f(I) = f(sh(I, j, +1)); });
      loop(r_slice(g, Nj, j), [&](I3 I) { f(I) = f(sh(I, j, -1)); });
    }
  }
}
template <class T> void BCv1D(const Grid& g, const SF<T>& f, int d, unsigned perdir) {  // VOFutil.jl:106-119 (= BCf!(d,f) :76-89)
  for (int j = 0; j < g.D; ++j) {
    const int64_t Nj = g.n[j];
    if (isper(perdir, j)) {
      loop(r_slice(g, 1, j), [&](I3 I) { f(I) = f(CIj(j, I, Nj - 1)); });
      loop(r_slice(g, Nj, j), [&](I3 I) { f(I) = f(CIj(j, I, 2)); });
    } else if (j == d) {
      loop(r_slice(g, 1, j), [&](I3 I) { f(I) = f(sh(I, j, +2)); });
    } else {
      loop(r_slice(g, 1, j), [&](I3 I) { f(I) = f(sh(I, j, +1)); });
      loop(r_slice(g, Nj, j), [&](I3 I) { f(I) = f(sh(I, j, -1)); });
    }
  }
}
template <class T> void BCv(const Grid& g, const VF<T>& f, unsigned perdir) {  // VOFutil.jl:91-104
  for (int d = 0; d < g.D; ++d)
    for (int j = 0; j < g.D; ++j) {
      const int64_t Nj = g.n[j];
      if (isper(perdir, j)) {
        loop(r_slice(g, 1, j), [&](I3 I) { f(I, d) = f(CIj(j, I, Nj - 1), d); });
        loop(r_slice(g, Nj, j), [&](I3 I) { f(I, d) = f(CIj(j, I, 2), d); });
      } else if (j == d) {
        loop(r_slice(g, 1, j), [&](I3 I) { f(I, d) = f(sh(I, j, +2), d); });
      } else {
        loop(r_slice(g, 1, j), [&](I3 I) { f(I, d) = f(sh(I, j, +1), d); });
        loop(r_slice(g, Nj, j), [&](I3 I) { f(I, d) = f(sh(I, j, -1), d); });
      }
    }
}
template <class T> void BCVOF(const Grid& g, const SF<T>& f, const SF<T>& al, const VF<T>& nh, unsigned perdir) {  // VOFutil.jl:44-62
  for (int j = 0; j < g.D; ++j) {
    const int64_t Nj = g.n[j];
    if (isper(perdir, j)) {
      auto fan = [&](I3 I, int64_t ii) {
        I3 J = CIj(j, I, ii);
        f(I) = f(J);
        for (int i = 0; i < g.D; ++i) nh(I, i) = nh(J, i);
        al(I) = al(J);
      };
      loop(r_slice(g, 1, j), [&](I3 I) { fan(I, Nj - 1); });
      loop(r_slice(g, Nj, j), [&](I3 I) { fan(I, 2); });
    } else {
      loop(r_slice(g, 1, j), [&](I3 I) { f(I) = f(sh(I, j, +1)); });
      loop(r_slice(g, Nj, j), [&](I3 I) { f(I) = f(sh(I, j, -1)); });
    }
  }
}
// WaterLily.BC!(a,A,saveexit,perdir) for a constant-tuple A (unpinned by any reference test)
template <class T> void BC_vec(const Grid& g, const VF<T>& a, const T* A, bool saveexit, unsigned perdir) {
  for (int i = 0; i < g.D; ++i)
    for (int j = 0; j < g.D; ++j) {
      const int64_t Nj = g.n[j];
      if (isper(perdir, j)) {
        loop(r_slice(g, 1, j), [&](I3 I) { a(I, i) = a(CIj(j, I, Nj - 1), i); });
        loop(r_slice(g, Nj, j), [&](I3 I) { a(I, i) = a(CIj(j, I, 2), i); });
      } else if (i == j) {
        for (int64_t s = 1; s <= 2; ++s) loop(r_slice(g, s, j), [&](I3 I) { a(I, i) = A[i]; });
        if (!saveexit || i > 0) loop(r_slice(g, Nj, j), [&](I3 I) { a(I, i) = A[i]; });
      } else {
        loop(r_slice(g, 1, j), [&](I3 I) { a(I, i) = a(sh(I, j, +1), i); });
        loop(r_slice(g, Nj, j), [&](I3 I) { a(I, i) = a(sh(I, j, -1), i); });
      }
    }
}

// ----------------------------------------------------------------------------------------------
// Field utilities, src/VOFutil.jl
// ----------------------------------------------------------------------------------------------
template <class T> void cleanWisp(const Grid& g, const SF<T>& f, T tol) {  // VOFutil.jl:127-136
  loop(r_inside(g), [&](I3 I) {
    T v = f(I);
    f(I) = (v < tol) ? T(0) : ((v > 1 - tol) ? T(1) : v);
  });
}
template <class T> void rhou2u(const Grid& g, const VF<T>& u, const VF<T>& ru, const SF<T>& f, T lr) {  // VOFutil.jl:198-201
  loop(r_inside(g), [&](I3 I) {
    for (int d = 0; d < g.D; ++d) u(I, d) = ru(I, d) / linInterpProp(phi_face(d, I, f), lr);
  });
}
template <class T> void u2rhou(const Grid& g, const VF<T>& ru, const VF<T>& u, const SF<T>& f, T lr) {  // VOFutil.jl:208-211
  loop(r_inside(g), [&](I3 I) {
    for (int d = 0; d < g.D; ++d) ru(I, d) = u(I, d) * linInterpProp(phi_face(d, I, f), lr);
  });
}
template <class T> void f2face(const Grid& g, const VF<T>& fF, const SF<T>& fC, unsigned perdir) {  // VOFutil.jl:229-234
  for (int d = 0; d < g.D; ++d) loop(r_inside(g), [&](I3 I) { fF(I, d) = phi_face(d, I, fC); });
  BCv(g, fF, perdir);
}

// applyVOF!(f,α,n̂,InterfaceSDF) (VOFutil.jl:9-37) with the SDF pre-sampled by the caller at the
// cell centre (sc) and at centre ± Δx e_i (sp[i], sm[i]); Δx = 0.01.
template <class T>
void applyVOF_samples(const Grid& g, const SF<T>& f, const SF<T>& al, const VF<T>& nh, const T* sc, const T* sp, const T* sm) {
  loop(r_inside(g), [&](I3 I) {
    const int64_t l = lin(g, I);
    T sumN = 0, sumN2 = 0, n[3] = {0, 0, 0};
    for (int i = 0; i < g.D; ++i) {
      T dd = sp[l + i * g.S] - sm[l + i * g.S];
      nh(I, i) = dd;
      n[i] = dd;
      sumN += dd;
      sumN2 += dd * dd;
    }
    al(I) = sumN / 2 - std::sqrt(sumN2) * sc[l];
    f(I) = getVolumeFractionD(g.D, n, al(I));
  });
  cleanWisp(g, f, 10 * std::numeric_limits<T>::epsilon());
}

// ----------------------------------------------------------------------------------------------
// Interface reconstruction + VOF face flux + 1-D VOF sweep:  src/normalEstimation.jl:10-28,
// src/advection.jl:80-137
// ----------------------------------------------------------------------------------------------
template <class T>
void reconstructInterface(const Grid& g, const SF<T>& f, const SF<T>& al, const VF<T>& nh, int ns, unsigned perdir) {
  loop(r_inside(g), [&](I3 I) {
    if (fullorempty(f(I))) {
      for (int i = 0; i < g.D; ++i) nh(I, i) = 0;
      al(I) = 0;
      return;
    }
    normalScheme(ns, g.D, f, nh, I);
    T n[3] = {0, 0, 0};
    for (int i = 0; i < g.D; ++i) n[i] = nh(I, i);
    al(I) = getInterceptD(g.D, n, f(I));
  });
  BCVOF(g, f, al, nh, perdir);
}

template <class T>
inline void getVOFFlux_face(const Grid& g, const SF<T>& ff, const SF<T>& f, const SF<T>& al, const VF<T>& nh, T dl, int d,
                            const I3& IF, const VF<T>& rhouf, T lr) {  // advection.jl:113-137
  if (dl == 0) return;
  I3 IC = (dl > 0) ? sh(IF, d, -1) : IF;
  T sumAbs = 0;
  for (int ii = 0; ii < g.D; ++ii) sumAbs += std::abs(nh(IC, ii));
  if (sumAbs == 0 || fullorempty(f(IC))) {
    ff(IF) = f(IC) * dl;
    rhouf(IF, d) += dl * lr + (1 - lr) * ff(IF);
    return;
  }
  T a = (dl > 0) ? al(IC) - nh(IC, d) * (1 - dl) : al(IC);
  T n[3] = {0, 0, 0};
  for (int ii = 0; ii < g.D; ++ii) n[ii] = nh(IC, ii) * ((ii == d) ? std::abs(dl) : T(1));
  ff(IF) = getVolumeFractionD(g.D, n, a) * dl;
  rhouf(IF, d) += dl * lr + (1 - lr) * ff(IF);
}
template <class T>
void getVOFFlux(const Grid& g, const SF<T>& ff, const SF<T>& f, const SF<T>& al, const VF<T>& nh, const VF<T>& u, const VF<T>& u0,
                T dt, int d, const VF<T>& rhouf, T lr) {  // advection.jl:108-112
  std::memset(ff.p, 0, sizeof(T) * g.S);
  loop(r_inside_uWB(g, d), [&](I3 IF) { getVOFFlux_face(g, ff, f, al, nh, dt / 2 * (u(IF, d) + u0(IF, d)), d, IF, rhouf, lr); });
}

struct FillReport {
  double maxf, minf;
  int64_t argmax[3], argmin[3];
  int dir;
  int status;  // bit0 overfill, bit1 underfill, -1 NaN, -5 "divergence ... is exploding!"
  double div_u0, div_u;  // |∇·u⁰|, |∇·u| at the reported cell (advection.jl:151,170)
};
// advection.jl:145-189: the diagnostic prints are dropped, the two throws are kept as negative statuses
template <class T> int reportFillError(const Grid& g, const SF<T>& f, const VF<T>& u, const VF<T>& u0, int d, T tol, FillReport* rep) {
  // findmax/findmin over ALL of f (ghosts included), first occurrence wins
  T mx = -std::numeric_limits<T>::infinity(), mn = std::numeric_limits<T>::infinity();
  int64_t imx = 0, imn = 0;
  bool nan = false;
#ifdef _OPENMP
#pragma omp parallel
#endif
  {
    T lmx = -std::numeric_limits<T>::infinity(), lmn = std::numeric_limits<T>::infinity();
    int64_t limx = 0, limn = 0;
    bool lnan = false;
#ifdef _OPENMP
#pragma omp for schedule(static) nowait
#endif
    for (int64_t l = 0; l < g.S; ++l) {
      T v = f.p[l];
      if (v != v) lnan = true;
      if (v > lmx) { lmx = v; limx = l; }
      if (v < lmn) { lmn = v; limn = l; }
    }
#ifdef _OPENMP
#pragma omp critical
#endif
    {
      if (lnan) nan = true;
      if (lmx > mx || (lmx == mx && limx < imx)) { mx = lmx; imx = limx; }
      if (lmn < mn || (lmn == mn && limn < imn)) { mn = lmn; imn = limn; }
    }
  }
  int st = 0;
  if (nan) st = -1;
  else {
    if (mx - 1 > tol) st |= 1;
    if (mn < -tol) st |= 2;
  }
  if (rep && (st != 0 || rep->status == 0)) {  // keep the latest offending sweep, else the latest sweep
    rep->maxf = mx; rep->minf = mn;
    rep->argmax[0] = imx % g.n[0] + 1; rep->argmax[1] = (imx / g.n[0]) % g.n[1] + 1; rep->argmax[2] = imx / (g.n[0] * g.n[1]) + 1;
    rep->argmin[0] = imn % g.n[0] + 1; rep->argmin[1] = (imn / g.n[0]) % g.n[1] + 1; rep->argmin[2] = imn / (g.n[0] * g.n[1]) + 1;
    rep->dir = d + 1;
    rep->div_u0 = rep->div_u = 0;
  }
  if (st > 0) {
    // ((du⁰+du > 10) || isnan(du⁰+du)) && error("divergence, …, is exploding!")   (:160 for the max cell, :180 for the min cell)
    auto check = [&](int64_t l) -> bool {
      I3 I{{l % g.n[0] + 1, (l / g.n[0]) % g.n[1] + 1, l / (g.n[0] * g.n[1]) + 1}};
      bool edge = false;
      for (int a = 0; a < g.D; ++a) edge = edge || I.i[a] >= g.n[a];  // div needs I+δ inside the array
      if (edge) return false;
      T d0 = 0, d1 = 0;
      for (int a = 0; a < g.D; ++a) { d0 += d_vec(a, I, u0); d1 += d_vec(a, I, u); }
      const double a0 = std::fabs((double)d0), a1 = std::fabs((double)d1);
      if (rep) { rep->div_u0 = a0; rep->div_u = a1; }
      return (a0 + a1 > 10) || (a0 + a1 != a0 + a1);
    };
    if (((st & 1) && check(imx)) || ((st & 2) && check(imn))) {
      if (rep) rep->status = -5;
      return -5;
    }
  }
  return st;
}

template <class T>
int advectVOF1d(const Grid& g, const SF<T>& f, const SF<T>& ff, const SF<T>& al, const VF<T>& nh, const VF<T>& u, const VF<T>& u0, T dt,
                const int8_t* cbar, const VF<T>& rhouf, T lr, int ns, int d, unsigned perdir, T tol, T filltol, FillReport* rep) {
  // advection.jl:80-89 (and the loop body of advectVOF! :65-72)
  reconstructInterface(g, f, al, nh, ns, perdir);
  getVOFFlux(g, ff, f, al, nh, u, u0, dt, d, rhouf, lr);
  loop(r_inside(g), [&](I3 I) {
    f(I) = f(I) + ((ff(I) - ff(sh(I, d, +1))) + (T(cbar[lin(g, I)]) * (d_vec(d, I, u) + d_vec(d, I, u0))) * dt / 2);
  });
  int st = reportFillError(g, f, u, u0, d, filltol, rep);
  cleanWisp(g, f, tol);
  BCf(g, f, perdir);
  return st;
}

template <class T> void compute_cbar(const Grid& g, const SF<T>& f, int8_t* cbar) {  // advection.jl:40, flow.jl:172
  loop(r_all(g), [&](I3 I) { cbar[lin(g, I)] = (f(I) < T(0.5)) ? 0 : 1; });
}

// advectVOF!(f,fᶠ,α,n̂,u,u⁰,Δt,c̄,ρuf,λρ,normalScheme; perdir,dirO)   advection.jl:34-78
template <class T>
int advectVOF(const Grid& g, T* f_, T* ff_, T* al_, T* nh_, T* u_, T* u0_, T Dt, int8_t* cbar, T* rhouf_, T lr, int ns, unsigned perdir,
              const int* dirO, FillReport* rep) {
  SF<T> f{f_, &g}, ff{ff_, &g}, al{al_, &g};
  VF<T> nh{nh_, &g}, u{u_, &g}, u0{u0_, &g}, rhouf{rhouf_, &g};
  const T tol = 10 * std::numeric_limits<T>::epsilon();
  std::memset(rhouf_, 0, sizeof(T) * g.S * g.D);
  compute_cbar(g, f, cbar);
  int status = 0;
  if (rep) { rep->status = 0; rep->dir = -1; }
  for (int iOp = 0; iOp < g.D; ++iOp) {  // Lie-Trotter, coefficients 1 (:57-58)
    int d = dirO[iOp] - 1;
    T dt = T(1) * Dt;
    int st = advectVOF1d(g, f, ff, al, nh, u, u0, dt, cbar, rhouf, lr, ns, d, perdir, tol, tol, rep);
    if (st < 0) { if (rep) rep->status = st; return st; }
    status |= st;
    if (rep) rep->status = status;
  }
  return status;
}

}  // namespace orc
