// oracle/oracle_core.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped, never timed as product).
//
// A plain C++17 restatement of the reference algorithm of TzuYaoHuang/InterfaceAdvection.jl for the
// VOF + CMOM advection path.  It keeps the reference's un-fused pass structure (one loop per `@loop`
// site, same order, same arrays, same aliasing) so that it can serve as the parity checker for the
// fused sm_100a kernels.  Arithmetic is evaluated in the reference's expression order without
// fast-math and without FMA contraction (build with -ffp-contract=off).
//
// Parity status: PLIC scalars, WY/MYC normals, VOF face flux, BCf!/BCv!/BCVOF!, applyVOF!, f2face!,
// getρ are PINNED by the reference's own known-answer tests (test/maintests.jl, see tests/).
// WH normal, SynDRoM (ϕq), advectρuu1D!, limiters, MPCFL and WaterLily's BC!/inside_u are
// "parity unpinned": no reference test holds values for them and the Julia reference cannot be run
// in this environment, so the source text is the only authority.
//
// Each function cites the reference file:line (relative to /root/reference) it follows.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

// ----------------------------------------------------------------------------------------------
// Index machinery mirroring Julia's 1-based CartesianIndex on column-major arrays with one ghost
// layer (array extents n = N.+2).  For D==2 the third extent is 1 and the third index is always 1.
// ----------------------------------------------------------------------------------------------
struct Grid {
  int D;
  int64_t n[3];
  int64_t S;  // elements of one scalar field
};
inline Grid make_grid(int D, const int64_t* Ng) {
  Grid g;
  g.D = D;
  g.n[0] = Ng[0];
  g.n[1] = Ng[1];
  g.n[2] = (D == 3) ? Ng[2] : 1;
  g.S = g.n[0] * g.n[1] * g.n[2];
  return g;
}
struct I3 {
  int64_t i[3];
};
inline I3 sh(I3 a, int d, int64_t s) {  // I + s*δ(d+1, I)   (d is 0-based here)
  a.i[d] += s;
  return a;
}
inline I3 CIj(int j, I3 a, int64_t k) {  // WaterLily CIj: replace j-th entry
  a.i[j] = k;
  return a;
}
inline int64_t lin(const Grid& g, const I3& a) {
  return (a.i[0] - 1) + g.n[0] * ((a.i[1] - 1) + g.n[1] * (a.i[2] - 1));
}

struct Range {
  int64_t lo[3], hi[3];
};
inline Range r_all(const Grid& g) {  // CartesianIndices(f)
  Range r;
  for (int d = 0; d < 3; ++d) { r.lo[d] = 1; r.hi[d] = g.n[d]; }
  return r;
}
inline Range r_inside(const Grid& g) {  // WaterLily inside(a): 2:size-1
  Range r = r_all(g);
  for (int d = 0; d < g.D; ++d) { r.lo[d] = 2; r.hi[d] = g.n[d] - 1; }
  return r;
}
inline Range r_inside_uWB(const Grid& g, int j) {  // src/util.jl:47-49
  Range r = r_inside(g);
  r.hi[j] = g.n[j];
  return r;
}
inline Range r_inside_u(const Grid& g, int j) {  // WaterLily inside_u(dims,j): j -> 3:n-1, others 2:n
  Range r = r_all(g);
  for (int d = 0; d < g.D; ++d) {
    if (d == j) { r.lo[d] = 3; r.hi[d] = g.n[d] - 1; }
    else { r.lo[d] = 2; r.hi[d] = g.n[d]; }
  }
  return r;
}
inline Range r_slice(const Grid& g, int64_t i, int j, int64_t low = 1) {  // WaterLily slice(dims,i,j,low)
  Range r = r_all(g);
  for (int d = 0; d < g.D; ++d) {
    if (d == j) { r.lo[d] = i; r.hi[d] = i; }
    else { r.lo[d] = low; r.hi[d] = g.n[d]; }
  }
  return r;
}

inline bool inside_others(const Grid& g, const I3& I, int j) {  // I ∈ 2:n-1 in every dimension but j (slice(N,·,j,2) runs to n)
  for (int d = 0; d < g.D; ++d)
    if (d != j && I.i[d] > g.n[d] - 1) return false;
  return true;
}

// `@loop body over I ∈ R` : every index independent (that is what lets the reference run the same
// body as a GPU kernel), so the OpenMP build may parallelise each pass.
template <class F>
inline void loop(const Range& r, F&& fn) {
  const int64_t n0 = r.hi[0] - r.lo[0] + 1, n1 = r.hi[1] - r.lo[1] + 1, n2 = r.hi[2] - r.lo[2] + 1;
  if (n0 <= 0 || n1 <= 0 || n2 <= 0) return;
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static)
#endif
  for (int64_t k = r.lo[2]; k <= r.hi[2]; ++k)
    for (int64_t j = r.lo[1]; j <= r.hi[1]; ++j)
      for (int64_t i = r.lo[0]; i <= r.hi[0]; ++i) fn(I3{{i, j, k}});
}

template <class T>
struct SF {  // scalar field view
  T* p;
  const Grid* g;
  T& operator()(const I3& a) const { return p[lin(*g, a)]; }
};
template <class T>
struct VF {  // vector field view, component index slowest
  T* p;
  const Grid* g;
  T& operator()(const I3& a, int c) const { return p[lin(*g, a) + (int64_t)c * g->S]; }
  SF<T> comp(int c) const { return SF<T>{p + (int64_t)c * g->S, g}; }
};

// ----------------------------------------------------------------------------------------------
// PLIC geometry, src/PLIC.jl
// ----------------------------------------------------------------------------------------------
template <class T> inline void sort2(T& a, T& b) {  // PLIC.jl:178
  if (!(a < b)) std::swap(a, b);
}
template <class T> inline void sort3(T& a, T& b, T& c) {  // PLIC.jl:186-191
  if (a > c) std::swap(a, c);
  if (a > b) std::swap(a, b);
  if (b > c) std::swap(b, c);
}

template <class T> inline T proot(T c0, T c1, T c2, T c3) {  // PLIC.jl:165-176
  T a0 = c0 / c3, a1 = c1 / c3, a2 = c2 / c3;
  T p0 = a1 / 3 - (a2 * a2) / 9;
  T q0 = (a1 * a2 - 3 * a0) / 6 - (a2 * a2 * a2) / 27;
  T a = q0 / std::sqrt(-(p0 * p0 * p0));
  T t = std::acos((a * a <= 1) ? a : T(0)) / 3;
  return std::sqrt(-p0) * (std::sqrt(T(3)) * std::sin(t) - std::cos(t)) - a2 / 3;
}

template <class T> inline T alpha2f(T m1, T m2, T a) {  // PLIC.jl:92  (2-D forward)
  return a < m1 ? (a * a) / ((2 * m1) * m2) : (a - m1 / 2) / m2;
}
template <class T> inline T alpha2f(T m1, T m2, T m3, T a) {  // PLIC.jl:93-107 (3-D forward)
  T m12 = m1 + m2;
  if (a < m1) return (a * a * a) / (((6 * m1) * m2) * m3);
  else if (a < m2) return (a * (a - m1)) / ((2 * m2) * m3) + (((m2 == 0) ? T(1) : m1 / m2) * m1) / (6 * m3);
  else if (a < std::min(m3, m12))
    return ((a * a) * (3 * m12 - a) + (m1 * m1) * (m1 - 3 * a) + (m2 * m2) * (m2 - 3 * a)) / (((6 * m1) * m2) * m3);
  else if (m3 < m12)
    return ((a * a) * (3 - 2 * a) + (m1 * m1) * (m1 - 3 * a) + (m2 * m2) * (m2 - 3 * a) + (m3 * m3) * (m3 - 3 * a)) /
           (((6 * m1) * m2) * m3);
  else return (2 * a - m12) / (2 * m3);
}
template <class T> inline T f2alpha(T m1, T m2, T v) {  // PLIC.jl:117 (2-D inverse)
  return v < m1 / (2 * m2) ? std::sqrt(((2 * m1) * m2) * v) : m2 * v + m1 / 2;
}
template <class T> inline T f2alpha(T m1, T m2, T m3, T v) {  // PLIC.jl:118-149 (3-D inverse)
  T m12 = m1 + m2;
  T p = ((6 * m1) * m2) * m3;
  T v1 = (((m2 == 0) ? T(1) : m1 / m2) * m1) / (6 * m3);
  T v2 = v1 + (m2 - m1) / (2 * m3);
  T v3 = (m3 < m12) ? ((m3 * m3) * (3 * m12 - m3) + (m1 * m1) * (m1 - 3 * m3) + (m2 * m2) * (m2 - 3 * m3)) / p
                    : m12 / (2 * m3);
  if (v < v1) return std::cbrt(p * v);
  else if (v < v2) return (m1 + std::sqrt(m1 * m1 + ((8 * m2) * m3) * (v - v1))) / 2;
  else if (v < v3) {
    T c0 = (m1 * m1 * m1 + m2 * m2 * m2) - p * v;
    T c1 = -3 * (m1 * m1 + m2 * m2);
    T c2 = 3 * m12;
    T c3 = -T(1);
    return proot(c0, c1, c2, c3);
  } else if (m3 < m12) {
    T c0 = ((m1 * m1 * m1 + m2 * m2 * m2) + m3 * m3 * m3) - p * v;
    T c1 = -3 * ((m1 * m1 + m2 * m2) + m3 * m3);
    T c2 = T(3);
    T c3 = -T(2);
    return proot(c0, c1, c2, c3);
  } else return m3 * v + m12 / 2;
}

template <class T> inline T getIntercept(T n1, T n2, T g) {  // PLIC.jl:20-29
  T t = std::abs(n1) + std::abs(n2);
  T a;
  if (g != T(0.5)) {
    T m1 = std::abs(n1) / t, m2 = std::abs(n2) / t;
    sort2(m1, m2);
    a = f2alpha(m1, m2, (g < T(0.5)) ? g : 1 - g);
  } else a = T(0.5);
  return ((g < T(0.5)) ? a : 1 - a) * t + std::min(n1, T(0)) + std::min(n2, T(0));
}
template <class T> inline T getIntercept(T n1, T n2, T n3, T g) {  // PLIC.jl:30-39
  T t = std::abs(n1) + std::abs(n2) + std::abs(n3);
  T a;
  if (g != T(0.5)) {
    T m1 = std::abs(n1) / t, m2 = std::abs(n2) / t, m3 = std::abs(n3) / t;
    sort3(m1, m2, m3);
    a = f2alpha(m1, m2, m3, (g < T(0.5)) ? g : 1 - g);
  } else a = T(0.5);
  return ((g < T(0.5)) ? a : 1 - a) * t + std::min(n1, T(0)) + std::min(n2, T(0)) + std::min(n3, T(0));
}
template <class T> inline T getVolumeFraction(T n1, T n2, T b) {  // PLIC.jl:59-70
  T t = std::abs(n1) + std::abs(n2);
  T a = (b - std::min(n1, T(0)) - std::min(n2, T(0))) / t;
  if (a <= 0 || a == T(0.5) || a >= 1) return std::min(std::max(a, T(0)), T(1));
  T m1 = std::abs(n1) / t, m2 = std::abs(n2) / t;
  sort2(m1, m2);
  T r = alpha2f(m1, m2, (a < T(0.5)) ? a : 1 - a);
  return (a < T(0.5)) ? r : 1 - r;
}
template <class T> inline T getVolumeFraction(T n1, T n2, T n3, T b) {  // PLIC.jl:71-82
  T t = std::abs(n1) + std::abs(n2) + std::abs(n3);
  T a = (b - std::min(n1, T(0)) - std::min(n2, T(0)) - std::min(n3, T(0))) / t;
  if (a <= 0 || a == T(0.5) || a >= 1) return std::min(std::max(a, T(0)), T(1));
  T m1 = std::abs(n1) / t, m2 = std::abs(n2) / t, m3 = std::abs(n3) / t;
  sort3(m1, m2, m3);
  T r = alpha2f(m1, m2, m3, (a < T(0.5)) ? a : 1 - a);
  return (a < T(0.5)) ? r : 1 - r;
}
template <class T> inline T getInterceptD(int D, const T* n, T g) {
  return D == 2 ? getIntercept(n[0], n[1], g) : getIntercept(n[0], n[1], n[2], g);
}
template <class T> inline T getVolumeFractionD(int D, const T* n, T b) {
  return D == 2 ? getVolumeFraction(n[0], n[1], b) : getVolumeFraction(n[0], n[1], n[2], b);
}

// ----------------------------------------------------------------------------------------------
// Small helpers: src/util.jl, src/VOFutil.jl
// ----------------------------------------------------------------------------------------------
template <class T> inline bool fullorempty(T fc) { return fc == 0 || fc == 1; }  // VOFutil.jl:151
template <class T> inline T linInterpProp(T f, T lam) { return lam + (1 - lam) * f; }  // VOFutil.jl:166 (base=1)
template <class T> inline T sgn(T x) { return T((x > 0) - (x < 0)); }
template <class T> inline T get3CellHeight(const SF<T>& f, const I3& I, int dir) {  // VOFutil.jl:158
  return f(I) + f(sh(I, dir, -1)) + f(sh(I, dir, +1));
}
template <class T> inline T phi_face(int a, const I3& I, const SF<T>& f) {  // WaterLily ϕ(a,I,f)
  return (f(I) + f(sh(I, a, -1))) / 2;
}
template <class T> inline T d_vec(int a, const I3& I, const VF<T>& u) {  // WaterLily ∂(a,I,u::vector)
  return u(sh(I, a, +1), a) - u(I, a);
}
template <class T> inline int myArgAbsMax(int D, const T* v) {  // util.jl:19-31 (0-based result)
  T mx = 0;
  int im = 0;
  for (int i = 0; i < D; ++i) {
    T cur = v[i] * v[i];
    if (cur > mx) { mx = cur; im = i; }
  }
  return im;
}
template <class T> inline T median3(T a, T b, T c) {  // WaterLily median(a,b,c)
  if (a > b) {
    if (b >= c) return b;
    if (a > c) return c;
  } else {
    if (b <= c) return b;
    if (a < c) return c;
  }
  return a;
}

// ----------------------------------------------------------------------------------------------
// Flux limiters λ(u,c,d), src/flow.jl:5-15  (+ WaterLily's quick/vanLeer/cds, recalled, unverifiable)
// ----------------------------------------------------------------------------------------------
enum Limiter { L_UPWIND = 0, L_MINMOD, L_KOREN, L_VANALBADA1, L_SWEBY, L_SUPERBEE, L_TVDCEN, L_TVDDOWN, L_QUICK, L_VANLEER, L_CDS };
template <class T> inline T sweby(T u, T c, T d, T gam) {  // flow.jl:12
  T s = sgn(d - u);
  if (c <= std::min(u, d) || c >= std::max(u, d)) return c;
  T m1 = std::min((s * gam) * (c - u), s * (d - c));
  T m2 = std::min(s * (c - u), (s * gam) * (d - c));
  return c + (s * std::max(T(0), std::max(m1, m2))) / 2;
}
template <class T> inline T limiter(int lam, T u, T c, T d) {
  switch (lam) {
    case L_UPWIND: return c;                                                        // flow.jl:5
    case L_MINMOD: return median3((3 * c - u) / 2, c, (c + d) / 2);                // flow.jl:6
    case L_KOREN: return median3((7 * c + d - 2 * u) / 6, c, median3(2 * c - u, c, d));  // flow.jl:7
    case L_VANALBADA1: {                                                            // flow.jl:8-11
      T al = c - u, be = d - c;
      T w = (al == be && al == 0) ? T(0) : (al + be) / (al * al + be * be);
      return c + (std::max(al * be, T(0)) * w) / 2;
    }
    case L_SWEBY: return sweby(u, c, d, T(1.5));   // flow.jl:12
    case L_SUPERBEE: return sweby(u, c, d, T(2));  // flow.jl:13
    case L_TVDCEN: {                               // flow.jl:14
      T s = sgn(d - u);
      if (c <= std::min(u, d) || c >= std::max(u, d)) return c;
      return c + s * std::min(s * (c - u), (s * (d - c)) / 2);
    }
    case L_TVDDOWN: {  // flow.jl:15
      T s = sgn(d - u);
      if (c <= std::min(u, d) || c >= std::max(u, d)) return c;
      return c + s * std::min(s * (c - u), s * (d - c));
    }
    case L_QUICK: return median3((5 * c + 2 * d - u) / 6, c, median3(10 * c - 9 * u, c, d));  // WaterLily quick
    case L_VANLEER:                                                                           // WaterLily vanLeer
      return (c <= std::min(u, d) || c >= std::max(u, d)) ? c : c + (d - c) * (c - u) / (d - u);
    case L_CDS: return (c + d) / 2;  // WaterLily cds
  }
  return c;
}

}  // namespace orc
