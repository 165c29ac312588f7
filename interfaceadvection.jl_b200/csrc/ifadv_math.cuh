// ifadv_math.cuh -- device-side scalar math of the VOF + CMOM path (sm_100a).
// PLIC forward/inverse problems (ref: src/PLIC.jl), interface-normal schemes (ref: src/normalEstimation.jl),
// flux limiters and the SynDRoM flux (ref: src/flow.jl:5-57).  Everything is evaluated in the reference's
// expression order; the translation unit is compiled with -fmad=false so no FMA contraction changes a
// rounding or flips one of the exact comparisons the algorithm branches on.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ifadv {

#define IFADV_DI __device__ __forceinline__

// ---- type-generic intrinsics -------------------------------------------------------------------------
IFADV_DI float t_abs(float x) { return fabsf(x); }
IFADV_DI double t_abs(double x) { return fabs(x); }
IFADV_DI float t_sqrt(float x) { return sqrtf(x); }
IFADV_DI double t_sqrt(double x) { return sqrt(x); }
IFADV_DI float t_cbrt(float x) { return cbrtf(x); }
IFADV_DI double t_cbrt(double x) { return cbrt(x); }
IFADV_DI float t_acos(float x) { return acosf(x); }
IFADV_DI double t_acos(double x) { return acos(x); }
IFADV_DI float t_sin(float x) { return sinf(x); }
IFADV_DI double t_sin(double x) { return sin(x); }
IFADV_DI float t_cos(float x) { return cosf(x); }
IFADV_DI double t_cos(double x) { return cos(x); }
IFADV_DI float t_min(float a, float b) { return fminf(a, b); }
IFADV_DI double t_min(double a, double b) { return fmin(a, b); }
IFADV_DI float t_max(float a, float b) { return fmaxf(a, b); }
IFADV_DI double t_max(double a, double b) { return fmax(a, b); }
template <class T> IFADV_DI T t_sign(T x) { return T((x > T(0)) - (x < T(0))); }
template <class T> IFADV_DI bool t_signbit(T x) { return signbit(x); }
template <class T> IFADV_DI void t_swap(T& a, T& b) { T t = a; a = b; b = t; }

// Division hooks.  Float64 is always IEEE-exact (parity <= 1e-12 against the oracle).  Float32 may be built with
// -DIFADV_FAST_F32 (csrc/Makefile default): reciprocal-multiply division (<= 2 ulp) and FMA contraction, well inside
// the 1e-5 Float32 tolerance of the north star, which removes ~10 instructions and one slow-path branch per division.
IFADV_DI double t_div(double a, double b) { return a / b; }
IFADV_DI double t_div6(double a) { return a / 6.0; }
#ifdef IFADV_FAST_F32
IFADV_DI float t_div(float a, float b) { return __fdividef(a, b); }
IFADV_DI float t_div6(float a) { return a * (1.0f / 6.0f); }
#else
IFADV_DI float t_div(float a, float b) { return a / b; }
IFADV_DI float t_div6(float a) { return a / 6.0f; }
#endif

template <class T> IFADV_DI bool fullorempty(T fc) { return fc == T(0) || fc == T(1); }      // VOFutil.jl:151
template <class T> IFADV_DI T lin_interp(T f, T lam, T oml) { return lam + oml * f; }        // VOFutil.jl:166, oml = 1-λ

// ρ at a face: linInterpProp(ϕ(d,I,f),λρ) = λ + (1-λ)*((a+b)/2)   (VOFutil.jl:176).  Fast Float32 folds the halving into the FMA.
template <class T> IFADV_DI T rho_face(T a, T b, T lam, T oml) {
#ifdef IFADV_FAST_F32
  if (sizeof(T) == 4) return fmaf((float)oml * 0.5f, (float)(a + b), (float)lam);
#endif
  return lam + oml * ((a + b) / T(2));
}

// ---- PLIC: src/PLIC.jl --------------------------------------------------------------------------------
template <class T> IFADV_DI void sort2(T& a, T& b) { if (!(a < b)) t_swap(a, b); }           // :178
template <class T> IFADV_DI void sort3(T& a, T& b, T& c) {                                   // :186-191
  if (a > c) t_swap(a, c);
  if (a > b) t_swap(a, b);
  if (b > c) t_swap(b, c);
}
template <class T> __device__ T proot(T c0, T c1, T c2, T c3) {                              // :165-176
  T a0 = c0 / c3, a1 = c1 / c3, a2 = c2 / c3;
  T p0 = a1 / T(3) - (a2 * a2) / T(9);
  T q0 = (a1 * a2 - T(3) * a0) / T(6) - (a2 * a2 * a2) / T(27);
  T a = q0 / t_sqrt(-(p0 * p0 * p0));
  T t = t_acos((a * a <= T(1)) ? a : T(0)) / T(3);
  return t_sqrt(-p0) * (t_sqrt(T(3)) * t_sin(t) - t_cos(t)) - a2 / T(3);
}
template <class T> IFADV_DI T alpha2f_2(T m1, T m2, T a) {                                   // :92
  return a < m1 ? (a * a) / ((T(2) * m1) * m2) : (a - m1 / T(2)) / m2;
}
template <class T> __device__ T alpha2f_3(T m1, T m2, T m3, T a) {                           // :93-107
  T m12 = m1 + m2;
  if (a < m1) return (a * a * a) / (((T(6) * m1) * m2) * m3);
  else if (a < m2) return (a * (a - m1)) / ((T(2) * m2) * m3) + (((m2 == T(0)) ? T(1) : m1 / m2) * m1) / (T(6) * m3);
  else if (a < t_min(m3, m12))
    return ((a * a) * (T(3) * m12 - a) + (m1 * m1) * (m1 - T(3) * a) + (m2 * m2) * (m2 - T(3) * a)) / (((T(6) * m1) * m2) * m3);
  else if (m3 < m12)
    return ((a * a) * (T(3) - T(2) * a) + (m1 * m1) * (m1 - T(3) * a) + (m2 * m2) * (m2 - T(3) * a) + (m3 * m3) * (m3 - T(3) * a)) /
           (((T(6) * m1) * m2) * m3);
  else return (T(2) * a - m12) / (T(2) * m3);
}
template <class T> IFADV_DI T f2alpha_2(T m1, T m2, T v) {                                   // :117
  return v < m1 / (T(2) * m2) ? t_sqrt(((T(2) * m1) * m2) * v) : m2 * v + m1 / T(2);
}
template <class T> __device__ T f2alpha_3(T m1, T m2, T m3, T v) {                           // :118-149
  T m12 = m1 + m2;
  T p = ((T(6) * m1) * m2) * m3;
  T v1 = (((m2 == T(0)) ? T(1) : m1 / m2) * m1) / (T(6) * m3);
  T v2 = v1 + (m2 - m1) / (T(2) * m3);
  T v3 = (m3 < m12) ? ((m3 * m3) * (T(3) * m12 - m3) + (m1 * m1) * (m1 - T(3) * m3) + (m2 * m2) * (m2 - T(3) * m3)) / p
                    : m12 / (T(2) * m3);
  if (v < v1) return t_cbrt(p * v);
  else if (v < v2) return (m1 + t_sqrt(m1 * m1 + ((T(8) * m2) * m3) * (v - v1))) / T(2);
  else if (v < v3) {
    T c0 = (m1 * m1 * m1 + m2 * m2 * m2) - p * v;
    T c1 = T(-3) * (m1 * m1 + m2 * m2);
    return proot(c0, c1, T(3) * m12, T(-1));
  } else if (m3 < m12) {
    T c0 = ((m1 * m1 * m1 + m2 * m2 * m2) + m3 * m3 * m3) - p * v;
    T c1 = T(-3) * ((m1 * m1 + m2 * m2) + m3 * m3);
    return proot(c0, c1, T(3), T(-2));
  } else return m3 * v + m12 / T(2);
}
// getIntercept(n̂, g): :20-39
template <class T, int D> __device__ T get_intercept(const T n[3], T g) {
  if (D == 2) {
    T t = t_abs(n[0]) + t_abs(n[1]);
    T a;
    if (g != T(0.5)) {
      T m1 = t_abs(n[0]) / t, m2 = t_abs(n[1]) / t;
      sort2(m1, m2);
      a = f2alpha_2(m1, m2, (g < T(0.5)) ? g : T(1) - g);
    } else a = T(0.5);
    return ((g < T(0.5)) ? a : T(1) - a) * t + t_min(n[0], T(0)) + t_min(n[1], T(0));
  } else {
    T t = t_abs(n[0]) + t_abs(n[1]) + t_abs(n[2]);
    T a;
    if (g != T(0.5)) {
      T m1 = t_abs(n[0]) / t, m2 = t_abs(n[1]) / t, m3 = t_abs(n[2]) / t;
      sort3(m1, m2, m3);
      a = f2alpha_3(m1, m2, m3, (g < T(0.5)) ? g : T(1) - g);
    } else a = T(0.5);
    return ((g < T(0.5)) ? a : T(1) - a) * t + t_min(n[0], T(0)) + t_min(n[1], T(0)) + t_min(n[2], T(0));
  }
}
// getVolumeFraction(n̂, b): :59-82
template <class T, int D> __device__ T get_volume_fraction(const T n[3], T b) {
  if (D == 2) {
    T t = t_abs(n[0]) + t_abs(n[1]);
    T a = (b - t_min(n[0], T(0)) - t_min(n[1], T(0))) / t;
    if (a <= T(0) || a == T(0.5) || a >= T(1)) return t_min(t_max(a, T(0)), T(1));
    T m1 = t_abs(n[0]) / t, m2 = t_abs(n[1]) / t;
    sort2(m1, m2);
    T r = alpha2f_2(m1, m2, (a < T(0.5)) ? a : T(1) - a);
    return (a < T(0.5)) ? r : T(1) - r;
  } else {
    T t = t_abs(n[0]) + t_abs(n[1]) + t_abs(n[2]);
    T a = (b - t_min(n[0], T(0)) - t_min(n[1], T(0)) - t_min(n[2], T(0))) / t;
    if (a <= T(0) || a == T(0.5) || a >= T(1)) return t_min(t_max(a, T(0)), T(1));
    T m1 = t_abs(n[0]) / t, m2 = t_abs(n[1]) / t, m3 = t_abs(n[2]) / t;
    sort3(m1, m2, m3);
    T r = alpha2f_3(m1, m2, m3, (a < T(0.5)) ? a : T(1) - a);
    return (a < T(0.5)) ? r : T(1) - r;
  }
}

// ---- flux limiters λ(u,c,d): src/flow.jl:5-15 (+ WaterLily quick / vanLeer / cds) ----------------------
template <class T> IFADV_DI T median3(T a, T b, T c) {  // WaterLily median: the middle value
  return t_max(t_min(a, b), t_min(t_max(a, b), c));
}
template <class T> IFADV_DI T sweby(T u, T c, T d, T gam) {
  T s = t_sign(d - u);
  if (c <= t_min(u, d) || c >= t_max(u, d)) return c;
  T m1 = t_min((s * gam) * (c - u), s * (d - c));
  T m2 = t_min(s * (c - u), (s * gam) * (d - c));
  return c + (s * t_max(T(0), t_max(m1, m2))) / T(2);
}
template <class T> __device__ __noinline__ T limiter_other(int lam, T u, T c, T d) {
  switch (lam) {
    case 0: return c;
    case 1: return median3((T(3) * c - u) / T(2), c, (c + d) / T(2));
    case 2: return median3(t_div6(T(7) * c + d - T(2) * u), c, median3(T(2) * c - u, c, d));
    case 3: {
      T al = c - u, be = d - c;
      T w = (al == be && al == T(0)) ? T(0) : (al + be) / (al * al + be * be);
      return c + (t_max(al * be, T(0)) * w) / T(2);
    }
    case 4: return sweby(u, c, d, T(1.5));
    case 5: return sweby(u, c, d, T(2));
    case 6: {
      T s = t_sign(d - u);
      if (c <= t_min(u, d) || c >= t_max(u, d)) return c;
      return c + s * t_min(s * (c - u), (s * (d - c)) / T(2));
    }
    case 7: {
      T s = t_sign(d - u);
      if (c <= t_min(u, d) || c >= t_max(u, d)) return c;
      return c + s * t_min(s * (c - u), s * (d - c));
    }
    case 8: return median3((T(5) * c + T(2) * d - u) / T(6), c, median3(T(10) * c - T(9) * u, c, d));
    case 9: return (c <= t_min(u, d) || c >= t_max(u, d)) ? c : c + (d - c) * (c - u) / (d - u);
    case 10: return (c + d) / T(2);
  }
  return c;
}
template <class T> IFADV_DI T limiter(int lam, T u, T c, T d) {
  if (lam == 2) return median3(t_div6(T(7) * c + d - T(2) * u), c, median3(T(2) * c - u, c, d));  // Koren, the default: inline
  return limiter_other(lam, u, c, d);  // the other limiters: one out-of-line copy per translation unit
}
// ϕq, the SynDRoM flux (src/flow.jl:37-57) for mass flux Psi, stencil (uu,cc,dd) and donor density mOld
template <class T> IFADV_DI T syndrom_blend(T vd, T Psi, T cc, T mOld, T dt) {
  T mOut = t_abs(Psi) * dt;
#ifdef IFADV_FAST_F32
  if (sizeof(T) == 4) {
    // (vb+vd)/2 = l2*cc + (1-l2)*vd in exact arithmetic; branch-free, within Float32 round-off of the reference order
    const T l2 = (mOut > mOld) ? T(1) : t_div(mOut, mOld);
    return Psi * (vd + l2 * (cc - vd));
  }
#endif
  T va = T(2) * cc - vd;
  if (mOut > mOld) return Psi * cc;
  T l2 = t_div(t_abs(mOut), mOld);
  T l1 = T(1) - l2;
  T vb = l2 * va + l1 * vd;
  return (Psi * (vb + vd)) / T(2);
}
template <class T> IFADV_DI T syndrom_flux(int lam, T Psi, T uu, T cc, T dd, T mOld, T dt) {
  return syndrom_blend(limiter(lam, uu, cc, dd), Psi, cc, mOld, dt);
}
// compile-time Koren (the package default, flow.jl): no limiter dispatch in the kernel body
template <bool KOREN, class T> IFADV_DI T syndrom_flux_t(int lam, T Psi, T uu, T cc, T dd, T mOld, T dt) {
  if (KOREN) return syndrom_blend(median3(t_div6(T(7) * cc + dd - T(2) * uu), cc, median3(T(2) * cc - uu, cc, dd)), Psi, cc, mOld, dt);
  return syndrom_blend(limiter_other(lam, uu, cc, dd), Psi, cc, mOld, dt);
}

// ---- interface normals: src/normalEstimation.jl -------------------------------------------------------
// BX: box accessor, B(dx,dy,dz) = f[I + (dx,dy,dz)] with the ghost rule of BCf! already applied.
#define IFADV_E(dir, s) (s) * ((dir) == 0), (s) * ((dir) == 1), (s) * ((dir) == 2)
template <class T, class BX> IFADV_DI T h3(const BX& B, int ox, int oy, int oz, int dir) {  // get3CellHeight, VOFutil.jl:158
  const int ex = (dir == 0), ey = (dir == 1), ez = (dir == 2);
  return B(ox, oy, oz) + B(ox - ex, oy - ey, oz - ez) + B(ox + ex, oy + ey, oz + ez);
}
template <class T, int D> IFADV_DI int arg_abs_max(const T n[3]) {  // util.jl:19-31
  T mx = T(0);
  int im = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T cur = n[i] * n[i];
    if (cur > mx) { mx = cur; im = i; }
  }
  return im;
}
template <class T> IFADV_DI T pick(const T n[3], int d) { return d == 0 ? n[0] : (d == 1 ? n[1] : n[2]); }

template <class T, int D, class BX> __device__ void normal_pcd(const BX& B, T n[3]) {  // :161-165
#pragma unroll
  for (int d = 0; d < D; ++d) n[d] = B(IFADV_E(d, -1)) - B(IFADV_E(d, +1));
}
template <class T, int D, class BX> __device__ void normal_column(const BX& B, T n[3]) {  // :75-97
  normal_pcd<T, D>(B, n);
  const int dom = arg_abs_max<T, D>(n);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (d == dom) {
      T s = t_sign(n[d]);
      n[d] = (s == T(0)) ? T(1) : s;
    } else {
      T hl = h3<T>(B, IFADV_E(d, -1), dom);
      T hr = h3<T>(B, IFADV_E(d, +1), dom);
      n[d] = (hl - hr) / T(2);
    }
  }
}
template <class T, int D, class BX> __device__ void normal_wy(const BX& B, T n[3]) {  // :35-68
  normal_pcd<T, D>(B, n);
  const int dom = arg_abs_max<T, D>(n);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (d == dom) {
      T s = t_sign(n[d]);
      n[d] = (s == T(0)) ? T(1) : s;
    } else {
      T hl = h3<T>(B, IFADV_E(d, -1), dom);
      T hc = h3<T>(B, 0, 0, 0, dom);
      T hr = h3<T>(B, IFADV_E(d, +1), dom);
      T v = (hl - hr) / T(2);
      if (fabs((double)v) > 0.5) v = ((double)v * ((double)hc - 1.5) >= 0.0) ? hc - hr : hl - hc;  // Float64 literals promote
      n[d] = v;
    }
  }
}
template <class T, int D, class BX> __device__ void normal_wh(const BX& B, T n[3]) {  // :105-154
  normal_column<T, D>(B, n);
  const int adom = arg_abs_max<T, D>(n);  // majorDir, util.jl:72-75
  const T ndom = pick(n, adom);
  const int sdom = t_signbit(ndom) ? -1 : +1;
  T an = t_abs(ndom);
  an = (an == T(0)) ? T(1) : an;
#pragma unroll
  for (int d = 0; d < D; ++d) n[d] /= an;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (d == adom) continue;
    const T slope = t_abs(n[d]);
    const int scur = t_signbit(n[d]) ? -1 : +1;
    const T hl = h3<T>(B, IFADV_E(d, -scur), adom);
    const T hc = h3<T>(B, 0, 0, 0, adom);
    const T hr = h3<T>(B, IFADV_E(d, +scur), adom);
    const T sumh = hl + hc + hr;
    const double s45 = 4.5 * (double)slope;  // `4.5slope` promotes to Float64 for either T
    if (s45 <= (double)sumh && (double)sumh <= 9.0 - s45) continue;
    const double thr = fmin((double)(T(1) / (T(2) * slope)), s45);  // min(1/2slope, 4.5slope)
    if ((double)sumh < s45) {
      const T wb = h3<T>(B, IFADV_E(adom, -sdom), d);
      if ((double)wb > thr) n[d] = (T)copysign(((double)hl - 0.5) / ((double)wb - 0.5), (double)scur);
    }
    if ((double)sumh > 9.0 - s45) {
      const T wt = h3<T>(B, IFADV_E(adom, +sdom), d);
      if ((double)wt < 3.0 - thr) n[d] = (T)copysign((2.5 - (double)hr) / (2.5 - (double)wt), (double)scur);
    }
  }
}
template <class T, int D, class BX> __device__ void normal_slic(const BX& B, T n[3]) {  // :172-176
  normal_pcd<T, D>(B, n);
  const int dom = arg_abs_max<T, D>(n);
#pragma unroll
  for (int i = 0; i < D; ++i) n[i] = (i == dom) ? t_sign(n[i]) : T(0);
}
template <class T, int D, class BX> __device__ T young_sum(const BX& B, int ox, int oy, int oz, int d) {  // :248-257
  // II ∈ I-δxy:I, III ∈ II:II+δxy, first dimension fastest
  const int c0 = (d == 0) ? 1 : 0;
  const int c1 = (D == 3) ? ((d == 2) ? 1 : 2) : -1;
  T a = T(0);
  if (D == 2) {
    for (int b1 = -1; b1 <= 0; ++b1)
      for (int a1 = 0; a1 <= 1; ++a1) a += B(ox + (b1 + a1) * (c0 == 0), oy + (b1 + a1) * (c0 == 1), oz);
  } else {
    for (int b2 = -1; b2 <= 0; ++b2)
      for (int b1 = -1; b1 <= 0; ++b1)
        for (int a2 = 0; a2 <= 1; ++a2)
          for (int a1 = 0; a1 <= 1; ++a1) {
            const int s0 = b1 + a1, s1 = b2 + a2;
            a += B(ox + s0 * (c0 == 0) + s1 * (c1 == 0), oy + s0 * (c0 == 1) + s1 * (c1 == 1), oz + s0 * (c0 == 2) + s1 * (c1 == 2));
          }
  }
  return a;
}
template <class T, int D, class BX> __device__ void normal_youngs(const BX& B, T n[3]) {  // :232-247
  T a = T(0);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    n[d] = (T)(((double)(young_sum<T, D>(B, IFADV_E(d, -1), d) - young_sum<T, D>(B, IFADV_E(d, +1), d))) * 0.5);
    a += t_abs(n[d]);
  }
  if (a == T(0)) {
#pragma unroll
    for (int d = 0; d < D; ++d) n[d] = (T)(1.0 / D);
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) n[d] /= a;
  }
}
template <class T, int D, class BX> __device__ void normal_cci(const BX& B, const T n[3], int dc, T out[3]) {  // :210-224
  T s = T(0);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (d == dc) {
      T sg = t_sign(n[d]);
      out[d] = (sg == T(0)) ? T(1) : sg;
    } else {
      T hu = h3<T>(B, IFADV_E(d, +1), dc);
      T hd = h3<T>(B, IFADV_E(d, -1), dc);
      out[d] = -(hu - hd) / T(2);
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) s += t_abs(out[d]);
#pragma unroll
  for (int d = 0; d < D; ++d) out[d] /= s;
}
template <class T, int D, class BX> __device__ void normal_myc(const BX& B, T n[3]) {  // :185-202
  normal_youngs<T, D>(B, n);
  T maxN = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) maxN = (t_abs(n[i]) > maxN) ? t_abs(n[i]) : maxN;
  T curm0 = T(0);
  int cciz = 0;
  T cur[3] = {T(0), T(0), T(0)};
  for (int iz = 0; iz < D; ++iz) {
    normal_cci<T, D>(B, n, iz, cur);
    const T c = t_abs(pick(cur, iz));
    if (c > curm0) cciz = iz;
    curm0 = c;  // unconditional, as in the reference (:195)
  }
  normal_cci<T, D>(B, n, cciz, cur);
  if (t_abs(pick(cur, cciz)) < maxN) {
#pragma unroll
    for (int i = 0; i < D; ++i) n[i] = cur[i];
  }
}
template <class T, int D, class BX> __device__ T cross_sum(const BX& B, int ox, int oy, int oz, int d) {  // :271-277
  T a = B(ox, oy, oz);
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const int ex = (k == 0), ey = (k == 1), ez = (k == 2);
    a += (k != d) ? T(1) * (B(ox - ex, oy - ey, oz - ez) + B(ox + ex, oy + ey, oz + ez)) : T(0);
  }
  return a;
}
template <class T, int D, class BX> __device__ void normal_cd(const BX& B, T n[3]) {  // :266-270
#pragma unroll
  for (int d = 0; d < D; ++d)
    n[d] = (T)(((double)(cross_sum<T, D>(B, IFADV_E(d, -1), d) - cross_sum<T, D>(B, IFADV_E(d, +1), d))) * 0.5);
}
template <class T, int D, class BX> __device__ void interface_normal(int scheme, const BX& B, T n[3]) {
  n[0] = n[1] = n[2] = T(0);
  switch (scheme) {
    case 0: normal_wh<T, D>(B, n); break;
    case 1: normal_wy<T, D>(B, n); break;
    case 2: normal_column<T, D>(B, n); break;
    case 3: normal_pcd<T, D>(B, n); break;
    case 4: normal_slic<T, D>(B, n); break;
    case 5: normal_myc<T, D>(B, n); break;
    case 6: normal_youngs<T, D>(B, n); break;
    case 7: normal_cd<T, D>(B, n); break;
    case 8: n[0] = T(0); n[1] = T(1); n[2] = T(0); break;  // XYLIC :279-283
  }
}

// PLIC volume flux through a face swept by dl, from the upwind cell with volume fraction fc
// (general branch of getVOFFlux!, src/advection.jl:131-134).  d = face direction (0-based).
template <class T, int D, class BX> IFADV_DI T plic_face_flux_inl(int scheme, const BX B, T fc, int d, T dl) {
  T n[3];
  interface_normal<T, D>(scheme, B, n);
  T sumAbs = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) sumAbs += t_abs(n[i]);
  if (sumAbs == T(0)) return fc * dl;  // advection.jl:125
  const T alpha = get_intercept<T, D>(n, fc);
  const T nd = pick(n, d);
  const T a = (dl > T(0)) ? alpha - nd * (T(1) - dl) : alpha;
  T m[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) m[i] = (i < D) ? n[i] * ((i == d) ? t_abs(dl) : T(1)) : T(0);
  return get_volume_fraction<T, D>(m, a) * dl;
}
// out-of-line copy (one ABI call per use; keeps kernels whose live state is small compact)
template <class T, int D, class BX> __device__ __noinline__ T plic_face_flux(int scheme, const BX B, T fc, int d, T dl) {
  return plic_face_flux_inl<T, D, BX>(scheme, B, fc, d, dl);
}

}  // namespace ifadv
