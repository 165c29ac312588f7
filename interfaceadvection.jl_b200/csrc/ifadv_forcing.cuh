// ifadv_forcing.cuh -- explicit forcing between advection and projection:
//   viscSurfTenρu! = fill!(r,0) + visc! + surfTen!      src/flow.jl:113-152, src/VOFutil.jl:186-191, src/surfaceTension.jl:8-101
//   updateU!                                             src/flow.jl:244-252
//   fill!(μ₀,1) + updateL!                               src/flow.jl:73,96,254-259
// The reference runs D² + 2 gather/scatter passes for visc!, D × 4 passes for surfTen! and 5 passes for updateU!; here:
//   visc_kernel      one thread per momentum cell evaluates the viscous stress flux through its own lower and upper j-faces for every
//                    (i,j) from f and u directly -- fFace = f2face!(f) + BCv! is never materialised (closed form of the ghost rules),
//                    Φ is not used -- and writes r once;
//   fbuffer_kernel   fbuffer = ϕ(d,·,f) with the ghost rules of BCf!(d,·) in one pass (the reference's array, because the column
//                    walks of the height function are unbounded);
//   surften_kernel   interface cells of fbuffer: Weymouth-Yue normal in registers, major direction, 3^(D-1) Popinet column heights,
//                    curvature, r[I,d] += η κ (-∂_d f);
//   update_u_kernel  ρu ← (a ρu⁰ + ρu + forcing·dt)·w, u ← ρu/ρ(f̄) (+ dt·w·g), forcing ← g in one pass.
// Arithmetic follows the reference expression by expression; this translation unit is compiled with -fmad=false and IEEE division in
// BOTH precisions, so Float32 and Float64 results equal the oracle's bit for bit.
// Only inside(f) entries of r are produced (the reference also accumulates into upper ghost entries that nothing reads).
#pragma once
#include "ifadv_math.cuh"
#include "ifadv_sweep.cuh"

namespace ifadv {

template <bool B> struct BoolK { static constexpr bool value = B; };

// Index J of a staggered field of component d after f2face! + BCv! / BCf!(d,·) (VOFutil.jl:76-104): the sequential plane passes
// j = 1..D resolve from the last dimension down.  Plane N_d of a non-periodic direction d is never written by the reference
// (`stale`): the lower dimensions keep resolving on that plane.  Returns false when the entry is stale.
template <int D> IFADV_DI bool stag_map(const Geo& g, int d, int v[3]) {
  bool ok = true;
#pragma unroll
  for (int j = D - 1; j >= 0; --j) {
    const int n = g.n[j];
    const bool per = (g.per >> j) & 1u;
    if (v[j] == 1) v[j] = per ? n - 1 : ((j == d) ? 3 : 2);
    else if (v[j] == n) {
      if (per) v[j] = 2;
      else if (j == d) ok = false;
      else v[j] = n - 1;
    }
  }
  return ok;
}
IFADV_DI long long stride_of(const Geo& g, int d) { return d == 0 ? 1 : (d == 1 ? g.s1 : g.s2); }

// fFace[J,d] = ϕ(d,J,f) with the ghost rules above; stale entries come from the caller's n̂ array like in the reference
template <class T, int D, bool INTERIOR> IFADV_DI T fface(const T* __restrict__ f, const T* __restrict__ stale, const Geo& g, int x, int y, int z, int d) {
  if (INTERIOR) {
    const long long l = lin3(g, x, y, z);
    return (__ldg(f + l) + __ldg(f + l - stride_of(g, d))) / T(2);
  }
  int v[3] = {x, y, z};
  const bool ok = stag_map<D>(g, d, v);
  const long long l = lin3(g, v[0], v[1], v[2]);
  if (!ok) return __ldg(stale + (long long)d * g.S + l);
  return (__ldg(f + l) + __ldg(f + l - stride_of(g, d))) / T(2);
}

// viscF(i,j,I,u,fFace,λμ,μ,λρ) = getμ(i,j,I,…)·(∂ⱼuᵢ + ∂ᵢuⱼ), flow.jl:141, VOFutil.jl:186-191
template <class T, int D, bool INTERIOR>
IFADV_DI T visc_flux(const T* __restrict__ f, const T* __restrict__ u, const T* __restrict__ stale, const Geo& g, int i, int j, int x, int y, int z,
                     T lmu, T omlmu, T mu, T lr, T omlr, T wlight) {
  const int ex = (j == 0), ey = (j == 1), ez = (j == 2), ix = (i == 0), iy = (i == 1), iz = (i == 2);
  const T f1 = fface<T, D, INTERIOR>(f, stale, g, x - ex, y - ey, z - ez, i);
  const T f2 = fface<T, D, INTERIOR>(f, stale, g, x, y, z, i);
  const T f3 = (i == j) ? f1 : fface<T, D, INTERIOR>(f, stale, g, x - ix, y - iy, z - iz, j);
  const T f4 = (i == j) ? f2 : fface<T, D, INTERIOR>(f, stale, g, x, y, z, j);
  const T s = (f1 + f2 + f3 + f4) / T(4);
  const T fm = (lr < T(1)) ? t_min(t_min(t_min(f1, f2), f3), f4) : t_max(t_max(t_max(f1, f2), f3), f4);
  const T w = (s > T(0.5)) ? T(1) : wlight;
  const T muI = mu * t_min(lin_interp(s, lmu, omlmu), w * lin_interp(fm, lr, omlr));
  const long long l = lin3(g, x, y, z);
  const T* ui = u + (long long)i * g.S;
  const T* uj = u + (long long)j * g.S;
  const T du = (__ldg(ui + l) - __ldg(ui + l - stride_of(g, j))) + (__ldg(uj + l) - __ldg(uj + l - stride_of(g, i)));
  return muI * du;
}

// fill!(r,0) + visc! on inside(f), flow.jl:114,120-152.  has_mu == 0: r ← 0 only.
template <class T, int D> __global__ void __launch_bounds__(128) visc_kernel(T* __restrict__ r, const T* __restrict__ u, const T* __restrict__ f,
                                                                            const T* __restrict__ stale, const Geo g, T lmu, T mu, T lr, int has_mu) {
  const int x = 2 + blockIdx.x * blockDim.x + threadIdx.x, y = 2 + blockIdx.y, z = (D == 3) ? 2 + blockIdx.z : 1;
  if (x > g.n[0] - 1) return;
  const long long l = lin3(g, x, y, z);
  T out[3] = {T(0), T(0), T(0)};
  if (has_mu) {
    const T omlmu = T(1) - lmu, omlr = T(1) - lr, wlight = lmu / lr;
    // every index the stencil touches is inside(f): no ghost rules (3 <= v <= n-2 in every dimension)
    const bool interior = x >= 3 && x <= g.n[0] - 2 && y >= 3 && y <= g.n[1] - 2 && (D == 2 || (z >= 3 && z <= g.n[2] - 2));
    const int v[3] = {x, y, z};
    auto body = [&](auto ic) {
      constexpr bool IN = decltype(ic)::value;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        T acc = T(0);
#pragma unroll
        for (int j = 0; j < D; ++j) {
          // lower j-face: Φ[I] = -viscF; r[I,i] += Φ[I]     (lowerBoundaryVisc! / inner loop)
          const T plo = -visc_flux<T, D, IN>(f, u, stale, g, i, j, x, y, z, lmu, omlmu, mu, lr, omlr, wlight);
          acc = acc + plo;
          // upper j-face: r[I,i] -= Φ[I+δj]; at the upper boundary += viscF(I+δj) (Neumann) or -= Φ[CIj(j,I+δj,2)] (periodic)
          int w0 = x + (j == 0), w1 = y + (j == 1), w2 = z + (j == 2);
          if (!IN && ((g.per >> j) & 1u) && v[j] + 1 == g.n[j]) {
            if (j == 0) w0 = 2; else if (j == 1) w1 = 2; else w2 = 2;
          }
          const T phi = -visc_flux<T, D, IN>(f, u, stale, g, i, j, w0, w1, w2, lmu, omlmu, mu, lr, omlr, wlight);
          acc = acc - phi;
        }
        out[i] = acc;
      }
    };
    if (interior) body(BoolK<true>{}); else body(BoolK<false>{});
  }
#pragma unroll
  for (int i = 0; i < D; ++i) r[(long long)i * g.S + l] = out[i];
}

// fbuffer = ϕ(d,·,f) on inside(f) + BCf!(d,fbuffer;perdir), surfaceTension.jl:10-11: all entries in one pass (stale ones are left alone)
template <class T, int D> __global__ void fbuffer_kernel(T* fb, const T* __restrict__ f, const Geo g, int d) {
  const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x, y = 1 + blockIdx.y, z = (D == 3) ? 1 + blockIdx.z : 1;
  if (x > g.n[0]) return;
  int v[3] = {x, y, z};
  const bool ok = stag_map<D>(g, d, v);
  const long long l = lin3(g, v[0], v[1], v[2]), lj = lin3(g, x, y, z);
  if (!ok) {  // plane N_d of a non-periodic d: the lower-dimension passes copy WITHIN the plane, its inside entries keep the caller's values
    if (l != lj) fb[lj] = fb[l];
    return;
  }
  fb[lj] = (__ldg(f + l) + __ldg(f + l - stride_of(g, d))) / T(2);
}

template <class T> IFADV_DI bool contain_interface(T f) { return T(0) < f && f < T(1); }  // VOFutil.jl:144

template <class T> struct GBox {  // 3^D box on a field whose ghost entries are materialised
  const T* p;
  long long l, s1, s2;
  IFADV_DI T operator()(int dx, int dy, int dz) const { return __ldg(p + l + dx + dy * s1 + dz * s2); }
};

// getPopinetHeightAdaptive(I,f,i,monotonic=true), surfaceTension.jl:76-99; sd: signed 1-based direction
template <class T, int D> IFADV_DI T popinet_height(const T* __restrict__ fb, const Geo& g, int x, int y, int z, int sd) {
  const int a = (sd < 0 ? -sd : sd) - 1, sg = (sd < 0) ? -1 : 1;
  const long long st = stride_of(g, a) * sg;
  const int n = g.n[a];
  const int c0 = (a == 0) ? x : (a == 1 ? y : z);
  const long long l0 = lin3(g, x, y, z);
  const T f0 = __ldg(fb + l0);
  T H = f0 - T(0.5);
  {
    T fnow = f0;
    bool fin = fnow < T(1);
    int c = c0;
    long long l = l0;
    while (!fin || contain_interface(fnow)) {
      c += sg; l += st;
      if (c < 1 || c > n) break;
      const T fi = __ldg(fb + l);
      fnow = (fi > fnow) ? T(0) : fi;
      H += fnow;
      fin = contain_interface(fnow) ? true : fin;
    }
  }
  {
    T fnow = f0;
    bool fin = fnow > T(0);
    int c = c0;
    long long l = l0;
    while (!fin || contain_interface(fnow)) {
      c -= sg; l -= st;
      if (c < 1 || c > n) break;
      const T fi = __ldg(fb + l);
      fnow = (fi < fnow) ? T(1) : fi;
      H += fnow - T(1);
      fin = contain_interface(fnow) ? true : fin;
    }
  }
  return H;
}
template <class T> IFADV_DI T root1p5(T a) { return t_sqrt(a * a * a); }  // surfaceTension.jl:101

// getCurvature(I,f,i), surfaceTension.jl:31-65
template <class T, int D> IFADV_DI T curvature(const T* __restrict__ fb, const Geo& g, int x, int y, int z, int sd) {
  const int ai = sd < 0 ? -sd : sd, sg = sd < 0 ? -1 : 1;
  if (D == 3) {
    const int ix = sg * (ai % 3 + 1), iy = (ai + 1) % 3 + 1;  // getXYdir, util.jl:65
    const int ax = (ix < 0 ? -ix : ix) - 1, sx = ix < 0 ? -1 : 1, ay = iy - 1;
    T H[3][3];
#pragma unroll
    for (int a = -1; a <= 1; ++a)
#pragma unroll
      for (int b = -1; b <= 1; ++b) {
        int v[3] = {x, y, z};
        v[ax] += a * sx;
        v[ay] += b;
        H[a + 1][b + 1] = popinet_height<T, D>(fb, g, v[0], v[1], v[2], sd);
      }
    const T filter = T(0.2);
    const T Hx = (H[2][1] - H[0][1]) / T(2);
    const T Hy = (H[1][2] - H[1][0]) / T(2);
    const T Hxx = ((H[2][1] + H[0][1] - T(2) * H[1][1]) + (H[2][0] + H[0][0] - T(2) * H[1][0]) * filter +
                   (H[2][2] + H[0][2] - T(2) * H[1][2]) * filter) / (T(1) + T(2) * filter);
    const T Hyy = ((H[1][2] + H[1][0] - T(2) * H[1][1]) + (H[0][2] + H[0][0] - T(2) * H[0][1]) * filter +
                   (H[2][2] + H[2][0] - T(2) * H[2][1]) * filter) / (T(1) + T(2) * filter);
    const T Hxy = (H[2][2] + H[0][0] - H[2][0] - H[0][2]) / T(4);
    return (Hxx * (T(1) + Hy * Hy) + Hyy * (T(1) + Hx * Hx) - T(2) * Hxy * Hx * Hy) / root1p5(T(1) + Hx * Hx + Hy * Hy);
  }
  const int ix = (ai == 1) ? -2 * sg : sg;  // getXdir, util.jl:64
  const int ax = (ix < 0 ? -ix : ix) - 1, sx = ix < 0 ? -1 : 1;
  T H[3];
#pragma unroll
  for (int a = -1; a <= 1; ++a) {
    int v[3] = {x, y, z};
    v[ax] += a * sx;
    H[a + 1] = popinet_height<T, D>(fb, g, v[0], v[1], v[2], sd);
  }
  const T Hx = (H[2] - H[0]) / T(2);
  const T Hxx = H[2] + H[0] - T(2) * H[1];
  return Hxx / root1p5(T(1) + Hx * Hx);
}

// calNormal! + applySurfTen! on inside(fbuffer), surfaceTension.jl:12-21
template <class T, int D> __global__ void __launch_bounds__(128) surften_kernel(T* __restrict__ r, const T* __restrict__ fb, const T* __restrict__ f,
                                                                               const Geo g, int d, T eta) {
  const int x = 2 + blockIdx.x * blockDim.x + threadIdx.x, y = 2 + blockIdx.y, z = (D == 3) ? 2 + blockIdx.z : 1;
  if (x > g.n[0] - 1) return;
  const long long l = lin3(g, x, y, z);
  if (!contain_interface(__ldg(fb + l))) return;
  GBox<T> B{fb, l, g.s1, (D == 3) ? g.s2 : 0};
  T n[3] = {T(0), T(0), T(0)};
  normal_wy<T, D>(B, n);                    // getInterfaceNormal_WY!(fbuffer,n̂,I)
  const int im = arg_abs_max<T, D>(n);      // majorDir(n̂,I), util.jl:72-75
  const int sd = t_signbit(pick(n, im)) ? -(im + 1) : (im + 1);
  const T kappa = curvature<T, D>(fb, g, x, y, z, sd);
  const long long ld = (long long)d * g.S + l;
  r[ld] = r[ld] + eta * kappa * -(__ldg(f + l) - __ldg(f + l - stride_of(g, d)));
}

// updateU!, flow.jl:244-252, over ALL entries (CartesianIndices(ρu)); ρu2u! on inside(f); accelerate! for a constant gravity vector
template <class T, int D> __global__ void update_u_kernel(T* __restrict__ u, T* __restrict__ ru, const T* __restrict__ ru0, T* __restrict__ fo,
                                                         const T* __restrict__ f, const Geo g, T dt, T lr, T w, T G0, T G1, T G2, int has_g) {
  const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x, y = 1 + blockIdx.y, z = (D == 3) ? 1 + blockIdx.z : 1;
  if (x > g.n[0]) return;
  const long long l = lin3(g, x, y, z);
  const bool in = x >= 2 && x <= g.n[0] - 1 && y >= 2 && y <= g.n[1] - 1 && (D == 2 || (z >= 2 && z <= g.n[2] - 1));
  const T a = T(1) / w - T(1), omlr = T(1) - lr, c = dt * w;
  const T fc = in ? __ldg(f + l) : T(0);
  const T G[3] = {G0, G1, G2};
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const long long ld = (long long)d * g.S + l;
    const T q = (a * __ldg(ru0 + ld) + ru[ld] + fo[ld] * dt) * w;
    ru[ld] = q;
    T un = u[ld];
    if (in) un = q / lin_interp((fc + __ldg(f + l - stride_of(g, d))) / T(2), lr, omlr);  // ρu2u!, VOFutil.jl:198-201
    if (has_g) un = c * G[d] + un;                                                         // axpy!(dt*w, forcing, u)
    if (in || has_g) u[ld] = un;
    fo[ld] = has_g ? G[d] : T(0);
  }
}

// μ₀[I,d] /= getρ(d,I,f,λρ) on inside(f) (fill_one: after fill!(μ₀,1)), flow.jl:254-257; BC! follows as a second launch
template <class T, int D> __global__ void update_l_kernel(T* __restrict__ mu0, const T* __restrict__ f, const Geo g, T lr, int fill_one) {
  const int x = 2 + blockIdx.x * blockDim.x + threadIdx.x, y = 2 + blockIdx.y, z = (D == 3) ? 2 + blockIdx.z : 1;
  if (x > g.n[0] - 1) return;
  const long long l = lin3(g, x, y, z);
  const T fc = __ldg(f + l), omlr = T(1) - lr;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const long long ld = (long long)d * g.S + l;
    const T rho = lin_interp((fc + __ldg(f + l - stride_of(g, d))) / T(2), lr, omlr);
    mu0[ld] = (fill_one ? T(1) : mu0[ld]) / rho;
  }
}

}  // namespace ifadv
