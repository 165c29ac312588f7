// ifadv_forcing.cuh -- explicit forcing between advection and projection:
//   viscSurfTenρu! = fill!(r,0) + visc! + surfTen!      src/flow.jl:113-152, src/VOFutil.jl:186-191, src/surfaceTension.jl:8-101
//   updateU!                                             src/flow.jl:244-252
//   fill!(μ₀,1) + updateL!                               src/flow.jl:73,96,254-259
// The reference runs D² + 2 gather/scatter passes for visc!, D × 4 passes for surfTen! and 5 passes for updateU!; here:
//   visc_kernel      one thread per momentum cell evaluates the viscous stress flux through its own lower and upper j-faces for every
//                    (i,j) from f and u directly -- fFace = f2face!(f) + BCv! is never materialised (closed form of the ghost rules),
//                    Φ is not used -- and writes r once;
//   visc3m_kernel    3-D: the same as a march along z from shared-memory planes, every face flux evaluated once;
//   stscan_kernel    one pass over f lists the interface cells of the D staggered fields f̄ = ϕ(d,·,f) (fbuffer is never
//                    materialised: its values, ghost rules of BCf!(d,·) included, are evaluated from f where the stencils need them);
//   surften_kernel   lane-dense over those lists: Weymouth-Yue normal in registers, major direction, 3^(D-1) Popinet column heights,
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

// The same for 3-D grids as a MARCH along z (the per-cell form re-reads ~200 values per cell through L2 and evaluates every face flux
// twice: 12.2 ms at 512³ Float32).  A CTA owns 32 x 8 columns and walks through `chunk` planes.  Per plane it stages the three
// staggered fields fFace[·,d] (ghost rules applied while loading, so every cell takes the same path) and the three velocity
// components with a one-cell halo in shared memory (ring of two planes), every thread evaluates the NINE lower-face fluxes of its cell
// exactly once, hands the x- and y-fluxes to its x-1 / y-1 neighbours through shared memory (one extra column / row of faces per tile)
// and keeps the z-flux for its own cell of the previous plane, whose sum it completes in the reference's order
// ((((((0 + Φx) - Φx⁺) + Φy) - Φy⁺) + Φz) - Φz⁺).  At a periodic upper boundary the reference takes Φ of plane 2 for plane N
// (upperBoundaryVisc!, Val{true}); the staged values of plane N are those of plane 2 when u carries periodic ghosts (BC!), which the
// entry point requires.
struct ViscM {
  static constexpr int TX = 32, TY = 8, NT = 256, PW = TX + 2, PH = TY + 2, PV = PW * PH;
  static constexpr int NFX = (TX + 1) * TY, NFY = TX * (TY + 1);
  template <class T> static constexpr size_t bytes() { return sizeof(T) * (size_t)(12 * PV + 3 * NFX + 3 * NFY); }
};
template <class T>
__global__ void __launch_bounds__(256) visc3m_kernel(T* __restrict__ r, const T* __restrict__ u, const T* __restrict__ f, const T* __restrict__ stale,
                                                     const Geo g, T lmu, T mu, T lr, int chunk) {
  constexpr int TX = ViscM::TX, TY = ViscM::TY, NT = ViscM::NT, PW = ViscM::PW, PV = ViscM::PV, NFX = ViscM::NFX, NFY = ViscM::NFY;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sFF = reinterpret_cast<T*>(smem_raw);  // [slot][d][PV]
  T* sU = sFF + 6 * PV;                     // [slot][c][PV]
  T* sFx = sU + 6 * PV;                     // [i][ty*(TX+1) + tx], tx = 0..TX
  T* sFy = sFx + 3 * NFX;                   // [i][ty*TX + tx],     ty = 0..TY
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int ox = 2 + blockIdx.x * TX, oy = 2 + blockIdx.y * TY;
  const int k0 = 2 + blockIdx.z * chunk, k1 = min(k0 + chunk, g.n[2]);  // planes k0 .. k1-1 are updated
  const T omlmu = T(1) - lmu, omlr = T(1) - lr, wlight = lmu / lr;
  // every staged index of this tile is inside(f) in x and y: plain averages, no ghost rules
  const bool fastxy = ox - 1 >= 2 && ox + TX <= g.n[0] - 1 && oy - 1 >= 2 && oy + TY <= g.n[1] - 1;
  auto load_plane = [&](int k) {
    const int slot = k & 1;
    const bool fast = fastxy && k >= 2 && k <= g.n[2] - 1;
    T* dF = sFF + slot * 3 * PV;
    T* dU = sU + slot * 3 * PV;
    for (int e = tid; e < PV; e += NT) {
      const int a = e % PW, b = e / PW;
      const int X = min(ox - 1 + a, g.n[0]), Y = min(oy - 1 + b, g.n[1]);
      const long long l = lin3(g, X, Y, k);
      if (fast) {
        const T fc = __ldg(f + l);
        dF[e] = (fc + __ldg(f + l - 1)) / T(2);
        dF[PV + e] = (fc + __ldg(f + l - g.s1)) / T(2);
        dF[2 * PV + e] = (fc + __ldg(f + l - g.s2)) / T(2);
      } else {
#pragma unroll
        for (int d = 0; d < 3; ++d) dF[d * PV + e] = fface<T, 3, false>(f, stale, g, X, Y, k, d);
      }
      dU[e] = __ldg(u + l);
      dU[PV + e] = __ldg(u + g.S + l);
      dU[2 * PV + e] = __ldg(u + 2 * g.S + l);
    }
  };
  // viscF(i,j,P) for the staged position p of plane k (slot s; sp = the slot of plane k-1)
  auto flux = [&](int i, int j, int p, int s, int sp) -> T {
    const T* F0 = sFF + s * 3 * PV;
    const T* F1 = sFF + sp * 3 * PV;
    const T* U0 = sU + s * 3 * PV;
    const T* U1 = sU + sp * 3 * PV;
    const int oj = (j == 0) ? 1 : PW, oi = (i == 0) ? 1 : PW;
    const T f1 = (j == 2) ? F1[i * PV + p] : F0[i * PV + p - oj];
    const T f2 = F0[i * PV + p];
    const T f3 = (i == j) ? f1 : ((i == 2) ? F1[j * PV + p] : F0[j * PV + p - oi]);
    const T f4 = (i == j) ? f2 : F0[j * PV + p];
    const T sv = (f1 + f2 + f3 + f4) / T(4);
    const T fm = (lr < T(1)) ? t_min(t_min(t_min(f1, f2), f3), f4) : t_max(t_max(t_max(f1, f2), f3), f4);
    const T w = (sv > T(0.5)) ? T(1) : wlight;
    const T muI = mu * t_min(lin_interp(sv, lmu, omlmu), w * lin_interp(fm, lr, omlr));
    const T uij = (j == 2) ? U1[i * PV + p] : U0[i * PV + p - oj];
    const T uji = (i == 2) ? U1[j * PV + p] : U0[j * PV + p - oi];
    const T du = (U0[i * PV + p] - uij) + (U0[j * PV + p] - uji);
    return muI * du;
  };
  const int x = ox + tx, y = oy + ty;
  const bool cellok = x <= g.n[0] - 1 && y <= g.n[1] - 1;
  const int p = (tx + 1) + PW * (ty + 1);
  T accp[3] = {T(0), T(0), T(0)};
  load_plane(k0 - 1);
#pragma unroll 1
  for (int k = k0; k <= k1; ++k) {
    load_plane(k);
    __syncthreads();  // plane k staged; the gathers of plane k-1 are done
    const int s = k & 1, sp = s ^ 1;
    const bool last = (k == k1);
    T Fl[3][3];  // [i][j]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Fl[i][2] = flux(i, 2, p, s, sp);
      if (!last) {
        Fl[i][0] = flux(i, 0, p, s, sp);
        Fl[i][1] = flux(i, 1, p, s, sp);
        sFx[i * NFX + ty * (TX + 1) + tx] = Fl[i][0];
        sFy[i * NFY + ty * TX + tx] = Fl[i][1];
      }
    }
    if (!last) {
      if (tid < TY) {  // the faces x = ox+TX of the tile's rows
        const int ph = (TX + 1) + PW * (tid + 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) sFx[i * NFX + tid * (TX + 1) + TX] = flux(i, 0, ph, s, sp);
      } else if (tid >= 32 && tid < 64) {  // the faces y = oy+TY of the tile's columns
        const int ph = (tx + 1) + PW * (TY + 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) sFy[i * NFY + TY * TX + tx] = flux(i, 1, ph, s, sp);
      }
      __syncthreads();  // the x / y fluxes of plane k are visible
    }
    if (k > k0 && cellok) {
      const long long l = lin3(g, x, y, k - 1);
#pragma unroll
      for (int i = 0; i < 3; ++i) r[(long long)i * g.S + l] = accp[i] - (-Fl[i][2]);
    }
    if (!last) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        T acc = T(0);
        acc = acc + (-Fl[i][0]);
        acc = acc - (-sFx[i * NFX + ty * (TX + 1) + tx + 1]);
        acc = acc + (-Fl[i][1]);
        acc = acc - (-sFy[i * NFY + (ty + 1) * TX + tx]);
        acc = acc + (-Fl[i][2]);
        accp[i] = acc;
      }
    }
  }
}

// ---- surface tension, surfaceTension.jl:8-101 ------------------------------------------------------------------------------------------
template <class T> IFADV_DI bool contain_interface(T f) { return T(0) < f && f < T(1); }  // VOFutil.jl:144

// fbuffer[J] as surfTen! sees it in its pass for direction d (surfaceTension.jl:10-11): ϕ(d,·,f) on inside(f) with the ghost rules of
// BCf!(d,·) -- evaluated from f on the fly, the array is not materialised.  Plane N_d of a non-periodic d is never written by that pass:
// there the reference still holds the previous direction's field (the caller's values for d = 1), hence the descent over d.
template <class T, int D> IFADV_DI T fbval(const T* __restrict__ f, const T* __restrict__ fbstale, const Geo& g, int x, int y, int z, int d) {
  int v[3] = {x, y, z};
  for (int dd = d; dd >= 0; --dd) {
    if (stag_map<D>(g, dd, v)) {
      const long long l = lin3(g, v[0], v[1], v[2]);
      return (__ldg(f + l) + __ldg(f + l - stride_of(g, dd))) / T(2);
    }
  }
  return __ldg(fbstale + lin3(g, v[0], v[1], v[2]));
}
template <class T, int D> struct SBox {  // 3^D box of fbuffer around a cell
  const T *f, *fbstale;
  const Geo& g;
  int x, y, z, d;
  IFADV_DI T operator()(int dx, int dy, int dz) const { return fbval<T, D>(f, fbstale, g, x + dx, y + dy, z + dz, d); }
};

// getPopinetHeightAdaptive(I,f,i,monotonic=true), surfaceTension.jl:76-99; sd: signed 1-based direction of the walk
template <class T, int D> IFADV_DI T popinet_height(const T* __restrict__ f, const T* __restrict__ fbstale, const Geo& g, int x, int y, int z, int d,
                                                    int sd) {
  const int a = (sd < 0 ? -sd : sd) - 1, sg = (sd < 0) ? -1 : 1;
  const int n = g.n[a];
  const int c0 = (a == 0) ? x : (a == 1 ? y : z);
  auto at = [&](int c) -> T { return fbval<T, D>(f, fbstale, g, a == 0 ? c : x, a == 1 ? c : y, a == 2 ? c : z, d); };
  const T f0 = at(c0);
  T H = f0 - T(0.5);
  {
    T fnow = f0;
    bool fin = fnow < T(1);
    int c = c0;
    while (!fin || contain_interface(fnow)) {
      c += sg;
      if (c < 1 || c > n) break;
      const T fi = at(c);
      fnow = (fi > fnow) ? T(0) : fi;
      H += fnow;
      fin = contain_interface(fnow) ? true : fin;
    }
  }
  {
    T fnow = f0;
    bool fin = fnow > T(0);
    int c = c0;
    while (!fin || contain_interface(fnow)) {
      c -= sg;
      if (c < 1 || c > n) break;
      const T fi = at(c);
      fnow = (fi < fnow) ? T(1) : fi;
      H += fnow - T(1);
      fin = contain_interface(fnow) ? true : fin;
    }
  }
  return H;
}
template <class T> IFADV_DI T root1p5(T a) { return t_sqrt(a * a * a); }  // surfaceTension.jl:101

// getCurvature(I,f,i), surfaceTension.jl:31-65
template <class T, int D> IFADV_DI T curvature(const T* __restrict__ f, const T* __restrict__ fbstale, const Geo& g, int x, int y, int z, int d, int sd) {
  const int ai = sd < 0 ? -sd : sd, sg = sd < 0 ? -1 : 1;
  if (D == 3) {
    const int ix = sg * (ai % 3 + 1), iy = (ai + 1) % 3 + 1;  // getXYdir, util.jl:65
    const int ax = (ix < 0 ? -ix : ix) - 1, sx = ix < 0 ? -1 : 1, ay = iy - 1;
    T H[3][3];
#pragma unroll 1
    for (int a = -1; a <= 1; ++a)
#pragma unroll 1
      for (int b = -1; b <= 1; ++b) {
        int v[3] = {x, y, z};
        v[ax] += a * sx;
        v[ay] += b;
        H[a + 1][b + 1] = popinet_height<T, D>(f, fbstale, g, v[0], v[1], v[2], d, sd);
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
#pragma unroll 1
  for (int a = -1; a <= 1; ++a) {
    int v[3] = {x, y, z};
    v[ax] += a * sx;
    H[a + 1] = popinet_height<T, D>(f, fbstale, g, v[0], v[1], v[2], d, sd);
  }
  const T Hx = (H[2] - H[0]) / T(2);
  const T Hxx = H[2] + H[0] - T(2) * H[1];
  return Hxx / root1p5(T(1) + Hx * Hx);
}

// One pass over inside(f): the interface cells of the D staggered fields (calNormal! / applySurfTen! act on those only) go to D lists,
// one warp-aggregated atomic per warp and direction.  cnt[d] counts every hit; entries beyond cap are dropped (the consumer then scans
// all cells of that direction).
template <class T, int D> __global__ void __launch_bounds__(128) stscan_kernel(const T* __restrict__ f, const Geo g, int* __restrict__ list,
                                                                              unsigned* __restrict__ cnt, unsigned cap) {
  const int x = 2 + blockIdx.x * blockDim.x + threadIdx.x;
  const int nrow = (int)gridDim.y * 4;
  (void)nrow;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    const int y = 2 + blockIdx.y * 4 + q, z = (D == 3) ? 2 + blockIdx.z : 1;
    const bool in = x <= g.n[0] - 1 && y <= g.n[1] - 1;
    const long long l = in ? lin3(g, x, y, z) : 0;
    const T fc = in ? __ldg(f + l) : T(0);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const T val = in ? (fc + __ldg(f + l - stride_of(g, d))) / T(2) : T(0);
      const bool hit = in && contain_interface(val);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(cnt + d, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        const unsigned k = base + (unsigned)__popc(m & ((1u << lane) - 1u));
        if (hit && k < cap) list[(size_t)d * cap + k] = (int)l;
      }
    }
  }
}

// calNormal! + applySurfTen! for one interface cell of direction d, surfaceTension.jl:12-21
template <class T, int D> IFADV_DI void surften_cell(T* __restrict__ r, const T* __restrict__ f, const T* __restrict__ fbstale, const Geo& g, int d, T eta,
                                                     int x, int y, int z) {
  SBox<T, D> B{f, fbstale, g, x, y, z, d};
  T n[3] = {T(0), T(0), T(0)};
  normal_wy<T, D>(B, n);                    // getInterfaceNormal_WY!(fbuffer,n̂,I)
  const int im = arg_abs_max<T, D>(n);      // majorDir(n̂,I), util.jl:72-75
  const int sd = t_signbit(pick(n, im)) ? -(im + 1) : (im + 1);
  const T kappa = curvature<T, D>(f, fbstale, g, x, y, z, d, sd);
  const long long l = lin3(g, x, y, z), ld = (long long)d * g.S + l;
  r[ld] = r[ld] + eta * kappa * -(__ldg(f + l) - __ldg(f + l - stride_of(g, d)));
}
// Lane-dense over the lists (persistent grid; the counts stay on the device).  A direction whose list overflowed scans every cell.
template <class T, int D> __global__ void __launch_bounds__(128) surften_kernel(T* __restrict__ r, const T* __restrict__ f, const T* __restrict__ fbstale,
                                                                               const Geo g, T eta, const int* __restrict__ list,
                                                                               const unsigned* __restrict__ cnt, unsigned cap) {
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 1
  for (int d = 0; d < D; ++d) {
    const unsigned n = cnt[d];
    if (n <= cap) {
      for (long long k = t0; k < (long long)n; k += stride) {
        const long long l = list[(size_t)d * cap + k];
        const int z = (D == 3) ? (int)(l / g.s2) : 0;
        const long long q = l - (long long)z * ((D == 3) ? g.s2 : 0);
        const int y = (int)(q / g.s1), x = (int)(q - (long long)y * g.s1);
        surften_cell<T, D>(r, f, fbstale, g, d, eta, x + 1, y + 1, (D == 3) ? z + 1 : 1);
      }
    } else {
      const long long nx = g.n[0] - 2, ny = g.n[1] - 2, nz = (D == 3) ? g.n[2] - 2 : 1;
      for (long long k = t0; k < nx * ny * nz; k += stride) {
        const int x = 2 + (int)(k % nx), y = 2 + (int)((k / nx) % ny), z = (D == 3) ? 2 + (int)(k / (nx * ny)) : 1;
        const long long l = lin3(g, x, y, z);
        if (contain_interface((__ldg(f + l) + __ldg(f + l - stride_of(g, d))) / T(2))) surften_cell<T, D>(r, f, fbstale, g, d, eta, x, y, z);
      }
    }
  }
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
