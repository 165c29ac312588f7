// ifadv_vofcell.cuh -- pure-VOF directional sweep (advect! / advectVOF!, src/advection.jl:34-137) for 3-D grids: CELL-PARALLEL.
//
// The CMOM sweeps march through rings of staged planes because a momentum cell needs a dozen neighbour values per direction.  A pure
// VOF cell needs five -- f of the cell and of its two neighbours along the sweep, the face velocity below and above -- and all of them
// are coalesced row reads, so the sweep is two kernels without shared memory, barriers or staging:
//   vofcell_kernel      a thread owns one (x,y) column of a chunk of z-planes; a CTA is 32 (x) x 8 (y) threads.  Per cell: δl of its
//                       two faces (advection.jl:110), the upwind cell's f (trivial flux f·δl, advection.jl:125-129), the dilation
//                       c̄(∂u+∂u⁰)δt/2, the update and cleanWisp! (advection.jl:83, VOFutil.jl:127-136), the fill-error extrema.
//                       Sweeps along z carry the upper face's flux to the next plane, sweeps along x / y evaluate both faces of a
//                       cell (the neighbour's copy is an L1 hit).  A cell with a face whose upwind cell holds an interface (a
//                       fraction of a percent) is NOT finished here: it goes to a list (one warp-aggregated atomic);
//   vofcell_fix_kernel  lane-dense over that list (persistent grid, the count stays on the device): the same cell update with the
//                       PLIC reconstruction of the flagged faces from the 3^3 box around the upwind cell (normal scheme -> intercept
//                       -> volume under the shifted plane, advection.jl:131-134).  Reconstructing in line would cost a divergent
//                       warp ~4000 issue slots per interface cell against ~50 for a plain cell (measured: 8x slower sweeps).
// Ghost layers of the ping-pong buffers are never read: f of a neighbour outside the interior comes from the cell BCf! would have
// copied (clamp = Neumann, wrap = periodic).  Same arithmetic, expression by expression, as the MOM = false instantiations of
// ifadv_along2.cuh / ifadv_xrow.cuh (bit-identical in Float64).  Algorithmic bytes: 4s+1 per cell (f, u_j, c̄ in; f out).
#pragma once
#include "ifadv_along2.cuh"

namespace ifadv {

template <class T> struct VFace { T ff, dl; bool plic; };
template <class T> struct Box27 {  // 3^3 box held by the thread
  T v[27];
  IFADV_DI T operator()(int dx, int dy, int dz) const { return v[(dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)]; }
};
template <class T> struct Box9 {  // 3^2 box (2-D grids)
  T v[9];
  IFADV_DI T operator()(int dx, int dy, int) const { return v[(dx + 1) + 3 * (dy + 1)]; }
};

// VOF flux through the lower face of the cell with index v along J (cells v-1 | v), advection.jl:108-137.  PLIC = false: a face that
// needs the reconstruction is only flagged.
template <class T, int J, bool PLIC, int D = 3>
IFADV_DI VFace<T> vof_face(const SweepP<T>& P, T usum, T flo, T fhi, int v, int cx, int cy, int cz) {
  const int nA = P.g.n[J];
  const bool perA = (P.g.per >> J) & 1u;
  T dl = P.hdt * usum;             // δt/2*(u+u⁰)
  dl = (dl != T(0)) ? dl : T(0);   // -0 -> +0: the zero-flux case of advection.jl:115 without a branch
  const bool up = dl > T(0);
  const T fc = up ? flo : fhi;     // upwind cell
  const int cu = up ? v - 1 : v;
  const bool gho = !perA && (cu < 2 || cu > nA - 1);  // ghost upwind cell on a non-periodic side: trivial flux
  const bool need = dl != T(0) && !gho && !fullorempty(fc);
  T ff = fc * dl;
  if (PLIC && need) {
    const int m = map1(cu, nA, perA);
    // the 3^D box around the upwind cell: all loads independent and in flight together (the normal schemes would otherwise pull them in
    // one by one through data-dependent branches), ghost rules folded into three index maps per axis
    const int bx = (J == 0) ? m : cx, by = (J == 1) ? m : cy, bz = (J == 2) ? m : cz;
    long long ox[3], oy[3], oz[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      ox[a] = mapc(bx + a - 1, P.g.n[0], P.g.per & 1u) - 1;
      oy[a] = (long long)(mapc(by + a - 1, P.g.n[1], P.g.per & 2u) - 1) * P.g.s1;
      oz[a] = (D == 3) ? (long long)(mapc(bz + a - 1, P.g.n[2], P.g.per & 4u) - 1) * P.g.s2 : 0;
    }
    if (D == 3) {
      Box27<T> B;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int a = 0; a < 3; ++a) B.v[a + 3 * b + 9 * c] = __ldg(P.f_in + ox[a] + oy[b] + oz[c]);
      ff = plic_face_flux_inl<T, D>(P.scheme, B, fc, J, dl);
    } else {
      Box9<T> B;
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int a = 0; a < 3; ++a) B.v[a + 3 * b] = __ldg(P.f_in + ox[a] + oy[b]);
      ff = plic_face_flux_inl<T, D>(P.scheme, B, fc, J, dl);
    }
  }
  return VFace<T>{ff, dl, need};
}

// update of one cell from its two face fluxes (advection.jl:83 + cleanWisp!), the optional ρuf output and the fill-error extrema
template <class T, int J>
IFADV_DI void vof_cell_finish(const SweepP<T>& P, long long l, int cJ, T fc, const VFace<T>& lo, const VFace<T>& hi, T dv, T& rmax, T& rmin,
                              unsigned& amax, unsigned& amin) {
  const long long sA = (J == 0) ? 1 : ((J == 1) ? P.g.s1 : P.g.s2);
  const unsigned lk = (unsigned)l;
  T fn = fc + ((lo.ff - hi.ff) + dv);
  rmax = max_nan(rmax, fn);
  rmin = t_min(rmin, fn);
  if (fn > T(1) || fn < T(0)) {  // only cells outside [0,1] can be reported (reportFillError, advection.jl:145-189)
    if (fn >= rmax) amax = lk;
    if (fn <= rmin) amin = lk;
  }
  fn = (fn < P.tol) ? T(0) : ((fn > P.onemtol) ? T(1) : fn);  // cleanWisp!
  P.f_out[l] = fn;
  if (P.rhouf_j != nullptr) {  // ρuf[·,d] on inside_uWB faces: δl·λρ + (1-λρ)fᶠ, VOFutil.jl:218
    P.rhouf_j[l] = lo.dl * P.lr + P.omlr * lo.ff;
    if (cJ == P.g.n[J] - 1) P.rhouf_j[l + sA] = hi.dl * P.lr + P.omlr * hi.ff;
  }
}

template <class T> IFADV_DI void vof_reduce(const SweepP<T>& P, T rmax, T rmin, unsigned amax, unsigned amin) {
  if (P.red == nullptr) return;
  int rnan = 0;
  if (rmax != rmax) { rnan = 1; rmax = -INFINITY; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const T omax = __shfl_xor_sync(0xffffffffu, rmax, off), omin = __shfl_xor_sync(0xffffffffu, rmin, off);
    const unsigned int oamax = __shfl_xor_sync(0xffffffffu, amax, off), oamin = __shfl_xor_sync(0xffffffffu, amin, off);
    const int onan = __shfl_xor_sync(0xffffffffu, rnan, off);
    if (omax > rmax) { rmax = omax; amax = oamax; }
    if (omin < rmin) { rmin = omin; amin = oamin; }
    rnan |= onan;
  }
  if ((threadIdx.x & 31) == 0) red_commit<T>(P.red, rmax, rmin, amax, amin, rnan);
}

// EDGE = false: every cell of the CTA's tile and both neighbours along the sweep are inside(f) and the tile is full -- no index maps,
// no ghost upwind cells, no lane shadows (chosen per CTA; all but the outermost tiles).  KB: planes whose loads are issued together.
template <class T, int J, bool SAMEU, bool EDGE, int KB>
IFADV_DI void vofcell_body(const SweepP<T>& P, const int chunk, int* __restrict__ list, unsigned* __restrict__ cnt, const unsigned cap) {
  const Geo& g = P.g;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int xr = 2 + blockIdx.x * 32 + tx, yr = 2 + blockIdx.y * 8 + ty;
  const bool ok = !EDGE || (xr <= g.n[0] - 1 && yr <= g.n[1] - 1);
  const int x = EDGE ? min(xr, g.n[0] - 1) : xr, y = EDGE ? min(yr, g.n[1] - 1) : yr;  // lanes beyond the grid shadow the last cell
  const int k0 = 2 + blockIdx.z * chunk, k1 = min(k0 + chunk, g.n[2]);  // planes k0 .. k1-1
  const int nA = g.n[J];
  const bool perA = (g.per >> J) & 1u;
  const unsigned s1 = (unsigned)g.s1, s2 = (unsigned)g.s2;
  const unsigned sA = (J == 0) ? 1u : ((J == 1) ? s1 : s2);
  const bool first = P.first != 0;
  const T dt = P.dt, hdt = P.hdt;
  const T* __restrict__ fin = P.f_in;
  const T* __restrict__ uj = P.uj;
  const T* __restrict__ u0j = P.u0j;
  T rmax = -INFINITY, rmin = INFINITY;
  unsigned int amax = 0, amin = 0;
  const int lane = threadIdx.x & 31;

  const unsigned lxy = (unsigned)(x - 1) + s1 * (unsigned)(y - 1);  // 32-bit element offsets (S < 2^31 is checked by the launcher)
  // J = 0, 1: the neighbour offsets along the sweep and the ghost status of the upwind candidates do not change with the plane
  const int cJ01 = (J == 0) ? x : y;
  const unsigned om = (J == 2 || !EDGE) ? (0u - sA) : (unsigned)(map1(cJ01 - 1, nA, perA) - cJ01) * sA;
  const unsigned op = (J == 2 || !EDGE) ? sA : (unsigned)(map1(cJ01 + 1, nA, perA) - cJ01) * sA;
  const bool gdn01 = EDGE && J != 2 && !perA && cJ01 - 1 < 2;       // lower neighbour is a ghost cell on a non-periodic side
  const bool gup01 = EDGE && J != 2 && !perA && cJ01 + 1 > nA - 1;  // upper neighbour
  const int kup = perA ? 2 : nA - 1;  // J = 2: the plane that stands for plane nA (wrap / clamp)
  // flux through a face between the cells `flo | fhi`; glo / ghi: that cell is a ghost cell on a non-periodic side
  auto face = [&](T usum, T flo, T fhi, bool glo, bool ghi) -> VFace<T> {
    T dl = hdt * usum;
    dl = (dl != T(0)) ? dl : T(0);
    const bool up = dl > T(0);
    const T fc = up ? flo : fhi;
    const bool gho = EDGE && (up ? glo : ghi);
    return VFace<T>{fc * dl, dl, dl != T(0) && !gho && !fullorempty(fc)};
  };
  struct In { T fc, fm, fp, ulo, uhi, u0lo, u0hi; int cb; };
  auto load = [&](int k) -> In {
    In r;
    const unsigned l = lxy + s2 * (unsigned)(k - 1);
    if (J == 2) {
      r.fp = __ldg(fin + (lxy + s2 * (unsigned)(((!EDGE || k + 1 <= nA - 1) ? k + 1 : kup) - 1)));
      r.uhi = __ldg(uj + (l + s2));
      r.u0hi = SAMEU ? r.uhi : __ldg(u0j + (l + s2));
      r.fc = r.fm = r.ulo = r.u0lo = T(0);  // rolled
    } else {
      r.fc = __ldg(fin + l);
      r.fm = __ldg(fin + (l + om));
      r.fp = __ldg(fin + (l + op));
      r.ulo = __ldg(uj + l);
      r.uhi = __ldg(uj + (l + sA));
      r.u0lo = SAMEU ? r.ulo : __ldg(u0j + l);
      r.u0hi = SAMEU ? r.uhi : __ldg(u0j + (l + sA));
    }
    r.cb = first ? 0 : (int)P.cbar[l];
    return r;
  };
  // J = 2: rolling state (f of planes k-1, k; flux and velocity of the lower face)
  T fm = T(0), fc = T(0), ulo = T(0), u0lo = T(0);
  VFace<T> lo{T(0), T(0), false};
  if (J == 2) {
    const unsigned l = lxy + s2 * (unsigned)(k0 - 1);
    fm = __ldg(fin + (lxy + s2 * (unsigned)(map1(k0 - 1, nA, perA) - 1)));
    fc = __ldg(fin + l);
    ulo = __ldg(uj + l);
    u0lo = SAMEU ? ulo : __ldg(u0j + l);
    lo = face(ulo + u0lo, fm, fc, !perA && k0 - 1 < 2, false);
  }
#pragma unroll 1
  for (int kb = k0; kb < k1; kb += KB) {
    In in[KB];
#pragma unroll
    for (int i = 0; i < KB; ++i) in[i] = load(min(kb + i, k1 - 1));
#pragma unroll
    for (int i = 0; i < KB; ++i) {
      const int k = kb + i;
      if (k < k1) {  // block-uniform
        const unsigned l = lxy + s2 * (unsigned)(k - 1);
        const In& cur = in[i];
        T fp = cur.fp, uhi = cur.uhi, u0hi = cur.u0hi;
        VFace<T> hi;
        if (J == 2) {
          hi = face(uhi + u0hi, fc, fp, false, !perA && k + 1 > nA - 1);
        } else {
          fc = cur.fc; fm = cur.fm; ulo = cur.ulo; u0lo = cur.u0lo;
          lo = face(ulo + u0lo, fm, fc, gdn01, false);
          hi = face(uhi + u0hi, fc, fp, false, gup01);
        }
        const int cb = first ? ((fc < T(0.5)) ? 0 : 1) : cur.cb;  // advection.jl:40 (c̄ from the incoming f)
        if (first && ok) P.cbar[l] = (int8_t)cb;
        const bool defer = ok && (lo.plic || hi.plic);
        if (ok && !defer) {
          const T div = (uhi - ulo) + (u0hi - u0lo);  // ∂(d,I,u)+∂(d,I,u⁰)
          const T dv = ((cb ? div : T(0)) * dt) / T(2);
          vof_cell_finish<T, J>(P, (long long)l, (J == 2) ? k : cJ01, fc, lo, hi, dv, rmax, rmin, amax, amin);
        }
        const unsigned m = __ballot_sync(0xffffffffu, defer);
        if (m) {
          const int leader = __ffs(m) - 1;
          unsigned base = 0;
          if (lane == leader) base = atomicAdd(cnt, (unsigned)__popc(m));
          base = __shfl_sync(0xffffffffu, base, leader);
          const unsigned q = base + (unsigned)__popc(m & ((1u << lane) - 1u));
          if (defer && q < cap) list[q] = (int)l;
        }
        if (J == 2) { fm = fc; fc = fp; ulo = uhi; u0lo = u0hi; lo = hi; }
      }
    }
  }
  vof_reduce<T>(P, rmax, rmin, amax, amin);
}

template <class T, int J, bool SAMEU, int KB, int MINB>
__global__ void __launch_bounds__(256, MINB) vofcell_kernel(const SweepP<T> P, const int chunk, int* __restrict__ list, unsigned* __restrict__ cnt,
                                                            const unsigned cap) {
  const Geo& g = P.g;
  const int ox = 2 + blockIdx.x * 32, oy = 2 + blockIdx.y * 8, k0 = 2 + blockIdx.z * chunk;
  // the tile is full and neither it nor its neighbours along the sweep touch a ghost index
  bool interior = ox + 31 <= g.n[0] - 1 && oy + 7 <= g.n[1] - 1;
  if (J == 0) interior = interior && ox >= 3 && ox + 32 <= g.n[0] - 1;
  if (J == 1) interior = interior && oy >= 3 && oy + 8 <= g.n[1] - 1;
  if (J == 2) interior = interior && k0 >= 3 && k0 + chunk <= g.n[2] - 1;
  if (interior) vofcell_body<T, J, SAMEU, false, KB>(P, chunk, list, cnt, cap);
  else vofcell_body<T, J, SAMEU, true, KB>(P, chunk, list, cnt, cap);
}

// the deferred cells, lane-dense.  When the list overflowed every cell is examined again instead.
template <class T, int J, bool SAMEU>
__global__ void __launch_bounds__(128) vofcell_fix_kernel(const SweepP<T> P, const int* __restrict__ list, const unsigned* __restrict__ cnt,
                                                          const unsigned cap) {
  const Geo& g = P.g;
  const int nA = g.n[J];
  const bool perA = (g.per >> J) & 1u;
  const long long sA = (J == 0) ? 1 : ((J == 1) ? g.s1 : g.s2);
  const bool first = P.first != 0;
  T rmax = -INFINITY, rmin = INFINITY;
  unsigned int amax = 0, amin = 0;
  const unsigned n = *cnt;
  const bool scan = n > cap;
  const long long nx = g.n[0] - 2, ny = g.n[1] - 2, nz = g.n[2] - 2;
  const long long total = scan ? nx * ny * nz : (long long)n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
    int x, y, z;
    if (scan) { x = 2 + (int)(q % nx); y = 2 + (int)((q / nx) % ny); z = 2 + (int)(q / (nx * ny)); }
    else {
      const long long l0 = list[q];
      const int zz = (int)(l0 / g.s2);
      const long long r0 = l0 - (long long)zz * g.s2;
      const int yy = (int)(r0 / g.s1);
      x = (int)(r0 - (long long)yy * g.s1) + 1; y = yy + 1; z = zz + 1;
    }
    const long long l = lin3(g, x, y, z);
    const int cJ = (J == 0) ? x : ((J == 1) ? y : z);
    const T fc = __ldg(P.f_in + l);
    const T fm = __ldg(P.f_in + l + (long long)(map1(cJ - 1, nA, perA) - cJ) * sA);
    const T fp = __ldg(P.f_in + l + (long long)(map1(cJ + 1, nA, perA) - cJ) * sA);
    const T ulo = __ldg(P.uj + l), uhi = __ldg(P.uj + l + sA);
    const T u0lo = SAMEU ? ulo : __ldg(P.u0j + l), u0hi = SAMEU ? uhi : __ldg(P.u0j + l + sA);
    const VFace<T> lo = vof_face<T, J, true>(P, ulo + u0lo, fm, fc, cJ, x, y, z);
    const VFace<T> hi = vof_face<T, J, true>(P, uhi + u0hi, fc, fp, cJ + 1, x, y, z);
    if (scan && !(lo.plic || hi.plic)) continue;  // finished by vofcell_kernel
    const int cb = first ? ((fc < T(0.5)) ? 0 : 1) : (int)P.cbar[l];
    const T div = (uhi - ulo) + (u0hi - u0lo);
    const T dv = ((cb ? div : T(0)) * P.dt) / T(2);
    vof_cell_finish<T, J>(P, l, cJ, fc, lo, hi, dv, rmax, rmin, amax, amin);
  }
  vof_reduce<T>(P, rmax, rmin, amax, amin);
}

// One directional sweep of a 2-D grid inside a cooperative kernel (the whole advect! step of small grids is ONE launch): phase 1 = every
// cell the cell-parallel way, cells next to an interface face go to `list`; grid barrier; phase 2 = the listed cells lane-dense with the
// PLIC reconstruction; grid barrier.  The grid-stride loops are warp-uniform (blockDim and gridDim·blockDim are multiples of 32).
// LOCAL (small grids, where the step is bound by latency and grid barriers, not by issue slots): a flagged cell is reconstructed in line
// by its own thread right away -- no list, one grid barrier less per sweep.
template <class T, int J, bool SAMEU, bool LOCAL, class GRID>
IFADV_DI void vof2d_cell_sweep(const SweepP<T>& P, GRID& grid, int* __restrict__ list, unsigned* cnt, const unsigned cap) {
  const Geo& g = P.g;
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x, gs = (long long)gridDim.x * blockDim.x;
  const int nx = g.n[0] - 2, ny = g.n[1] - 2, nA = g.n[J], lane = threadIdx.x & 31;
  const bool perA = (g.per >> J) & 1u, first = P.first != 0;
  const long long ncell = (long long)nx * ny, nround = (ncell + 31) / 32 * 32, sA = (J == 0) ? 1 : g.s1;
  T rmax = -INFINITY, rmin = INFINITY;
  unsigned int amax = 0, amin = 0;
  auto cell = [&](long long c, bool plic_pass, bool ok) -> bool {  // returns: the cell is deferred (phase 1)
    const int x = 2 + (int)(c % nx), y = 2 + (int)(c / nx);
    const long long l = lin3(g, x, y, 1);
    const int cJ = (J == 0) ? x : y;
    const T fc = __ldg(P.f_in + l);
    const T fm = __ldg(P.f_in + l + (long long)(map1(cJ - 1, nA, perA) - cJ) * sA);
    const T fp = __ldg(P.f_in + l + (long long)(map1(cJ + 1, nA, perA) - cJ) * sA);
    const T ulo = __ldg(P.uj + l), uhi = __ldg(P.uj + l + sA);
    const T u0lo = SAMEU ? ulo : __ldg(P.u0j + l), u0hi = SAMEU ? uhi : __ldg(P.u0j + l + sA);
    VFace<T> lo, hi;
    if (plic_pass) { lo = vof_face<T, J, true, 2>(P, ulo + u0lo, fm, fc, cJ, x, y, 1); hi = vof_face<T, J, true, 2>(P, uhi + u0hi, fc, fp, cJ + 1, x, y, 1); }
    else { lo = vof_face<T, J, false, 2>(P, ulo + u0lo, fm, fc, cJ, x, y, 1); hi = vof_face<T, J, false, 2>(P, uhi + u0hi, fc, fp, cJ + 1, x, y, 1); }
    const bool flagged = lo.plic || hi.plic;
    const int cb = first ? ((fc < T(0.5)) ? 0 : 1) : (int)P.cbar[l];
    if (!plic_pass && first && ok) P.cbar[l] = (int8_t)cb;
    if (ok && (plic_pass ? flagged : !flagged)) {
      const T div = (uhi - ulo) + (u0hi - u0lo);
      const T dv = ((cb ? div : T(0)) * P.dt) / T(2);
      vof_cell_finish<T, J>(P, l, cJ, fc, lo, hi, dv, rmax, rmin, amax, amin);
    }
    return ok && flagged;
  };
  for (long long c = gt; c < nround; c += gs) {
    const bool ok = c < ncell;
    const bool defer = cell(ok ? c : 0, false, ok);
    if (LOCAL) {
      if (defer) cell(c, true, true);
      continue;
    }
    const unsigned m = __ballot_sync(0xffffffffu, defer);
    if (m) {
      const int leader = __ffs(m) - 1;
      unsigned base = 0;
      if (lane == leader) base = atomicAdd(cnt, (unsigned)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      const unsigned q = base + (unsigned)__popc(m & ((1u << lane) - 1u));
      if (defer && q < cap) list[q] = (int)(ok ? c : 0);
    }
  }
  if (!LOCAL) {
    grid.sync();
    const unsigned n = *(volatile unsigned*)cnt;
    if (n <= cap) {
      for (long long q = gt; q < (long long)n; q += gs) cell(list[q], true, true);
    } else {  // the list overflowed: examine every cell again
      for (long long c = gt; c < ncell; c += gs) cell(c, true, true);
    }
  }
  vof_reduce<T>(P, rmax, rmin, amax, amin);
  grid.sync();
}

}  // namespace ifadv
