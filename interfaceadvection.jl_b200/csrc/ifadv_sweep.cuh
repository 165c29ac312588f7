// ifadv_sweep.cuh -- the fused directional sweep kernel (sm_100a).
//
// ONE kernel per directional sweep replaces the ≈27 full-field + ≈80 boundary-plane launches of the
// reference: interface reconstruction (normalEstimation.jl:10-28), PLIC face
// flux (advection.jl:108-137), volume-fraction update with dilation (advection.jl:83), wisp snapping
// (VOFutil.jl:127-136), fill-error reduction (advection.jl:145-149) and -- when MOM -- the CMOM momentum
// transport with the SynDRoM limiter (flow.jl:197-231), including every boundary rule the reference applies
// through BC!/BCf!/BCv!/BCVOF! between its passes.  None of the reference's intermediates (u★, Φ, n̂, α, fᶠ,
// ρuf, dρ, ρ̄∂ⱼuⱼ, r) ever reach HBM: a CTA stages the face fluxes and u★ of its tile (+halo) in shared
// memory and finishes the cell update from there.
//
// Ghost handling: every read of f / ρu / c̄ goes through an index map that sends a ghost or out-of-range
// index to the interior cell whose value BCf!/BC! would have copied there (clamp = Neumann, wrap =
// periodic), so intermediate buffers never need their ghost layers filled.
#pragma once
#include "ifadv_math.cuh"

namespace ifadv {

struct Geo {
  int n[3];          // array extents incl. ghosts (n[2] = 1 for D == 2)
  long long s1, s2;  // strides of dims 1 and 2 (dim 0 has stride 1)
  long long S;       // elements of a scalar field
  unsigned per;      // periodic mask
};

IFADV_DI int wrapc(int v, int n) {
  const int m = n - 2;
  while (v < 2) v += m;
  while (v > n - 1) v -= m;
  return v;
}
IFADV_DI int clampc(int v, int n) { return min(max(v, 2), n - 1); }
IFADV_DI int mapc(int v, int n, bool per) { return per ? wrapc(v, n) : clampc(v, n); }
IFADV_DI long long lin3(const Geo& g, int x, int y, int z) { return (long long)(x - 1) + g.s1 * (y - 1) + g.s2 * (z - 1); }

template <class T> struct SweepP {
  const T* f_in;
  T* f_out;
  const T* u;     // u[:, 1] (all components)
  const T* u0;    // u⁰[:, 1]
  const T* uj;    // u[:, j]
  const T* u0j;   // u⁰[:, j]
  int8_t* cbar;   // written when first, read otherwise
  const T* rhou_in;
  T* rhou_out;
  const T* uOld;
  const T* drho;
  const T* uexit;  // exitBC (BC!'s saveexit, flow.jl:197,207): the array whose plane N of component x holds the saved exit value of u★; else null
  T* rhouf_j;  // optional (pure VOF): ρuf[:, j]
  T dt, hdt, idt, lr, omlr, tol, onemtol;
  T A[3];
  long long coff[3];  // component offsets d*S
  Geo g;
  int scheme, lim, first;
  int fused;  // sweep 1 of the fused entry: ρu_in = BC!(uOld*ρ(f̄)) is formed on the fly (u2ρu! + BC! folded in)
  unsigned long long* red;  // [0] max key, [1] min key, [2] argmax pack, [3] argmin pack, [4] nan count
  int kz0, kz1;  // planes [kz0, kz1) of dimension 3 this launch updates (1-based): 2..n[2] on one GPU; the owned planes of a z-slab,
                 // whose ghost planes (filled by the exchange) are read like any other interior plane   [3-D lean kernels only]
};

// order-preserving map double -> uint64 (for atomicMax/atomicMin)
IFADV_DI unsigned long long ord_key(double v) {
  long long b = __double_as_longlong(v);
  return (b < 0) ? (unsigned long long)(~b) : ((unsigned long long)b | 0x8000000000000000ull);
}
IFADV_DI unsigned int ord_key32(float v) {
  int b = __float_as_int(v);
  return (b < 0) ? (unsigned int)(~b) : ((unsigned int)b | 0x80000000u);
}

// Commit a warp's fill-error extrema.  Almost every warp ends with max = 1, min = 0 exactly, so look before issuing an atomic: tens of
// thousands of warps hitting ONE address serialise at the L2 (measured on the pure-VOF sweep at 256³: 0.043 ms of a 0.17 ms launch).
template <class T> IFADV_DI void red_commit(unsigned long long* red, T rmax, T rmin, unsigned amax, unsigned amin, int rnan) {
  const volatile unsigned long long* vr = red;
  if (rmax > -INFINITY) {
    const unsigned long long k = ord_key((double)rmax), kp = ((unsigned long long)ord_key32((float)rmax) << 32) | amax;
    if (k > vr[0]) atomicMax(red + 0, k);
    if (kp > vr[2]) atomicMax(red + 2, kp);
  }
  if (rmin < INFINITY) {
    const unsigned long long k = ord_key((double)rmin), kp = ((unsigned long long)ord_key32((float)rmin) << 32) | amin;
    if (k < vr[1]) atomicMin(red + 1, k);
    if (kp < vr[3]) atomicMin(red + 3, kp);
  }
  if (rnan) atomicAdd(red + 4, 1ull);
}

template <class T, int D> struct FMap {  // f with the BCf! ghost rule applied through the index map
  const T* f;
  Geo g;
  IFADV_DI T operator()(int x, int y, int z) const {
    x = mapc(x, g.n[0], g.per & 1u);
    y = mapc(y, g.n[1], g.per & 2u);
    z = (D == 3) ? mapc(z, g.n[2], g.per & 4u) : 1;
    return __ldg(f + lin3(g, x, y, z));
  }
};
template <class T, int D> struct FBox {  // 3^D box accessor around a cell for the normal schemes
  FMap<T, D> F;
  int cx, cy, cz;
  IFADV_DI T operator()(int dx, int dy, int dz) const { return F(cx + dx, cy + dy, cz + dz); }
};

// Tile geometry of one CTA for sweep direction J.
template <int D, int J, int TX, int TY, int TZ> struct Tile {
  static constexpr int T0 = TX, T1 = TY, T2 = (D == 3) ? TZ : 1;
  static constexpr int CELLS = T0 * T1 * T2;
  // face region: [-1, T_J] along J, [-1, T_k) across
  static constexpr int F0 = T0 + ((J == 0) ? 2 : 1), F1 = T1 + ((J == 1) ? 2 : 1), F2 = (D == 3) ? T2 + ((J == 2) ? 2 : 1) : 1;
  static constexpr int FN = F0 * F1 * F2;
  // u★ region: [-2, T_J+1] along J, tile across
  static constexpr int U0 = T0 + ((J == 0) ? 4 : 0), U1 = T1 + ((J == 1) ? 4 : 0), U2 = (D == 3) ? T2 + ((J == 2) ? 4 : 0) : 1;
  static constexpr int UN = U0 * U1 * U2;
  static constexpr int ULO0 = (J == 0) ? 2 : 0, ULO1 = (J == 1) ? 2 : 0, ULO2 = (D == 3 && J == 2) ? 2 : 0;
  static constexpr int FLO2 = (D == 3) ? 1 : 0;
  static IFADV_DI int fidx(int lx, int ly, int lz) { return (lx + 1) + F0 * ((ly + 1) + F1 * (lz + FLO2)); }
  static IFADV_DI int uidx(int lx, int ly, int lz) { return (lx + ULO0) + U0 * ((ly + ULO1) + U1 * (lz + ULO2)); }
  template <class T> static constexpr size_t smem_bytes(bool mom) { return sizeof(T) * (2 * (size_t)FN + (mom ? (size_t)D * UN : 0)); }
};

// one tile (bx, by, bz) of the sweep; the kernel below runs one tile per CTA, vof2d_step_kernel loops over tiles
template <class T, int D, int J, int TX, int TY, int TZ, bool MOM, int NT>
IFADV_DI void sweep_tile(const SweepP<T>& P, const int bx, const int by, const int bz, unsigned char* smem_raw) {
  using TL = Tile<D, J, TX, TY, TZ>;
  T* sFF = reinterpret_cast<T*>(smem_raw);  // volume flux fᶠ through lower J-faces
  T* sM = sFF + TL::FN;                     // mass flux: ρuf/δt (MOM) or ρuf (pure VOF)
  T* sU = sM + TL::FN;                      // u★ components (MOM)

  const Geo g = P.g;
  const int tid = threadIdx.x;
  const int o0 = 2 + bx * TL::T0, o1 = 2 + by * TL::T1, o2 = (D == 3) ? 2 + bz * TL::T2 : 1;
  const bool per0 = g.per & 1u, per1 = g.per & 2u, per2 = (D == 3) && (g.per & 4u);
  const bool perJ = (J == 0) ? per0 : ((J == 1) ? per1 : per2);
  const int nJ = g.n[J];
  const FMap<T, D> F{P.f_in, g};

  // ---- stage 1: VOF face fluxes of the tile (+ halo) ---------------------------------------------------
  for (int idx = tid; idx < TL::FN; idx += NT) {
    const int a0 = idx % TL::F0, a1 = (idx / TL::F0) % TL::F1, a2 = idx / (TL::F0 * TL::F1);
    const int l0 = a0 - 1, l1 = a1 - 1, l2 = a2 - TL::FLO2;
    int v0 = o0 + l0, v1 = o1 + l1, v2 = (D == 3) ? o2 + l2 : 1;
    // entries never read: two halo coordinates at once
    const int nhalo = (l0 < 0) + (l1 < 0) + ((D == 3) && (l2 < 0));
    T ff = T(0), m = T(0);
    bool need = nhalo <= 1;
    // tile overhang across, faces beyond the upper boundary face along J
    if (J != 0 && v0 > g.n[0] - 1) need = false;
    if (J != 1 && v1 > g.n[1] - 1) need = false;
    if (D == 3 && J != 2 && v2 > g.n[2] - 1) need = false;
    const int p = (J == 0) ? v0 : ((J == 1) ? v1 : v2);
    if (p > nJ) need = false;
    if (!perJ && p < 2) need = false;  // face 1 only exists through the velocity BC on ρuf
    if (need) {
      // map to the interior-equivalent face
      if (J != 0) v0 = mapc(v0, g.n[0], per0); else if (perJ) v0 = wrapc(v0, g.n[0]);
      if (J != 1) v1 = mapc(v1, g.n[1], per1); else if (perJ) v1 = wrapc(v1, g.n[1]);
      if (D == 3) { if (J != 2) v2 = mapc(v2, g.n[2], per2); else if (perJ) v2 = wrapc(v2, g.n[2]); }
      const long long li = lin3(g, v0, v1, v2);
      const T dl = P.hdt * (__ldg(P.uj + li) + __ldg(P.u0j + li));  // δt/2*(u+u⁰), advection.jl:110
      if (dl != T(0)) {                                             // advection.jl:115
        // upwind cell (advection.jl:120)
        int c0 = v0, c1 = v1, c2 = v2;
        if (dl > T(0)) { if (J == 0) c0 -= 1; else if (J == 1) c1 -= 1; else c2 -= 1; }
        const int cJ = (J == 0) ? c0 : ((J == 1) ? c1 : c2);
        // a ghost upwind cell on a non-periodic side has no reconstruction (its n̂ is never set: VOFutil.jl:50-53)
        const bool ghost = !perJ && (cJ < 2 || cJ > nJ - 1);
        const T fc = F(c0, c1, c2);
        if (ghost || fullorempty(fc)) ff = fc * dl;  // advection.jl:125-126
        else {
          FBox<T, D> B{F, c0, c1, c2};
          ff = plic_face_flux<T, D>(P.scheme, B, fc, J, dl);  // advection.jl:131-134
        }
        m = dl * P.lr + P.omlr * ff;  // fᶠ2ρuf, VOFutil.jl:218
        if (MOM) m = m * P.idt;       // rmul!(ρuf, inv(δt)), flow.jl:207
      }
    }
    sFF[idx] = ff;
    sM[idx] = m;
  }

  // ---- stage 1b: u★ = BC!(ρu/ρ(f̄)) along the sweep line (flow.jl:197, VOFutil.jl:198-201) -----------------
  if (MOM) {
    for (int idx = tid; idx < D * TL::UN; idx += NT) {
      const int i = idx / TL::UN, r = idx - i * TL::UN;
      const int a0 = r % TL::U0, a1 = (r / TL::U0) % TL::U1, a2 = r / (TL::U0 * TL::U1);
      int v0 = o0 + a0 - TL::ULO0, v1 = o1 + a1 - TL::ULO1, v2 = (D == 3) ? o2 + a2 - TL::ULO2 : 1;
      bool need = true;
      if (J != 0 && v0 > g.n[0] - 1) need = false;
      if (J != 1 && v1 > g.n[1] - 1) need = false;
      if (D == 3 && J != 2 && v2 > g.n[2] - 1) need = false;
      const int p = (J == 0) ? v0 : ((J == 1) ? v1 : v2);
      if (!perJ && (p < 1 || p > nJ)) need = false;
      T val = T(0);
      if (need) {
        const int vi = (i == 0) ? v0 : ((i == 1) ? v1 : v2);
        const bool peri = (i == 0) ? per0 : ((i == 1) ? per1 : per2);
        if (!peri && i == 0 && vi == g.n[0] && P.uexit != nullptr)  // exitBC: BC! with saveexit keeps plane N of component x
          val = __ldg(P.uexit + lin3(g, v0, mapc(v1, g.n[1], per1), (D == 3) ? mapc(v2, g.n[2], per2) : 1));
        else if (!peri && (vi == 1 || vi == 2 || vi == g.n[i])) val = P.A[i];  // Dirichlet planes of BC!
        else {
          v0 = (i == 0) ? (peri ? wrapc(v0, g.n[0]) : v0) : mapc(v0, g.n[0], per0);
          v1 = (i == 1) ? (peri ? wrapc(v1, g.n[1]) : v1) : mapc(v1, g.n[1], per1);
          if (D == 3) v2 = (i == 2) ? (peri ? wrapc(v2, g.n[2]) : v2) : mapc(v2, g.n[2], per2);
          const T fb = (F(v0, v1, v2) + F(v0 - (i == 0), v1 - (i == 1), v2 - (i == 2))) / T(2);
          val = __ldg(P.rhou_in + (long long)i * g.S + lin3(g, v0, v1, v2)) / lin_interp(fb, P.lr, P.omlr);
        }
      }
      sU[idx] = val;
    }
  }
  __syncthreads();

  // ---- stage 2: cell update ------------------------------------------------------------------------------
  double rmax = -INFINITY, rmin = INFINITY;
  unsigned int amax = 0, amin = 0;
  int rnan = 0;
  for (int idx = tid; idx < TL::CELLS; idx += NT) {
    const int l0 = idx % TL::T0, l1 = (idx / TL::T0) % TL::T1, l2 = idx / (TL::T0 * TL::T1);
    const int k0 = o0 + l0, k1 = o1 + l1, k2 = (D == 3) ? o2 + l2 : 1;
    if (k0 > g.n[0] - 1 || k1 > g.n[1] - 1 || (D == 3 && k2 > g.n[2] - 1)) continue;
    const long long lk = lin3(g, k0, k1, k2);
    const long long sJ = (J == 0) ? 1 : ((J == 1) ? g.s1 : g.s2);
    const int e0 = (J == 0), e1 = (J == 1), e2 = (J == 2);
    const T fK = __ldg(P.f_in + lk);
    int cb;
    if (P.first) { cb = (fK < T(0.5)) ? 0 : 1; P.cbar[lk] = (int8_t)cb; }  // flow.jl:172 / advection.jl:40
    else cb = P.cbar[lk];
    const T div = (__ldg(P.uj + lk + sJ) - __ldg(P.uj + lk)) + (__ldg(P.u0j + lk + sJ) - __ldg(P.u0j + lk));  // ∂(d,I,u)+∂(d,I,u⁰)
    const int fi = TL::fidx(l0, l1, l2), fiu = TL::fidx(l0 + e0, l1 + e1, l2 + e2);
    // f[I] += fᶠ[I]-fᶠ[I+δ] + c̄[I]*(∂u+∂u⁰)*δt/2          advection.jl:83
    T fn = fK + ((sFF[fi] - sFF[fiu]) + ((T(cb) * div) * P.dt) / T(2));
    {  // reportFillError's extrema (advection.jl:146-148), before cleanWisp!
      const double fd = (double)fn;
      if (fn != fn) rnan = 1;
      if (fd > rmax) { rmax = fd; amax = (unsigned int)lk; }
      if (fd < rmin) { rmin = fd; amin = (unsigned int)lk; }
    }
    fn = (fn < P.tol) ? T(0) : ((fn > P.onemtol) ? T(1) : fn);  // cleanWisp!, VOFutil.jl:127-136
    P.f_out[lk] = fn;
    if (!MOM && P.rhouf_j != nullptr) {
      P.rhouf_j[lk] = sM[fi];
      const int kJ = (J == 0) ? k0 : ((J == 1) ? k1 : k2);
      if (kJ == nJ - 1) P.rhouf_j[lk + sJ] = sM[fiu];  // inside_uWB includes the upper boundary face
    }
    if (MOM) {
      const int kJ = (J == 0) ? k0 : ((J == 1) ? k1 : k2);
      // ρ̄∂ⱼuⱼ at this cell (flow.jl:216)
      const T dilK = (lin_interp(T(cb), P.lr, P.omlr) * div) / T(2);
      // BC-aware mass flux (velocity BC! on ρuf, flow.jl:207): Dirichlet planes of component j
      auto Mat = [&](int m0, int m1, int m2) -> T {
        const int pp = ((J == 0) ? o0 + m0 : ((J == 1) ? o1 + m1 : o2 + m2));
        if (!perJ && (pp == 1 || pp == 2 || (pp == nJ && !(J == 0 && P.uexit != nullptr)))) return P.A[J];  // exitBC keeps the computed flux of face N
        return sM[TL::fidx(m0, m1, m2)];
      };
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const int d0 = (i == 0), d1 = (i == 1), d2 = (i == 2);
        const bool peri = (i == 0) ? per0 : ((i == 1) ? per1 : per2);
        // face-centred old f of a (virtual) momentum cell: dρ after f2face!+BCv! (flow.jl:205, VOFutil.jl:91-104)
        auto drho_at = [&](int c0, int c1, int c2) -> T {
          if (i == J && !perJ) {
            const int cJ = (J == 0) ? c0 : ((J == 1) ? c1 : c2);
            if (cJ == nJ) return __ldg(P.drho + (long long)i * g.S + lin3(g, c0, c1, c2));  // plane never written by f2face!
            if (cJ == 1) { if (J == 0) c0 = 3; else if (J == 1) c1 = 3; else c2 = 3; }   // BCv!: f[I] = f[I+2δ]
          }
          return (F(c0, c1, c2) + F(c0 - d0, c1 - d1, c2 - d2)) / T(2);
        };
        // momentum flux through the lower J-face of the i-momentum cell at tile-local (m0,m1,m2)
        auto flux = [&](int m0, int m1, int m2) -> T {
          const int q0 = o0 + m0, q1 = o1 + m1, q2 = (D == 3) ? o2 + m2 : 1;
          const int pp = (J == 0) ? q0 : ((J == 1) ? q1 : q2);
          const T Psi = (Mat(m0, m1, m2) + Mat(m0 - d0, m1 - d1, m2 - d2)) / T(2);  // ϕ(i,CI(I,j),ρuf)
          const T* su = sU + i * TL::UN;
          const int ui = TL::uidx(m0, m1, m2);
          const int us = (J == 0) ? 1 : ((J == 1) ? TL::U0 : TL::U0 * TL::U1);
          const T um1 = su[ui - us], uc = su[ui];
          T uu, cc, dd;
          if (!perJ && pp == 2) {  // ϕuL, flow.jl:28-31
            if (Psi > T(0)) { uu = T(2) * um1 - uc; cc = um1; dd = uc; }
            else { uu = su[ui + us]; cc = uc; dd = um1; }
          } else if (!perJ && pp == nJ) {  // ϕuR, flow.jl:32-35
            if (Psi < T(0)) { uu = T(2) * uc - um1; cc = uc; dd = um1; }
            else { uu = su[ui - 2 * us]; cc = um1; dd = uc; }
          } else {  // ϕu, flow.jl:20-23 (ϕuP on a periodic boundary is the same stencil through the wrap)
            if (Psi > T(0)) { uu = su[ui - 2 * us]; cc = um1; dd = uc; }
            else { uu = su[ui + us]; cc = uc; dd = um1; }
          }
          // donor momentum cell (flow.jl:42) and its old mass (flow.jl:50)
          const int sh = (Psi > T(0)) ? 1 : 0;
          const T mOld = lin_interp(drho_at(q0 - sh * e0, q1 - sh * e1, q2 - sh * e2), P.lr, P.omlr);
          return syndrom_flux(P.lim, Psi, uu, cc, dd, mOld, P.dt);
        };
        const T Flo = flux(l0, l1, l2);
        const T Fhi = flux(l0 + e0, l1 + e1, l2 + e2);
        // ρ̄∂ⱼuⱼ at I-δi after BCf! (flow.jl:217): evaluated at the interior-equivalent cell
        T dilM;
        {
          const int ki = (i == 0) ? k0 : ((i == 1) ? k1 : k2);
          if (!peri && ki == 2) dilM = dilK;  // Neumann ghost copies the first interior cell
          else {
            int c0 = k0 - d0, c1 = k1 - d1, c2 = k2 - d2;
            if (peri) { if (i == 0) c0 = wrapc(c0, g.n[0]); else if (i == 1) c1 = wrapc(c1, g.n[1]); else c2 = wrapc(c2, g.n[2]); }
            const long long lc = lin3(g, c0, c1, c2);
            const int cbm = P.first ? ((__ldg(P.f_in + lc) < T(0.5)) ? 0 : 1) : (int)P.cbar[lc];
            const T divm = (__ldg(P.uj + lc + sJ) - __ldg(P.uj + lc)) + (__ldg(P.u0j + lc + sJ) - __ldg(P.u0j + lc));
            dilM = (lin_interp(T(cbm), P.lr, P.omlr) * divm) / T(2);
          }
        }
        const long long lki = (long long)i * g.S + lk;
        // r = Φ[I] - Φ[I+δj] + uOld*ϕ(i,I,ρ̄∂ⱼuⱼ);  ρu += δt*r          flow.jl:223-231
        const T r = (Flo - Fhi) + __ldg(P.uOld + lki) * ((dilK + dilM) / T(2));
        P.rhou_out[lki] = __ldg(P.rhou_in + lki) + P.dt * r;
      }
      (void)kJ;
    }
  }

  // ---- fill-error reduction: warp shuffles, then one atomic per CTA (replaces findmax/findmin + host sync) --
  if (P.red != nullptr) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double omax = __shfl_xor_sync(0xffffffffu, rmax, off), omin = __shfl_xor_sync(0xffffffffu, rmin, off);
      const unsigned int oamax = __shfl_xor_sync(0xffffffffu, amax, off), oamin = __shfl_xor_sync(0xffffffffu, amin, off);
      const int onan = __shfl_xor_sync(0xffffffffu, rnan, off);
      if (omax > rmax) { rmax = omax; amax = oamax; }
      if (omin < rmin) { rmin = omin; amin = oamin; }
      rnan |= onan;
    }
    if ((tid & 31) == 0) {
      if (rmax > -INFINITY) {
        atomicMax(P.red + 0, ord_key(rmax));
        atomicMax(P.red + 2, ((unsigned long long)ord_key32((float)rmax) << 32) | amax);
      }
      if (rmin < INFINITY) {
        atomicMin(P.red + 1, ord_key(rmin));
        atomicMin(P.red + 3, ((unsigned long long)ord_key32((float)rmin) << 32) | amin);
      }
      if (rnan) atomicAdd(P.red + 4, 1ull);
    }
  }
}

template <class T, int D, int J, int TX, int TY, int TZ, bool MOM, int NT>
__global__ void __launch_bounds__(NT) sweep_kernel(const SweepP<T> P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  sweep_tile<T, D, J, TX, TY, TZ, MOM, NT>(P, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, smem_raw);
}

}  // namespace ifadv
