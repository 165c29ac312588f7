// ifadv_xsweep.cuh -- fused CMOM directional sweep along x (J = 0) for 3-D grids: LEAN plane marching.
//
// x is the contiguous dimension, so the sweep direction runs ACROSS the lanes of a warp: a CTA owns a tile of 32 (x) x TY (y)
// columns, marches along z, and a thread owns CPT cells of every plane (rows ty, ty+TR, ..).  What the stencil reaches along
// x or y comes from the neighbour threads through shared memory; what it reaches along z (previous mass flux, dilation, f)
// stays in the owner's registers or in 4-slot rings.  The work of a plane is split into three stages that run skewed by
// one plane each, so that ONE barrier per step separates every producer from its consumers:
//     step k:   S1(k+2)  u★ = BC!(ρu/ρ(f̄)) (flow.jl:197), VOF face flux + mass flux (advection.jl:108-137), dilation (flow.jl:216)
//               S2(k+1)  SynDRoM momentum flux through the lower x-face of every cell (flow.jl:20-57,223)
//               S3(k)    update of f (advection.jl:83, cleanWisp!) and ρu (flow.jl:224-231), fill-error extrema
// Interface faces marked in S1 are reconstructed lane-dense at the top of the next step (second barrier only then).
// The steps are unrolled by four, so every ring slot is a compile-time constant; tiles that touch a non-periodic x
// boundary run the XB instantiation of the same body (Dirichlet planes of BC!, ϕuL/ϕuR, ghost upwind cells), all other
// tiles a body with those rules folded away.  The halo of the tile (u★ at x-2,x-1,x+32,x+33; mass flux / dilation at x-1
// and y-1; the face x+32) is computed by the first three warps from the same shared planes (role bits per halo entry).
// Arithmetic (expression by expression) and boundary rules as in ifadv_march.cuh<J=0>; reference lines cited there.
#pragma once
#include "ifadv_along2.cuh"

namespace ifadv {

template <int TY> struct XTile {
  static constexpr int WX = 37, WY = TY + 3, PL = WX * WY, NC = 32 * TY, NH = PL - NC;
  // planes: f x4, u_x x2, u⁰_x x2, ρu 2x3, u★ 2x3, M x4, Φ/fᶠ 2x4, ρ̄∂u x4, fᶠ of reconstructed faces x1
  static constexpr int NPL = 37;
  template <class T> static constexpr size_t smem_bytes() { return sizeof(T) * (size_t)PL * (NPL + 1) + sizeof(int) * (PL + 8); }
};

enum : unsigned {
  XF_NEEDM = 1u << 0,  // the face carries a VOF flux                     va <= nA && (perA || va >= 2)
  XF_DIRA = 1u << 1,   // Dirichlet plane of component x (BC!)            va in {1,2,nA}, x not periodic
  XF_DIRAM = 1u << 2,  // ... at va-1
  XF_DILSH = 1u << 3,  // Neumann ghost of ρ̄∂ⱼuⱼ along x: use the cell at va+1 (va == 1)
  XF_LVAR = 1u << 4,   // ϕuL face (va == 2)
  XF_RVAR = 1u << 5,   // ϕuR face (va == nA)
  XF_GHLO = 1u << 6,   // cell va-1 is a ghost cell on a non-periodic side (no PLIC reconstruction)
  XF_GHHI = 1u << 7,   // cell va itself is one
  // roles of a halo entry
  XH_US = 1u << 8,     // u★ (columns -2,-1,32,33)
  XH_M = 1u << 9,      // VOF / mass flux (columns -1, 32; row -1)
  XH_DIL = 1u << 10,   // dilation (column -1; row -1)
  XH_FL = 1u << 11,    // SynDRoM flux of the face x+32
  XH_DIRB = 1u << 12,  // Dirichlet plane of component y at this row
  XF_EXIT = 1u << 13,  // exitBC: plane nA of component x keeps its saved value (BC! with saveexit) and the face nA its computed mass flux
};

IFADV_DI unsigned xflags(int va, int nA, bool perA, bool exitbc = false) {
  unsigned f = 0;
  if (va <= nA && (perA || va >= 2)) f |= XF_NEEDM;
  if (!perA) {
    if (va == 1 || va == 2 || (va == nA && !exitbc)) f |= XF_DIRA;
    if (va == nA && exitbc) f |= XF_EXIT;
    if (va - 1 == 1 || va - 1 == 2 || va - 1 == nA) f |= XF_DIRAM;
    if (va == 1) f |= XF_DILSH;
    if (va == 2) f |= XF_LVAR;
    if (va == nA) f |= XF_RVAR;
    if (va - 1 < 2 || va - 1 > nA - 1) f |= XF_GHLO;
    if (va < 2 || va > nA - 1) f |= XF_GHHI;
  }
  return f;
}

// 3^3 box accessor on the 4-slot f ring (x fastest, rows along y, ring along z)
template <class T, int PL, int WX> struct XBox {
  const T* sF;
  int e, rel;  // rel = (plane of the box centre) - ks
  IFADV_DI T operator()(int dx, int dy, int dz) const { return sF[((rel + dz) & 3) * PL + e + dx + dy * WX]; }
};

// SAMEU: u¹ and u² are one array (see ifadv_along2.cuh): no u⁰ copy stream, the u⁰ values are the u values
template <class T, int CPT, bool FUSED, bool KOREN, bool XB, int NT, bool SAMEU>
IFADV_DI void xsweep_body(const SweepP<T>& P, const int chunk, T* sm) {
  constexpr int TR = NT / 32, TY = TR * CPT;
  using TL = XTile<TY>;
  constexpr int WX = TL::WX, PL = TL::PL;
  static_assert(TL::NH <= NT, "one halo entry per thread");
  constexpr int OF = 0, OU = OF + 4 * PL, OU0 = OU + 2 * PL, OR = OU0 + 2 * PL, OUS = OR + 6 * PL, OM = OUS + 6 * PL, OFL = OM + 4 * PL,
                ODIL = OFL + 8 * PL, OFX = ODIL + 4 * PL, ODL = OFX + PL, OEND = ODL + PL;
  int* sList = reinterpret_cast<int*>(sm + OEND);
  int* sCnt = sList + PL;  // 4 counters: marks of step k go to counter (k+1-ks)&3

  const Geo& g = P.g;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int nA = g.n[0], nB = g.n[1], nC = g.n[2];
  const unsigned s1 = (unsigned)g.s1, s2 = (unsigned)g.s2;
  const bool perA = g.per & 1u, perB = (g.per >> 1) & 1u, perC = (g.per >> 2) & 1u;
  const unsigned cB = (unsigned)P.coff[1], cC = (unsigned)P.coff[2];
  const int ox = 2 + blockIdx.x * 32, oy = 2 + blockIdx.y * TY;
  const int k0 = P.kz0 + blockIdx.z * chunk, k1 = min(k0 + chunk, P.kz1);
  const int ks = k1 - 4 * ((k1 - k0 + 3 + 3) / 4);  // >= 3 warm-up steps, a multiple of four steps in total
  const T lr = P.lr, omlr = P.omlr, dt = P.dt;
  const T lam1 = lin_interp(T(1), lr, omlr);
  const T AA = P.A[0], AB = P.A[1], AC = P.A[2];
  const bool first = FUSED ? true : (P.first != 0);  // the fused sweep is always sweep 1
  const T* const rsrc = FUSED ? P.uOld : P.rhou_in;  // fused sweep 1: the ρu ring carries uOld, ρu = BC!(uOld*ρ(f̄)) on the fly
  if (tid < 4) sCnt[tid] = 0;

  // ---- per-thread constants: own cells (one column, CPT rows) and (for the first NH threads) one halo entry --------------------
  const int vx = ox + tx;
  const unsigned cflg = XB ? xflags(vx, nA, perA, P.uexit != nullptr) : XF_NEEDM;
  const int e0 = (tx + 3) + WX * (ty + 2);  // shared entry of cell 0; cell j adds j*TR*WX
  // x as stored minus x mapped: non-zero only in a ghost column (the thread at va = nA evaluates the boundary face from u_x[nA])
  const unsigned gox = XB ? (unsigned)((perA ? wrapc(vx, nA) : min(max(vx, 1), nA)) - mapc(vx, nA, perA)) : 0u;
  unsigned gm[CPT];
  bool valid[CPT], dirB[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int vy = oy + ty + j * TR;
    gm[j] = (unsigned)(mapc(vx, nA, perA) - 1) + (unsigned)(mapc(vy, nB, perB) - 1) * s1;
    valid[j] = vx <= nA - 1 && vy <= nB - 1;
    dirB[j] = !perB && (vy == 2 || vy == nB);
  }
  int eh = 0;
  unsigned ghm = 0, gho = 0, hflg = 0;
  const bool hasH = tid < TL::NH;
  if (hasH) {
    int lx, ly;
    const int h = tid;
    if (h < TY) { lx = 32; ly = h; hflg = XH_US | XH_M | XH_FL; }                    // face x+32
    else if (h < 2 * TY) { lx = -1; ly = h - TY; hflg = XH_US | XH_M | XH_DIL; }     // column x-1
    else if (h < 3 * TY) { lx = -2; ly = h - 2 * TY; hflg = XH_US; }
    else if (h < 4 * TY) { lx = 33; ly = h - 3 * TY; hflg = XH_US; }
    else if (h < 4 * TY + 33) { lx = h - 4 * TY; ly = -1; hflg = XH_M | (lx < 32 ? XH_DIL : 0u); }  // row y-1
    else if (h < 5 * TY + 33) { lx = -3; ly = h - (4 * TY + 33); }                   // from here on: only f (PLIC boxes, ρ(f̄) of the halo)
    else if (h < 5 * TY + 33 + WX) { lx = h - (5 * TY + 33) - 3; ly = -2; }
    else if (h < 5 * TY + 33 + 2 * WX) { lx = h - (5 * TY + 33 + WX) - 3; ly = TY; }
    else { const int r = h - (5 * TY + 33 + 2 * WX); ly = -1; lx = (r < 3) ? r - 3 : 33; }
    eh = (lx + 3) + WX * (ly + 2);
    const int va = ox + lx, vb = oy + ly;
    const unsigned yo = (unsigned)(mapc(vb, nB, perB) - 1) * s1;
    ghm = (unsigned)(mapc(va, nA, perA) - 1) + yo;
    gho = (unsigned)((perA ? wrapc(va, nA) : min(max(va, 1), nA)) - 1) + yo;  // x as stored (component x, u_x faces)
    if (XB) hflg |= xflags(va, nA, perA, P.uexit != nullptr); else hflg |= XF_NEEDM;
    if (!perB && (vb == 2 || vb == nB)) hflg |= XH_DIRB;
  }
  const bool hUS = (hflg & XH_US) != 0, hM = (hflg & XH_M) != 0, hDIL = (hflg & XH_DIL) != 0, hFL = (hflg & XH_FL) != 0;

  auto pm = [&](int v) -> unsigned { return (unsigned)(map1(v, nC, perC) - 1) * s2; };
  auto po = [&](int v) -> unsigned { return (unsigned)(own1(v, nC, perC) - 1) * s2; };
  auto dirCf = [&](int v) -> bool { return !perC && (v == 2 || v == nC); };
  constexpr unsigned SZ = sizeof(T);
  const unsigned sb = (unsigned)__cvta_generic_to_shared(sm);
  const unsigned se0 = sb + (unsigned)e0 * SZ, seh = sb + (unsigned)eh * SZ;
  constexpr unsigned JS = TR * WX * SZ;  // byte stride between the cells of a thread

  // copies of plane v into ring slots q4 (f) / q2 (u, ρu)
  auto issue = [&](const unsigned pmv, const unsigned pov, unsigned q4, unsigned q2) {
    const unsigned dF = (OF + q4 * PL) * SZ, dU = (OU + q2 * PL) * SZ, dU0 = (OU0 + q2 * PL) * SZ, dR = (OR + q2 * 3 * PL) * SZ;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const unsigned gmv = gm[j] + pmv;
      cp_async_s(se0 + dF + j * JS, P.f_in + gmv);
      cp_async_s(se0 + dU + j * JS, P.u + (gmv + gox));
      if (!SAMEU) cp_async_s(se0 + dU0 + j * JS, P.u0 + (gmv + gox));
      cp_async_s(se0 + dR + j * JS, ((XB && (cflg & XF_EXIT)) ? P.uexit : rsrc) + (gmv + gox));  // exitBC: saved exit value of u★
      cp_async_s(se0 + dR + PL * SZ + j * JS, rsrc + (gmv + cB));
      cp_async_s(se0 + dR + 2 * PL * SZ + j * JS, rsrc + (gm[j] + cC + pov));
    }
    if (hasH) {
      cp_async_s(seh + dF, P.f_in + (ghm + pmv));
      if (hM) {
        cp_async_s(seh + dU, P.u + (gho + pmv));
        if (!SAMEU) cp_async_s(seh + dU0, P.u0 + (gho + pmv));
      }
      if (hUS) {
        cp_async_s(seh + dR, ((XB && (hflg & XF_EXIT)) ? P.uexit : rsrc) + (gho + pmv));
        cp_async_s(seh + dR + PL * SZ, rsrc + (ghm + cB + pmv));
        cp_async_s(seh + dR + 2 * PL * SZ, rsrc + (ghm + cC + pov));
      }
    }
  };

  // ---- rolling register state per own cell (values entering step k) ----------------------------------------------------------------
  T us1[3][CPT];            // own u★ of plane k+1
  T h1[3][CPT], h0[3][CPT]; // ρ at the lower x / y / z faces of the own cell in planes k+1, k (h0 only in the fused sweep)
  T M1[CPT];                // own mass flux (lower x-face) of plane k+1 (plane k: shared M ring)
  T FF1[CPT];               // own fᶠ of plane k+1 (plane k: the Φ/fᶠ buffer written in S2)
  T dv1[CPT], dv0[CPT];     // c̄(∂u+∂u⁰)δt/2 of planes k+1, k
  T Fl0[3][CPT];            // own SynDRoM fluxes of plane k
  T f1[CPT];                // own f(k+1) (f(k): shared f ring)
  bool mk1[CPT];
  T hFF1 = T(0);            // halo face x+32: fᶠ of plane k+1
  bool hmk1 = false;
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
#pragma unroll
    for (int r = 0; r < 3; ++r) { us1[r][j] = T(0); h1[r][j] = h0[r][j] = T(1); Fl0[r][j] = T(0); }
    M1[j] = FF1[j] = dv1[j] = dv0[j] = T(0);
    f1[j] = T(0);
    mk1[j] = false;
  }
  T rmax = -INFINITY, rmin = INFINITY;
  unsigned int amax = 0, amin = 0;

  // ---- prologue ------------------------------------------------------------------------------------------------------------------
  issue(pm(ks), po(ks), 0, 0); issue(pm(ks + 1), po(ks + 1), 1, 1);  // f(ks), f(ks+1) feed ρ(f̄) / the PLIC boxes of the first steps; u, ρu of these planes are unused
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  issue(pm(ks + 2), po(ks + 2), 2, 0);
  cp_async_commit();
  for (int i = tid; i < 4 * PL; i += NT) { sm[OM + i] = T(0); sm[ODIL + i] = T(0); }
  for (int i = tid; i < 6 * PL; i += NT) sm[OUS + i] = T(0);
  for (int i = tid; i < 8 * PL; i += NT) sm[OFL + i] = T(0);
#pragma unroll
  for (int j = 0; j < CPT; ++j) f1[j] = sm[OF + PL + e0 + j * TR * WX];

  unsigned lkU = (unsigned)(ks - 1) * s2;  // (k-1)*s2, offset of plane k
  unsigned pm3 = pm(ks + 3), po3 = po(ks + 3);  // offsets of plane k+3 (mapped / as stored), rolled with the march
  unsigned pm2 = pm(ks + 2);                    // mapped offset of plane k+2 (c̄)

  // SynDRoM fluxes through the lower x-face of the cell at entry e, plane k+1 (ring phase I): shared by own cells and the halo face x+32
  auto face_flux = [&](auto ic, const int k, const int e, const unsigned flg, const unsigned yoff_drho, const T Mown, const T Mprev,
                       const T* uc, const T* hown, T* out) {
    constexpr int I = decltype(ic)::value;
    const T* US = sm + OUS + ((I + 1) & 1) * 3 * PL;
    const T* Mp = sm + OM + ((I + 1) & 3) * PL;
    const T* F1 = sm + OF + ((I + 1) & 3) * PL;  // f(k+1)
    const T* F0 = sm + OF + (I & 3) * PL;        // f(k)
    const bool dira = XB && (flg & XF_DIRA), Lvar = XB && (flg & XF_LVAR), Rvar = XB && (flg & XF_RVAR);
    const T Mc = dira ? AA : Mown;  // velocity BC! on ρuf: Dirichlet planes of component x (flow.jl:207)
    // density of the donor cell x-1: the same ρ(f̄) its u★ was formed with
    const T fm = F1[e - 1];
    T hm[3];
    hm[0] = rho_face(fm, F1[e - 2], lr, omlr);
    hm[1] = rho_face(fm, F1[e - 1 - WX], lr, omlr);
    hm[2] = rho_face(fm, F0[e - 1], lr, omlr);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      T Mo;
      if (r == 0) Mo = (XB && (flg & XF_DIRAM)) ? AA : Mp[e - 1];
      else if (r == 1) Mo = dira ? AA : Mp[e - WX];
      else Mo = dira ? AA : Mprev;
      const T Psi = (Mc + Mo) / T(2);
      const T* su = US + r * PL;
      const T um1 = su[e - 1], ucr = uc[r];
      const bool pos = Psi > T(0);
      T uu, cc, dd;
      if (Lvar) {  // ϕuL, flow.jl:28-31
        if (pos) { uu = T(2) * um1 - ucr; cc = um1; dd = ucr; }
        else { uu = su[e + 1]; cc = ucr; dd = um1; }
      } else if (Rvar) {  // ϕuR, flow.jl:32-35
        if (Psi < T(0)) { uu = T(2) * ucr - um1; cc = ucr; dd = um1; }
        else { uu = su[e - 2]; cc = um1; dd = ucr; }
      } else {  // ϕu, flow.jl:20-23
        uu = pos ? su[e - 2] : su[e + 1];
        cc = pos ? um1 : ucr;
        dd = pos ? ucr : um1;
      }
      T mOld = pos ? hm[r] : hown[r];
      if (XB && r == 0) {
        if (Lvar && pos) mOld = rho_face(F1[e + 1], F1[e], lr, omlr);  // donor index 1: BCv! copies plane 3 = (f(3)+f(2))/2
        if (Rvar && !pos)                                                // donor index nA: the plane f2face! never writes
          mOld = lin_interp(__ldg(P.drho + ((unsigned)(nA - 1) + yoff_drho + (unsigned)(mapc(k + 1, nC, perC) - 1) * s2)), lr, omlr);
      }
      out[r] = syndrom_flux_t<KOREN>(P.lim, Psi, uu, cc, dd, mOld, dt);
    }
  };

  auto step = [&](auto ic, const int k) {
    constexpr int I = decltype(ic)::value;  // (k-ks)&3
    const int rel = k - ks;
    // c̄(k+2) for this step's dilation: the read is issued before the barrier, so its latency overlaps the wait
    int cb2[CPT];
    int cbh2 = 0;
#pragma unroll
    for (int j = 0; j < CPT; ++j) cb2[j] = 0;
    if (!first) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) cb2[j] = (int)P.cbar[gm[j] + pm2];
      if (hDIL) cbh2 = (int)P.cbar[ghm + pm2];
    }
    cp_async_wait_all();
    __syncthreads();  // the copies issued during step k-1 have landed; every read / write of step k-1 is done

    const bool store = k >= k0;
    const bool dirC2 = dirCf(k + 2), dirC0 = dirCf(k);
    // A. copies of plane k+3
    issue(pm3, po3, (I + 3) & 3, (I + 3) & 1);
    cp_async_commit();
    {  // offsets of plane k+4: one stride further unless the plane is within the boundary band (block-uniform)
      const int v = k + 4;
      pm2 = pm3;
      if (v >= 3 && v <= nC - 1) { pm3 += s2; po3 += s2; }
      else { pm3 = pm(v); po3 = po(v); }
    }

    // P. lane-dense PLIC reconstruction of the interface faces of plane k+1 marked in step k-1 (advection.jl:131-134)
    {
      const int cnt = sCnt[rel & 3];  // block-uniform
      if (cnt > 0) {
        T* Mp = sm + OM + ((rel + 1) & 3) * PL;
        for (int i = tid; i < cnt; i += NT) {
          const int e = sList[i];
          const T dl = sm[ODL + i];
          const int eu = (dl > T(0)) ? e - 1 : e;  // upwind cell
          XBox<T, PL, WX> B{sm + OF, eu, rel + 1};
          const T ff = plic_face_flux_inl<T, 3>(P.scheme, B, sm[OF + ((rel + 1) & 3) * PL + eu], 0, dl);
          sm[OFX + e] = ff;
          Mp[e] = (dl * lr + omlr * ff) * P.idt;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          if (mk1[j]) { FF1[j] = sm[OFX + e0 + j * TR * WX]; M1[j] = Mp[e0 + j * TR * WX]; }
        if (hmk1) hFF1 = sm[OFX + eh];
      }
      if (tid == 0) sCnt[(rel - 1) & 3] = 0;
    }

    // ring slots of this step
    constexpr int qF2 = OF + ((I + 2) & 3) * PL, qF1 = OF + ((I + 1) & 3) * PL;
    constexpr int qU = OU + ((I + 2) & 1) * PL, qU0 = (SAMEU ? OU : OU0) + ((I + 2) & 1) * PL, qR = OR + ((I + 2) & 1) * 3 * PL;
    constexpr int wUS = OUS + ((I + 2) & 1) * 3 * PL, wM = OM + ((I + 2) & 3) * PL, wD = ODIL + ((I + 2) & 3) * PL;
    constexpr int wFL = OFL + ((I + 1) & 1) * 4 * PL, rFL = OFL + (I & 1) * 4 * PL;
    constexpr int rD0 = ODIL + (I & 3) * PL, rDm = ODIL + ((I + 3) & 3) * PL;
    const int cm = (rel + 1) & 3;  // counter of the marks of this step

    // VOF face flux + mass flux through the lower x-face of the cell at entry e (plane k+2); returns fᶠ, writes M, marks interface faces
    auto vof_face = [&](const int e, const unsigned flg, const T fup, const T fown, T& FFo, T& Mo, bool& mko) {
      FFo = T(0); Mo = T(0); mko = false;
      if (!XB || (flg & XF_NEEDM)) {
        T dl = P.hdt * (sm[qU + e] + sm[qU0 + e]);  // δt/2*(u+u⁰), advection.jl:110
        dl = (dl != T(0)) ? dl : T(0);              // -0 -> +0: the zero-flux case of advection.jl:115 without a branch
        const bool up = dl > T(0);
        const T fc = up ? fup : fown;               // upwind cell x-1 / x, advection.jl:120
        const bool gh = XB && (flg & (up ? XF_GHLO : XF_GHHI));
        if (dl != T(0) && !gh && !fullorempty(fc)) {
          const int i = atomicAdd(&sCnt[cm], 1);    // interface face: reconstructed lane-dense at the top of the next step
          sList[i] = e;
          sm[ODL + i] = dl;
          mko = true;
        } else {
          FFo = fc * dl;                            // advection.jl:125-126
          Mo = (dl * lr + omlr * FFo) * P.idt;      // fᶠ2ρuf (VOFutil.jl:218), rmul!(ρuf, inv(δt)) (flow.jl:207)
        }
      }
      sm[wM + e] = Mo;
    };

    T us2[3][CPT], h2[3][CPT], M2[CPT], FF2[CPT], dv2[CPT], f2[CPT], Fl1[3][CPT];
    bool mk2[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int e = e0 + j * TR * WX;
      // S1(k+2): u★ = BC!(ρu/ρ(f̄)) (flow.jl:197, VOFutil.jl:198-201)
      f2[j] = sm[qF2 + e];
      const T fxm = sm[qF2 + e - 1];
      h2[0][j] = rho_face(f2[j], fxm, lr, omlr);
      h2[1][j] = rho_face(f2[j], sm[qF2 + e - WX], lr, omlr);
      h2[2][j] = rho_face(f2[j], f1[j], lr, omlr);
      const T qa2 = sm[qR + e], qb2 = sm[qR + PL + e], qc2 = sm[qR + 2 * PL + e];
      // fused: ρu = u*ρ (u2ρu!) formed on the fly, then u★ = ρu/ρ, rounding as the two passes would
      const T ra = t_div(FUSED ? qa2 * h2[0][j] : qa2, h2[0][j]);
      const T rb = t_div(FUSED ? qb2 * h2[1][j] : qb2, h2[1][j]);
      const T rc = t_div(FUSED ? qc2 * h2[2][j] : qc2, h2[2][j]);
      us2[0][j] = (XB && (cflg & XF_DIRA)) ? AA : ((XB && (cflg & XF_EXIT)) ? qa2 : ra);  // Dirichlet planes of BC! / saved exit plane
      us2[1][j] = dirB[j] ? AB : rb;
      us2[2][j] = dirC2 ? AC : rc;
#pragma unroll
      for (int r = 0; r < 3; ++r) sm[wUS + r * PL + e] = us2[r][j];
      // S1(k+2): VOF flux, mass flux, dilation
      vof_face(e, cflg, fxm, f2[j], FF2[j], M2[j], mk2[j]);
      const T div = (sm[qU + e + 1] - sm[qU + e]) + (sm[qU0 + e + 1] - sm[qU0 + e]);  // ∂(d,I,u)+∂(d,I,u⁰)
      if (first) cb2[j] = (f2[j] < T(0.5)) ? 0 : 1;                                  // flow.jl:172 (c̄ from the incoming f)
      dv2[j] = ((cb2[j] ? div : T(0)) * dt) / T(2);                                   // c̄[I]*(∂u+∂u⁰)*δt/2 of advection.jl:83
      sm[wD + e] = ((cb2[j] ? lam1 : lr) * div) / T(2);                               // flow.jl:216
    }
    // halo duties of S1(k+2)
    T hFF2 = T(0);
    bool hmk2 = false;
    if (hUS | hM) {
      const T fh = sm[qF2 + eh], fhxm = sm[qF2 + eh - 1];
      if (hUS) {
        const T ha = rho_face(fh, fhxm, lr, omlr), hb = rho_face(fh, sm[qF2 + eh - WX], lr, omlr), hc = rho_face(fh, sm[qF1 + eh], lr, omlr);
        const T a = sm[qR + eh], b = sm[qR + PL + eh], c = sm[qR + 2 * PL + eh];
        const T ra = t_div(FUSED ? a * ha : a, ha), rb = t_div(FUSED ? b * hb : b, hb), rc = t_div(FUSED ? c * hc : c, hc);
        sm[wUS + eh] = (XB && (hflg & XF_DIRA)) ? AA : ((XB && (hflg & XF_EXIT)) ? a : ra);
        sm[wUS + PL + eh] = (hflg & XH_DIRB) ? AB : rb;
        sm[wUS + 2 * PL + eh] = dirC2 ? AC : rc;
      }
      if (hM) {
        T m;
        vof_face(eh, hflg, fhxm, fh, hFF2, m, hmk2);
        if (hDIL) {
          const int e2 = (XB && (hflg & XF_DILSH)) ? eh + 1 : eh;  // BCf! (Neumann) on ρ̄∂ⱼuⱼ along the sweep direction, flow.jl:217
          const T div = (sm[qU + e2 + 1] - sm[qU + e2]) + (sm[qU0 + e2 + 1] - sm[qU0 + e2]);
          const int cb = first ? ((fh < T(0.5)) ? 0 : 1) : cbh2;
          sm[wD + eh] = ((cb ? lam1 : lr) * div) / T(2);
        }
      }
    }

    T q0[3][CPT], uo[3][CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
#pragma unroll
      for (int r = 0; r < 3; ++r) q0[r][j] = uo[r][j] = T(0);
    }
    if (store) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const unsigned o = gm[j] + lkU;
        q0[0][j] = __ldg(rsrc + o);
        q0[1][j] = __ldg(rsrc + (o + cB));
        q0[2][j] = __ldg(rsrc + (o + cC));
        if (!FUSED) {
          uo[0][j] = __ldg(P.uOld + o);
          uo[1][j] = __ldg(P.uOld + (o + cB));
          uo[2][j] = __ldg(P.uOld + (o + cC));
        }
      }
    }
    // S2(k+1): SynDRoM momentum flux through the lower x-face of the own cells and of the halo face x+32
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int e = e0 + j * TR * WX;
      T ucr[3] = {us1[0][j], us1[1][j], us1[2][j]}, hw[3] = {h1[0][j], h1[1][j], h1[2][j]}, fl[3];
      face_flux(ic, k, e, cflg, gm[j] - (unsigned)(mapc(vx, nA, perA) - 1), M1[j], sm[OM + (I & 3) * PL + e], ucr, hw, fl);
#pragma unroll
      for (int r = 0; r < 3; ++r) { Fl1[r][j] = fl[r]; sm[wFL + r * PL + e] = fl[r]; }
      sm[wFL + 3 * PL + e] = FF1[j];
    }
    if (hFL) {
      const T* US = sm + OUS + ((I + 1) & 1) * 3 * PL;
      const T fh = sm[qF1 + eh];
      T ucr[3] = {US[eh], US[PL + eh], US[2 * PL + eh]}, fl[3];
      T hw[3] = {rho_face(fh, sm[qF1 + eh - 1], lr, omlr), rho_face(fh, sm[qF1 + eh - WX], lr, omlr),
                 rho_face(fh, sm[OF + (I & 3) * PL + eh], lr, omlr)};
      face_flux(ic, k, eh, hflg, ghm - (unsigned)(mapc(ox + 32, nA, perA) - 1), sm[OM + ((I + 1) & 3) * PL + eh], sm[OM + (I & 3) * PL + eh],
                ucr, hw, fl);
#pragma unroll
      for (int r = 0; r < 3; ++r) sm[wFL + r * PL + eh] = fl[r];
      sm[wFL + 3 * PL + eh] = hFF1;
    }

    // S3(k): update of the own cells.  ρu (and uOld) of plane k are re-read through L2 into registers (the shared ρu ring only
    // holds the plane u★ is formed from); the reads are in flight while the SynDRoM fluxes above are evaluated.
    const unsigned lk0 = lkU;
    lkU += s2;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int e = e0 + j * TR * WX;
      if (store && valid[j]) {
        const unsigned lk = lk0 + gm[j];  // owned cells are interior: the mapped offset is the cell itself
        const T f0 = sm[OF + (I & 3) * PL + e];
        if (first) P.cbar[lk] = (int8_t)((f0 < T(0.5)) ? 0 : 1);
        T fn = f0 + ((sm[rFL + 3 * PL + e] - sm[rFL + 3 * PL + e + 1]) + dv0[j]);  // advection.jl:83
        rmax = max_nan(rmax, fn);
        rmin = t_min(rmin, fn);
        if (fn > T(1) || fn < T(0)) {  // only cells outside [0,1] can be reported (reportFillError, advection.jl:145-189)
          if (fn >= rmax) amax = lk;
          if (fn <= rmin) amin = lk;
        }
        fn = (fn < P.tol) ? T(0) : ((fn > P.onemtol) ? T(1) : fn);  // cleanWisp!
        P.f_out[lk] = fn;
        const T dK = sm[rD0 + e];
        T qa = q0[0][j], qb = q0[1][j], qc = q0[2][j];
        T oa = uo[0][j], ob = uo[1][j], oc = uo[2][j];
        if (FUSED) {  // u2ρu! + BC!(ρu,uBC): Dirichlet plane 2 of the normal component holds uBC
          oa = qa; ob = qb; oc = qc;
          qa = (XB && (cflg & XF_LVAR)) ? AA : qa * h0[0][j];
          qb = dirB[j] ? AB : qb * h0[1][j];
          qc = dirC0 ? AC : qc * h0[2][j];
        }
        // r = Φ[I] - Φ[I+δj] + uOld*ϕ(i,I,ρ̄∂ⱼuⱼ);  ρu += δt*r          flow.jl:223-231
        const T ra = (Fl0[0][j] - sm[rFL + e + 1]) + oa * ((dK + sm[rD0 + e - 1]) / T(2));
        const T rb = (Fl0[1][j] - sm[rFL + PL + e + 1]) + ob * ((dK + sm[rD0 + e - WX]) / T(2));
        const T rc = (Fl0[2][j] - sm[rFL + 2 * PL + e + 1]) + oc * ((dK + sm[rDm + e]) / T(2));
        P.rhou_out[lk] = qa + dt * ra;
        P.rhou_out[lk + cB] = qb + dt * rb;
        P.rhou_out[lk + cC] = qc + dt * rc;
      }
      // roll the register pipeline
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        Fl0[r][j] = Fl1[r][j]; us1[r][j] = us2[r][j]; h0[r][j] = h1[r][j]; h1[r][j] = h2[r][j];
      }
      FF1[j] = FF2[j]; M1[j] = M2[j]; mk1[j] = mk2[j];
      dv0[j] = dv1[j]; dv1[j] = dv2[j]; f1[j] = f2[j];
    }
    hFF1 = hFF2; hmk1 = hmk2;
  };

  for (int k = ks; k < k1; k += 4) {
    step(IntC<0>{}, k);
    step(IntC<1>{}, k + 1);
    step(IntC<2>{}, k + 2);
    step(IntC<3>{}, k + 3);
  }

  // ---- fill-error reduction ------------------------------------------------------------------------------------------------------------
  if (P.red != nullptr) {
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
    if ((tid & 31) == 0) {
      if (rmax > -INFINITY) {
        atomicMax(P.red + 0, ord_key((double)rmax));
        atomicMax(P.red + 2, ((unsigned long long)ord_key32((float)rmax) << 32) | amax);
      }
      if (rmin < INFINITY) {
        atomicMin(P.red + 1, ord_key((double)rmin));
        atomicMin(P.red + 3, ((unsigned long long)ord_key32((float)rmin) << 32) | amin);
      }
      if (rnan) atomicAdd(P.red + 4, 1ull);
    }
  }
}

template <class T, int CPT, bool FUSED, bool KOREN, int NT, int MINB, bool SAMEU = false>
__global__ void __launch_bounds__(NT, MINB) xsweep_kernel(const SweepP<T> P, const int chunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const int ox = 2 + blockIdx.x * 32, nA = P.g.n[0];
  // tiles whose columns (incl. the halo -3..+33) reach a ghost column of a non-periodic x boundary take the body with the boundary rules
  const bool xb = !(P.g.per & 1u) && (ox - 3 < 2 || ox + 33 > nA - 1);
  if (xb) xsweep_body<T, CPT, FUSED, KOREN, true, NT, SAMEU>(P, chunk, sm);
  else xsweep_body<T, CPT, FUSED, KOREN, false, NT, SAMEU>(P, chunk, sm);
}

}  // namespace ifadv
