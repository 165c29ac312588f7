// ifadv_along.cuh -- fused directional sweep along y or z (J = 1, 2) for 3-D grids: REGISTER MARCHING.
//
// A CTA owns a tile of 32 (x) x TC (cross direction c) columns and marches along the sweep direction a = J through a
// chunk of planes; thread (tx, tc) owns ONE column.  Everything the stencil reaches ALONG the sweep -- the 4-point u★
// line of each component, the VOF / mass / momentum flux through the lower face, the previous dilation, the face
// velocities -- rolls through REGISTERS, so each flux is evaluated exactly once and never leaves the thread.  Shared
// memory only carries what a thread needs from its x-1 / c-1 neighbours (mass flux, dilation, f) and the
// cp.async-staged input planes:
//     f        8-slot ring, halo -2..+1 in x and c (the 3^3 PLIC box of a halo column), loaded 3 planes ahead
//     u_a,u⁰_a 4-slot rings (faces), ρu 4-slot ring x 3 components, uOld 2-slot ring x 3 components (no halo)
// Per plane: ONE wait + barrier (S1), issue the next plane's copies, u★(k+2), VOF flux / mass flux at face k+1,
// dilation(k), barrier (S2), [lane-dense PLIC of marked interface faces], SynDRoM fluxes at face k+1, update of cell k.
// A chunk starts 4 planes early (warm-up) to fill the register pipeline.
// Reference lines as in ifadv_march.cuh / ifadv_sweep.cuh (same boundary rules, same expression order).
#pragma once
#include "ifadv_march.cuh"

namespace ifadv {

template <int TCT> struct ATile {  // TCT = tile extent along c (rows) = 8 * columns-per-thread
  static constexpr int WX = 35, WC = TCT + 3, PLH = WX * WC, NC = 32 * TCT, NH = PLH - NC, NHU = 32 + TCT;
  static constexpr int RF = 8, RU = 4, RR = 4, RO = 2;
  // halo planes: F, U, U0, M x2, FX (+ Dil x2) ; core planes (CMOM): ρu, uOld ; + interface list + counter
  template <class T> static constexpr size_t smem_bytes(bool mom) {
    return sizeof(T) * ((size_t)PLH * (RF + 2 * RU + 3 + (mom ? 2 : 0)) + (mom ? (size_t)NC * 3 * (RR + RO) : 0)) + sizeof(int) * (PLH + 4);
  }
};

// 3^3 box accessor on the 8-slot f ring of the along kernel
template <class T, int J, int PLH, int WX> struct ABox {
  const T* sF;
  int e, pl;
  IFADV_DI T operator()(int dx, int dy, int dz) const {
    const int da = (J == 1) ? dy : dz, dc = (J == 1) ? dz : dy;
    return sF[((pl + da) & 7) * PLH + e + dx + dc * WX];
  }
};

template <class T, int J, int CPT, bool MOM, bool FUSED, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) along_kernel(const SweepP<T> P, const int chunk) {
  constexpr int TR = NT / 32;      // rows of threads
  constexpr int TCT = TR * CPT;    // tile rows: a thread owns CPT columns, rows tc, tc+TR, ...
  using TL = ATile<TCT>;
  static_assert(J == 1 || J == 2, "sweeps along x use the in-plane kernel");
  static_assert(TL::NH <= NT, "one halo entry per thread");
  constexpr int DCC = (J == 1) ? 2 : 1;  // global dimension of the cross direction c
  constexpr int WX = TL::WX, PLH = TL::PLH, NC = TL::NC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  T* sF = sm;
  T* sU = sF + TL::RF * PLH;
  T* sU0 = sU + TL::RU * PLH;
  T* sM = sU0 + TL::RU * PLH;                 // 2 slots (face & 1): mass flux
  T* sFX = sM + 2 * PLH;                      // fᶠ of reconstructed interface faces (handed back to the owning thread)
  T* sDil = sFX + PLH;                        // 2 slots (plane & 1)           [CMOM]
  T* sR = sDil + (MOM ? 2 * PLH : 0);         // [(plane&3)*3 + role][NC]     [CMOM]
  T* sO = sR + (MOM ? TL::RR * 3 * NC : 0);   // [(plane&1)*3 + role][NC]     [CMOM]
  int* sList = reinterpret_cast<int*>(sO + (MOM ? TL::RO * 3 * NC : 0));
  int* sCnt = sList + PLH;

  const Geo& g = P.g;
  const int tid = threadIdx.x, tx = tid & 31, tc = tid >> 5;
  const int nA = g.n[J], nX = g.n[0], nCc = g.n[DCC];
  const long long st3[3] = {1, g.s1, g.s2};
  const long long sA = st3[J], sCc = st3[DCC];
  const bool perA = (g.per >> J) & 1u, perX = g.per & 1u, perC = (g.per >> DCC) & 1u;
  const long long cA = P.coff[J], cC = P.coff[DCC];  // component offsets (host-computed; component x has offset 0)
  const int ox = 2 + blockIdx.x * 32, oc = 2 + blockIdx.y * TCT;
  const int k0 = 2 + blockIdx.z * chunk, k1 = min(k0 + chunk, nA);
  const T lr = P.lr, omlr = P.omlr, dt = P.dt;
  const T AA = P.A[J], AXv = P.A[0], ACv = P.A[DCC];
  if (tid == 0) *sCnt = 0;

  // ---- per-thread constants: own columns and (for the first NH threads) one halo entry ---------------------------------------
  const int vx = ox + tx;
  const bool dirX = !perX && (vx == 2 || vx == nX);
  int eo[CPT], go[CPT];
  bool valid[CPT], dirC[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int lc = tc + j * TR, vcc = oc + lc;
    eo[j] = (tx + 2) + WX * (lc + 2);
    go[j] = (mapc(vx, nX, perX) - 1) + (int)((mapc(vcc, nCc, perC) - 1) * sCc);
    valid[j] = vx <= nX - 1 && vcc <= nCc - 1;
    dirC[j] = !perC && (vcc == 2 || vcc == nCc);
  }
  int eh = 0, gh = 0;
  const bool hasH = tid < TL::NH;
  const bool hU = tid < TL::NHU;
  if (hasH) {
    // the 32+TCT entries that also carry face velocities / mass flux / dilation come first, so only the first warps
    // execute the halo flux / dilation code; the remaining entries only feed the 3^3 PLIC box
    int lx, lc;
    const int h = tid;
    if (h < 32) { lx = h; lc = -1; }
    else if (h < 32 + TCT) { lx = -1; lc = h - 32; }
    else if (h < 32 + TCT + WX) { lc = -2; lx = h - (32 + TCT) - 2; }
    else if (h < 35 + TCT + WX) { lc = -1; const int r = h - (32 + TCT + WX); lx = (r < 2) ? r - 2 : 32; }
    else if (h < 35 + TCT + 2 * WX) { lc = TCT; lx = h - (35 + TCT + WX) - 2; }
    else if (h < 35 + 2 * TCT + 2 * WX) { lx = -2; lc = h - (35 + TCT + 2 * WX); }
    else { lx = 32; lc = h - (35 + 2 * TCT + 2 * WX); }
    eh = (lx + 2) + WX * (lc + 2);
    gh = (mapc(ox + lx, nX, perX) - 1) + (int)((mapc(oc + lc, nCc, perC) - 1) * sCc);
  }

  // plane offsets along a: mapped (f, tangential components, c̄, uOld) and as stored (component a, face velocities).
  // They are block-uniform and roll with the march, so only pm(k+4), po(k+4) are evaluated per plane.
  auto pm = [&](int v) -> long long { return (long long)(map1(v, nA, perA) - 1) * sA; };
  auto po = [&](int v) -> long long { return (long long)(own1(v, nA, perA) - 1) * sA; };
  auto dirAf = [&](int v) -> bool { return !perA && (v == 1 || v == 2 || v == nA); };
  constexpr unsigned SZ = sizeof(T);
  const unsigned sb = (unsigned)__cvta_generic_to_shared(sm);
  const unsigned aFh = sb + eh * SZ, aUh = sb + (TL::RF * PLH + eh) * SZ;
  const unsigned aR = sb + (unsigned)((sR - sm) + tid) * SZ;  // ρu ring: + ((v&3)*3 + r)*NC*SZ + j*NT*SZ
  const unsigned aO = sb + (unsigned)((sO - sm) + tid) * SZ;  // uOld ring: + ((v&1)*3 + r)*NC*SZ + j*NT*SZ

  auto issue_f = [&](int v, long long pmv) {
    const unsigned so = sb + (unsigned)(v & 7) * (PLH * SZ);
    const T* fp = P.f_in + pmv;
#pragma unroll
    for (int j = 0; j < CPT; ++j) cp_async_s(so + eo[j] * SZ, fp + go[j]);
    if (hasH) cp_async_s(aFh + (unsigned)(v & 7) * (PLH * SZ), fp + gh);
  };
  auto issue_u = [&](int v, long long pov) {
    const unsigned so = (unsigned)(v & 3) * (PLH * SZ);
    const T* up = P.uj + pov;
    const T* u0p = P.u0j + pov;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      cp_async_s(sb + so + (TL::RF * PLH + eo[j]) * SZ, up + go[j]);
      cp_async_s(sb + so + ((TL::RF + TL::RU) * PLH + eo[j]) * SZ, u0p + go[j]);
    }
    if (MOM && hU) {
      cp_async_s(aUh + so, up + gh);
      cp_async_s(aUh + so + TL::RU * PLH * SZ, u0p + gh);
    }
  };
  constexpr bool fused = MOM && FUSED;
  const T* rsrc = fused ? P.uOld : P.rhou_in;  // fused sweep 1: the ρu ring carries uOld, ρu = BC!(uOld*ρ(f̄)) is formed on the fly
  auto issue_ru = [&](int v, long long pmv, long long pov) {
    const unsigned d = aR + (unsigned)((v & 3) * 3) * (NC * SZ);
    const T* ra = rsrc + cA + pov;
    const T* rx = rsrc + pmv;
    const T* rc = rsrc + cC + pmv;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      cp_async_s(d + j * NT * SZ, ra + go[j]);
      cp_async_s(d + (NC + j * NT) * SZ, rx + go[j]);
      cp_async_s(d + (2 * NC + j * NT) * SZ, rc + go[j]);
    }
  };
  auto issue_uold = [&](int v, long long pmv) {
    const unsigned d = aO + (unsigned)((v & 1) * 3) * (NC * SZ);
    const T* o0 = P.uOld + pmv;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      cp_async_s(d + j * NT * SZ, o0 + cA + go[j]);
      cp_async_s(d + (NC + j * NT) * SZ, o0 + go[j]);
      cp_async_s(d + (2 * NC + j * NT) * SZ, o0 + cC + go[j]);
    }
  };
  int cbo_n[CPT], cbh_n = 0;  // c̄ of the next plane (own columns, halo entry), read one plane ahead
#pragma unroll
  for (int j = 0; j < CPT; ++j) cbo_n[j] = 0;
  auto load_cbar = [&](long long pmv) {
    if (!P.first) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) cbo_n[j] = (int)P.cbar[pmv + go[j]];
      if (MOM && hU) cbh_n = (int)P.cbar[pmv + gh];
    }
  };

  // ---- rolling register state -------------------------------------------------------------------------------------------------------
  T usA[CPT][4], usX[CPT][4], usC[CPT][4];  // u★ at planes k-1, k, k+1, k+2
  T FloA[CPT], FloX[CPT], FloC[CPT], FFlo[CPT], Mlo[CPT], dilm1[CPT], uk[CPT], u0k[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
#pragma unroll
    for (int i = 0; i < 4; ++i) usA[j][i] = usX[j][i] = usC[j][i] = T(0);
    FloA[j] = FloX[j] = FloC[j] = FFlo[j] = Mlo[j] = dilm1[j] = T(0);
  }
  T rmax = -INFINITY, rmin = INFINITY;
  unsigned int amax = 0, amin = 0;
  int rnan = 0;

  // ---- prologue: planes needed by the first warm-up step ks = k0-4 ----------------------------------------------------------------------
  const int ks = k0 - 4;
  issue_f(ks - 1, pm(ks - 1)); issue_f(ks, pm(ks)); issue_f(ks + 1, pm(ks + 1)); issue_f(ks + 2, pm(ks + 2));
  issue_u(ks, po(ks)); issue_u(ks + 1, po(ks + 1));
  if (MOM) {
    issue_ru(ks, pm(ks), po(ks)); issue_ru(ks + 1, pm(ks + 1), po(ks + 1)); issue_ru(ks + 2, pm(ks + 2), po(ks + 2));
    if (!fused) issue_uold(ks, pm(ks));
  }
  cp_async_commit();
  load_cbar(pm(ks));
  cp_async_wait_all();
  __syncthreads();
  long long pmA = pm(ks + 1), pmB = pm(ks + 2), pmC = pm(ks + 3), poB = po(ks + 2), poC = po(ks + 3);  // pm(k+1..k+3), po(k+2..k+3)
#pragma unroll
  for (int j = 0; j < CPT; ++j) { uk[j] = sU[(ks & 3) * PLH + eo[j]]; u0k[j] = sU0[(ks & 3) * PLH + eo[j]]; }
  long long lk0 = (long long)(ks - 1) * sA;  // (k-1)*sA, rolled with the march

#pragma unroll 1
  for (int k = ks; k < k1; ++k) {
    const int p = k + 1, q = k + 2;
    if (k > ks) {
      cp_async_wait_all();
      __syncthreads();  // S1: everything issued during step k-1 has landed; all reads of step k-1 are done
    }
    int cbo[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) cbo[j] = cbo_n[j];
    const int cbh = cbh_n;
    const bool dq = dirAf(q), dp = dirAf(p), dpm = dirAf(k);
    // A. next plane's copies (HBM latency hides behind this plane's arithmetic)
    issue_f(k + 3, pmC);
    issue_u(k + 2, poB);
    if (MOM) { issue_ru(k + 3, pmC, poC); if (!fused) issue_uold(k + 1, pmA); }
    cp_async_commit();
    load_cbar(pmA);
    pmA = pmB; pmB = pmC; poB = poC;
    {  // pm(k+4), po(k+4): one stride further unless the plane is within the boundary band (block-uniform)
      const int v = k + 4;
      if (v >= 4 && v <= nA - 2) { pmC += sA; poC += sA; }
      else { pmC = pm(v); poC = po(v); }
    }

    const T* Fk = sF + (k & 7) * PLH;
    const T* Fp = sF + (p & 7) * PLH;
    const T* Fq = sF + (q & 7) * PLH;
    const T* Up = sU + (p & 3) * PLH;
    const T* U0p = sU0 + (p & 3) * PLH;
    T* Mp = sM + (p & 1) * PLH;
    const bool needp = p <= nA && (perA || p >= 2);
    T FFhi[CPT], Mhi[CPT], div[CPT], fK[CPT], dilk[CPT], up1[CPT], u0p1[CPT];
    int cb[CPT];
    bool marked[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int e = eo[j];
      // B. u★ of plane q = k+2 (flow.jl:197): BC!(ρu/ρ(f̄))
      if (MOM) {
        const T* R = sR + ((q & 3) * 3) * NC + tid + j * NT;
        const T fq = Fq[e];
        const T ha = rho_face(fq, Fp[e], lr, omlr), hx = rho_face(fq, Fq[e - 1], lr, omlr),
                hc = rho_face(fq, Fq[e - WX], lr, omlr);
        // fused: ρu = u*ρ (u2ρu!, VOFutil.jl:208-211) and straight back to u★ = ρu/ρ, rounding as the two passes would
        const T ra = t_div(fused ? R[0] * ha : R[0], ha);
        const T rx = t_div(fused ? R[NC] * hx : R[NC], hx);
        const T rc = t_div(fused ? R[2 * NC] * hc : R[2 * NC], hc);
        usA[j][3] = dq ? AA : ra;  // Dirichlet planes of BC!
        usX[j][3] = dirX ? AXv : rx;
        usC[j][3] = dirC[j] ? ACv : rc;
      }
      // C. VOF flux + mass flux through face p = k+1 (advection.jl:108-137)
      up1[j] = Up[e]; u0p1[j] = U0p[e];
      FFhi[j] = T(0); Mhi[j] = T(0); marked[j] = false;
      if (needp) {
        const T dl = P.hdt * (up1[j] + u0p1[j]);  // δt/2*(u+u⁰)
        if (dl != T(0)) {
          const bool up = dl > T(0);
          const T fc = up ? Fk[e] : Fp[e];  // upwind cell
          const int cu = up ? k : p;
          const bool ghost = !perA && (cu < 2 || cu > nA - 1);
          if (ghost || fullorempty(fc)) {
            FFhi[j] = fc * dl;
            Mhi[j] = dl * lr + omlr * FFhi[j];
            if (MOM) Mhi[j] = Mhi[j] * P.idt;
          } else { sList[atomicAdd(sCnt, 1)] = e; marked[j] = true; }
        }
      }
      Mp[e] = Mhi[j];
      // D. dilation of plane k (flow.jl:216)
      div[j] = (up1[j] - uk[j]) + (u0p1[j] - u0k[j]);  // ∂(d,I,u)+∂(d,I,u⁰)
      fK[j] = Fk[e];
      cb[j] = P.first ? ((fK[j] < T(0.5)) ? 0 : 1) : cbo[j];  // flow.jl:172 (c̄ from the incoming f)
      dilk[j] = T(0);
      if (MOM) {
        dilk[j] = (lin_interp(T(cb[j]), lr, omlr) * div[j]) / T(2);
        sDil[(k & 1) * PLH + e] = dilk[j];
      }
    }
    // C/D for the halo entry (mass flux and dilation that the x-1 / c-1 neighbours of the tile edge need)
    if (MOM && hU) {
      T m = T(0);
      if (needp) {
        const T dl = P.hdt * (Up[eh] + U0p[eh]);
        if (dl != T(0)) {
          const bool up = dl > T(0);
          const T fc = up ? Fk[eh] : Fp[eh];
          const int cu = up ? k : p;
          const bool ghost = !perA && (cu < 2 || cu > nA - 1);
          if (ghost || fullorempty(fc)) {
            m = dl * lr + omlr * (fc * dl);
            m = m * P.idt;
          } else sList[atomicAdd(sCnt, 1)] = eh;
        }
      }
      Mp[eh] = m;
      const T* Uk = sU + (k & 3) * PLH;
      const T* U0k = sU0 + (k & 3) * PLH;
      const T dh = (Up[eh] - Uk[eh]) + (U0p[eh] - U0k[eh]);
      const int ch = P.first ? ((Fk[eh] < T(0.5)) ? 0 : 1) : cbh;
      sDil[(k & 1) * PLH + eh] = (lin_interp(T(ch), lr, omlr) * dh) / T(2);
    }
    __syncthreads();  // S2: mass flux of face p, dilation of plane k and the interface list are complete
    {
      const int cnt = *sCnt;  // block-uniform
      if (cnt > 0) {
        // lane-dense PLIC reconstruction of the marked faces (general branch of getVOFFlux!, advection.jl:131-134)
        for (int i = tid; i < cnt; i += NT) {
          const int e = sList[i];
          const T dl = P.hdt * (Up[e] + U0p[e]);
          const int pl = (dl > T(0)) ? k : p;
          ABox<T, J, PLH, WX> B{sF, e, pl};
          const T ff = plic_face_flux<T, 3>(P.scheme, B, sF[(pl & 7) * PLH + e], J, dl);
          T m = dl * lr + omlr * ff;
          if (MOM) m = m * P.idt;
          sFX[e] = ff;
          Mp[e] = m;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          if (marked[j]) { FFhi[j] = sFX[eo[j]]; Mhi[j] = Mp[eo[j]]; }
      }
    }
    const long long lkp = lk0;  // (k-1)*sA
    lk0 += sA;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int e = eo[j];
      // G. SynDRoM momentum flux through face p of the three momentum cells of this column (flow.jl:20-57,223)
      T FhiA = T(0), FhiX = T(0), FhiC = T(0);
      if (MOM) {
        const T Mc = dp ? AA : Mhi[j];  // velocity BC! on ρuf (flow.jl:207)
        const bool Lvar = !perA && p == 2, Rvar = !perA && p == nA;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          T Mo;
          if (r == 0) Mo = dpm ? AA : Mlo[j];
          else if (r == 1) Mo = dp ? AA : Mp[e - 1];
          else Mo = dp ? AA : Mp[e - WX];
          const T Psi = (Mc + Mo) / T(2);
          const T* us = (r == 0) ? usA[j] : ((r == 1) ? usX[j] : usC[j]);
          const bool pos = Psi > T(0);
          T uu, cc, dd;
          if (Lvar) {  // ϕuL
            if (pos) { uu = T(2) * us[1] - us[2]; cc = us[1]; dd = us[2]; }
            else { uu = us[3]; cc = us[2]; dd = us[1]; }
          } else if (Rvar) {  // ϕuR
            if (Psi < T(0)) { uu = T(2) * us[2] - us[1]; cc = us[2]; dd = us[1]; }
            else { uu = us[0]; cc = us[1]; dd = us[2]; }
          } else {  // ϕu
            uu = pos ? us[0] : us[3];
            cc = pos ? us[1] : us[2];
            dd = pos ? us[2] : us[1];
          }
          // donor momentum cell: plane k (Ψ>0) or p; its face-centred old f (dρ after f2face!+BCv!, flow.jl:205)
          const T* Fd = pos ? Fk : Fp;
          T fo;
          if (r == 0) {
            const T* Fdm = pos ? sF + ((k - 1) & 7) * PLH : Fk;
            fo = (Fd[e] + Fdm[e]) / T(2);
            if (Lvar && pos) fo = (Fq[e] + Fp[e]) / T(2);  // donor index 1: BCv! copies plane 3 = (f(3)+f(2))/2
            if (Rvar && !pos) fo = __ldg(P.drho + cA + (long long)(nA - 1) * sA + go[j]);  // donor index nA: never written by f2face!
          } else if (r == 1) fo = (Fd[e] + Fd[e - 1]) / T(2);
          else fo = (Fd[e] + Fd[e - WX]) / T(2);
          const T fl = syndrom_flux(P.lim, Psi, uu, cc, dd, lin_interp(fo, lr, omlr), dt);
          if (r == 0) FhiA = fl; else if (r == 1) FhiX = fl; else FhiC = fl;
        }
      }
      // H. update of cell k
      if (k >= k0 && valid[j]) {
        const long long lk = lkp + go[j];
        if (P.first) P.cbar[lk] = (int8_t)cb[j];
        T fn = fK[j] + ((FFlo[j] - FFhi[j]) + ((T(cb[j]) * div[j]) * dt) / T(2));  // advection.jl:83
        if (fn != fn) rnan = 1;
        if (fn > rmax) { rmax = fn; amax = (unsigned int)lk; }
        if (fn < rmin) { rmin = fn; amin = (unsigned int)lk; }
        fn = (fn < P.tol) ? T(0) : ((fn > P.onemtol) ? T(1) : fn);  // cleanWisp!
        P.f_out[lk] = fn;
        if (!MOM && P.rhouf_j != nullptr) {
          P.rhouf_j[lk] = Mlo[j];
          if (k == nA - 1) P.rhouf_j[lk + sA] = Mhi[j];  // inside_uWB includes the upper boundary face
        }
        if (MOM) {
          const T* R = sR + ((k & 3) * 3) * NC + tid + j * NT;
          const T* O = fused ? R : sO + ((k & 1) * 3) * NC + tid + j * NT;
          const T* Dk = sDil + (k & 1) * PLH;
          const T dNa = (!perA && k == 2) ? dilk[j] : dilm1[j];  // BCf! (Neumann) on ρ̄∂ⱼuⱼ along the sweep direction
          T qA = R[0], qX = R[NC], qC = R[2 * NC];  // ρu before the sweep
          if (fused) {  // u2ρu! + BC!(ρu,uBC): Dirichlet plane 2 of the normal component holds uBC
            const T* Fm = sF + ((k - 1) & 7) * PLH;
            qA = dpm ? AA : qA * rho_face(fK[j], Fm[e], lr, omlr);
            qX = dirX ? AXv : qX * rho_face(fK[j], Fk[e - 1], lr, omlr);
            qC = dirC[j] ? ACv : qC * rho_face(fK[j], Fk[e - WX], lr, omlr);
          }
          // r = Φ[I] - Φ[I+δj] + uOld*ϕ(i,I,ρ̄∂ⱼuⱼ);  ρu += δt*r          flow.jl:223-231
          const T rA = (FloA[j] - FhiA) + O[0] * ((dilk[j] + dNa) / T(2));
          const T rX = (FloX[j] - FhiX) + O[NC] * ((dilk[j] + Dk[e - 1]) / T(2));
          const T rC = (FloC[j] - FhiC) + O[2 * NC] * ((dilk[j] + Dk[e - WX]) / T(2));
          P.rhou_out[cA + lk] = qA + dt * rA;
          P.rhou_out[lk] = qX + dt * rX;
          P.rhou_out[cC + lk] = qC + dt * rC;
        }
      }
      // I. roll the register pipeline
      FloA[j] = FhiA; FloX[j] = FhiX; FloC[j] = FhiC; FFlo[j] = FFhi[j]; Mlo[j] = Mhi[j]; dilm1[j] = dilk[j];
      uk[j] = up1[j]; u0k[j] = u0p1[j];
#pragma unroll
      for (int i = 0; i < 3; ++i) { usA[j][i] = usA[j][i + 1]; usX[j][i] = usX[j][i + 1]; usC[j][i] = usC[j][i + 1]; }
    }
    if (tid == 0) *sCnt = 0;
  }

  // ---- fill-error reduction ------------------------------------------------------------------------------------------------------------
  if (P.red != nullptr) {
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
      red_commit<T>(P.red, rmax, rmin, amax, amin, rnan);
    }
  }
}

}  // namespace ifadv
