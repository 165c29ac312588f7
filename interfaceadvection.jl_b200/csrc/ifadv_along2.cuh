// ifadv_along2.cuh -- fused directional sweep along y or z (J = 1, 2) for 3-D grids: LEAN register marching.
//
// Same tile / shared-memory layout and the same arithmetic (expression by expression) as its retired first generation, restructured
// so that the interior of the domain runs a branch-free body with compile-time ring slots and ONE barrier per plane:
//   * software skew: step k evaluates the VOF / mass flux of face k+2 and the dilation of plane k+1 (written to the
//     shared M / Dil rings for the x-1 / c-1 neighbours) and, in the same step, the SynDRoM fluxes of face k+1 and the
//     update of cell k from the ring slots written one step earlier -- the barrier at the top of the step is the only one;
//     interface faces marked in step k are reconstructed lane-dense at the top of step k+1 (second barrier only then);
//   * FAST steps (3 <= k <= nA-4: every plane the step touches is a plain interior plane) are instantiated four times
//     with static ring slots ((k-ks)&3 known at compile time; the 8-slot f ring as two half rings whose bases swap per
//     group) and every boundary predicate folded to a constant; the remaining steps run the GENERIC instantiation of
//     the same body (Dirichlet planes of BC!, ϕuL/ϕuR, ghost upwind cells, clamped / wrapped plane offsets);
//   * 32-bit element offsets (one add per column and plane, one IMAD.WIDE per access) instead of 64-bit offset chains;
//   * face densities ρ(f̄) of the three faces of a cell are evaluated once (for u★) and carried in registers for the
//     SynDRoM donor density of the next two faces and for the fused u2ρu! product;
//   * fill-error extrema by FMNMX (NaN-propagating max); the location is only tracked for cells outside [0,1].
// Reference lines as in ifadv_march.cuh / ifadv_sweep.cuh.
#pragma once
#include "ifadv_march.cuh"

namespace ifadv {

// tile of the register-marching kernels: 32 columns along x, TCT rows along the cross direction c, halo of 1 below / 2 above in both
template <int TCT> struct ATile {  // TCT = tile extent along c (rows) = 8 * columns-per-thread
  static constexpr int WX = 35, WC = TCT + 3, PLH = WX * WC, NC = 32 * TCT, NH = PLH - NC, NHU = 32 + TCT;
  static constexpr int RF = 8, RU = 4, RR = 4, RO = 2;
  // halo planes: F, U, U0, M x2, FX (+ Dil x2) ; core planes (CMOM): ρu, uOld ; + interface list + counter
  template <class T> static constexpr size_t smem_bytes(bool mom) {
    return sizeof(T) * ((size_t)PLH * (RF + 2 * RU + 3 + (mom ? 2 : 0)) + (mom ? (size_t)NC * 3 * (RR + RO) : 0)) + sizeof(int) * (PLH + 4);
  }
};


template <bool B> struct BoolC { static constexpr bool value = B; };
template <int V> struct IntC { static constexpr int value = V; };

IFADV_DI float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
IFADV_DI double max_nan(double a, double b) { return (a != a || b != b) ? (a + b) : fmax(a, b); }

// 3^3 box accessor on the 8-slot f ring, ring phase relative to the chunk start
template <class T, int J, int PLH, int WX> struct A2Box {
  const T* sF;
  int e, rel;  // rel = (plane of the box centre) - ks
  IFADV_DI T operator()(int dx, int dy, int dz) const {
    const int da = (J == 1) ? dy : dz, dc = (J == 1) ? dz : dy;
    return sF[((rel + da) & 7) * PLH + e + dx + dc * WX];
  }
};

// SAMEU: u¹ and u² are ONE array (both advectfq! calls of MPFMomStep!: flow.jl:92, and :70 where u⁰≡u) -- the u⁰ copy stream, its
// ring and its registers drop out; the arithmetic is unchanged (u + u, (u2-u1) + (u2-u1)), so the results are bit-identical.
template <class T, int J, int CPT, bool MOM, bool FUSED, bool KOREN, int NT, int MINB, bool SAMEU = false>
__global__ void __launch_bounds__(NT, MINB) along2_kernel(const SweepP<T> P, const int chunk) {
  constexpr int TR = NT / 32;      // rows of threads
  constexpr int TCT = TR * CPT;    // tile rows: a thread owns CPT columns, rows tc, tc+TR, ...
  using TL = ATile<TCT>;
  static_assert(J == 1 || J == 2, "sweeps along x use the in-plane kernel");
  static_assert(TL::NH <= NT, "one halo entry per thread");
  constexpr int DCC = (J == 1) ? 2 : 1;  // global dimension of the cross direction c
  constexpr int WX = TL::WX, PLH = TL::PLH, NC = TL::NC;
  constexpr bool fused = MOM && FUSED;
  const bool first = fused ? true : (P.first != 0);  // the fused sweep is always sweep 1
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  // element offsets of the shared arrays
  constexpr int OF = 0;                                   // f ring, 8 slots
  constexpr int OU = OF + TL::RF * PLH;                   // u_a ring, 4 slots
  constexpr int OU0 = OU + TL::RU * PLH;                  // u⁰_a ring, 4 slots
  constexpr int OM = OU0 + TL::RU * PLH;                  // mass flux, 2 slots (face & 1)
  constexpr int OFX = OM + 2 * PLH;                       // fᶠ of reconstructed interface faces
  constexpr int ODIL = OFX + PLH;                         // dilation, 2 slots (plane & 1)           [CMOM]
  constexpr int OR = ODIL + (MOM ? 2 * PLH : 0);          // ρu ring  [(plane&3)*3 + role][NC]       [CMOM]
  constexpr int OO = OR + (MOM ? TL::RR * 3 * NC : 0);    // uOld ring [(plane&1)*3 + role][NC]      [CMOM]
  constexpr int OEND = OO + (MOM ? TL::RO * 3 * NC : 0);
  int* sList = reinterpret_cast<int*>(sm + OEND);
  int* sCnt = sList + PLH;  // 4 counters: marks of step k go to counter (k+1-ks)&3

  const Geo& g = P.g;
  const int tid = threadIdx.x, tx = tid & 31, tc = tid >> 5;
  const int nA = g.n[J], nX = g.n[0], nCc = g.n[DCC];
  const unsigned sA = (unsigned)((J == 1) ? g.s1 : g.s2), sCc = (unsigned)((DCC == 1) ? g.s1 : g.s2);
  const bool perA = (g.per >> J) & 1u, perX = g.per & 1u, perC = (g.per >> DCC) & 1u;
  // dimension 3 is restricted to the planes [kz0, kz1) (z-slabs): the march range for J == 2, the tile rows for J == 1
  const int ox = 2 + blockIdx.x * 32, oc = ((J == 1) ? P.kz0 : 2) + blockIdx.y * TCT;
  const int chi = (J == 1) ? P.kz1 : nCc;  // first cross index not updated
  const int k0 = ((J == 2) ? P.kz0 : 2) + blockIdx.z * chunk, k1 = min(k0 + chunk, (J == 2) ? P.kz1 : nA);
  const int ks = k0 - 4;
  const T lr = P.lr, omlr = P.omlr, dt = P.dt;
  const T lam1 = lin_interp(T(1), lr, omlr);
  const T AA = P.A[J], AXv = P.A[0], ACv = P.A[DCC];
  if (tid < 4) sCnt[tid] = 0;

  // one 64-bit base per array; component offsets (32-bit element offsets, component x has offset 0) go into the thread offsets
  const T* const rsrc = fused ? P.uOld : P.rhou_in;  // fused sweep 1: the ρu ring carries uOld, ρu = BC!(uOld*ρ(f̄)) on the fly
  const unsigned cA = (unsigned)P.coff[J], cC = (unsigned)P.coff[DCC];
  const T* const ubase = P.u;    // u[:, 1]
  const T* const u0base = P.u0;  // u⁰[:, 1]

  // ---- per-thread constants: own columns and (for the first NH threads) one halo entry -----------------------------------------
  const int vx = ox + tx;
  const bool dirX = !perX && (vx == 2 || vx == nX);
  const int e0 = (tx + 2) + WX * (tc + 2);  // shared entry of column 0; column j adds j*TR*WX
  unsigned go[CPT];
  bool valid[CPT], dirC[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int vcc = oc + tc + j * TR;
    go[j] = (unsigned)(mapc(vx, nX, perX) - 1) + (unsigned)(mapc(vcc, nCc, perC) - 1) * sCc;
    valid[j] = vx <= nX - 1 && vcc < chi;
    dirC[j] = !perC && (vcc == 2 || vcc == nCc);
  }
  int eh = 0;
  unsigned gh = 0;
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
    gh = (unsigned)(mapc(ox + lx, nX, perX) - 1) + (unsigned)(mapc(oc + lc, nCc, perC) - 1) * sCc;
  }

  // plane offsets along a (block-uniform, 32-bit element offsets): mapped (f, tangential components, c̄, uOld) and as
  // stored (component a, face velocities)
  auto pm = [&](int v) -> unsigned { return (unsigned)(map1(v, nA, perA) - 1) * sA; };
  auto po = [&](int v) -> unsigned { return (unsigned)(own1(v, nA, perA) - 1) * sA; };
  auto dirAf = [&](int v) -> bool { return !perA && (v == 1 || v == 2 || v == nA); };
  constexpr unsigned SZ = sizeof(T);
  const unsigned sb = (unsigned)__cvta_generic_to_shared(sm);
  const unsigned se0 = sb + (unsigned)e0 * SZ;   // + slot offsets (+ j*TR*WX*SZ)
  const unsigned seh = sb + (unsigned)eh * SZ;
  const unsigned st0 = sb + (unsigned)tid * SZ;  // core planes: + (OR/OO + ..)*SZ + j*NT*SZ

  // ---- rolling register state (see the header: values entering step k) ---------------------------------------------------------
  T us[3][CPT][4];                                 // u★ of planes k-1, k, k+1, (k+2): [A, X, C]
  T Flo[3][CPT];                                   // SynDRoM flux through face k
  T FFlo[CPT], Mlo[CPT], FFhi[CPT], Mhi[CPT];      // VOF / mass flux through faces k, k+1
  T dilm1[CPT], dil0[CPT], dv0[CPT];               // dilation of planes k-1, k; c̄(∂u+∂u⁰)δt/2 of plane k
  T f0[CPT], f1[CPT], u1[CPT], u01[CPT];           // f(k), f(k+1); u_a, u⁰_a at face k+1
  T h0[3][CPT], h1[3][CPT];                        // ρ at the lower a / x / c faces of cells k, k+1
  bool mk[CPT];                                    // face k+1 of this column is an interface face (reconstructed lane-dense)
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int i = 0; i < 4; ++i) us[r][j][i] = T(0);
      Flo[r][j] = T(0); h0[r][j] = T(1); h1[r][j] = T(1);
    }
    FFlo[j] = Mlo[j] = FFhi[j] = Mhi[j] = dilm1[j] = dil0[j] = dv0[j] = T(0);
    mk[j] = false;
  }
  T rmax = -INFINITY, rmin = INFINITY;
  unsigned int amax = 0, amin = 0;

  // ---- prologue: planes needed by the first step ks --------------------------------------------------------------------------------
  {
    auto ld_f = [&](int v) {
      const T* fp = P.f_in + pm(v);
      const unsigned so = (unsigned)(((v - ks) & 7) * PLH + OF) * SZ;
#pragma unroll
      for (int j = 0; j < CPT; ++j) cp_async_s(se0 + so + j * (TR * WX * SZ), fp + go[j]);
      if (hasH) cp_async_s(seh + so, fp + gh);
    };
    auto ld_u = [&](int v) {
      const unsigned so = (unsigned)(((v - ks) & 3) * PLH) * SZ;
      const T* up = ubase + (po(v) + cA);
      const T* u0p = u0base + (po(v) + cA);
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        cp_async_s(se0 + so + OU * SZ + j * (TR * WX * SZ), up + go[j]);
        if (!SAMEU) cp_async_s(se0 + so + OU0 * SZ + j * (TR * WX * SZ), u0p + go[j]);
      }
      if (MOM && hU) {
        cp_async_s(seh + so + OU * SZ, up + gh);
        if (!SAMEU) cp_async_s(seh + so + OU0 * SZ, u0p + gh);
      }
    };
    auto ld_ru = [&](int v) {
      const unsigned d = st0 + (unsigned)(OR + ((v - ks) & 3) * 3 * NC) * SZ;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        cp_async_s(d + j * NT * SZ, rsrc + (po(v) + cA + go[j]));
        cp_async_s(d + (NC + j * NT) * SZ, rsrc + (pm(v) + go[j]));
        cp_async_s(d + (2 * NC + j * NT) * SZ, rsrc + (pm(v) + cC + go[j]));
      }
    };
    ld_f(ks); ld_f(ks + 1); ld_f(ks + 2);
    ld_u(ks + 1); ld_u(ks + 2);
    if (MOM) {
      ld_ru(ks); ld_ru(ks + 1); ld_ru(ks + 2);
      if (!fused) {
        const unsigned d = st0 + (unsigned)(OO + 0) * SZ;  // plane ks -> slot 0
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          cp_async_s(d + j * NT * SZ, P.uOld + (pm(ks) + cA + go[j]));
          cp_async_s(d + (NC + j * NT) * SZ, P.uOld + (pm(ks) + go[j]));
          cp_async_s(d + (2 * NC + j * NT) * SZ, P.uOld + (pm(ks) + cC + go[j]));
        }
      }
    }
    cp_async_commit();
    // the M / Dil rings are read one step after they are written: start from zeros
    for (int i = tid; i < 2 * PLH; i += NT) { sm[OM + i] = T(0); if (MOM) sm[ODIL + i] = T(0); }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int e = e0 + j * TR * WX;
      f0[j] = sm[OF + e];                                  // plane ks   -> slot 0
      f1[j] = sm[OF + 1 * PLH + e];                        // plane ks+1 -> slot 1
      u1[j] = sm[OU + 1 * PLH + e];
      u01[j] = SAMEU ? u1[j] : sm[OU0 + 1 * PLH + e];
    }
  }

  // uniform rolling offsets
  unsigned lkU = (unsigned)(ks - 1) * sA;  // (k-1)*sA: offset of plane k when it is a plain interior plane
  unsigned fLo = 0, fHi = 4 * PLH;         // element offsets of the two half rings of f: slots (k-ks)&7 in 0..3 / 4..7 of the current group


  auto step = [&](auto fastc, auto ic, const int k) {
    constexpr bool FAST = decltype(fastc)::value;
    constexpr int I = decltype(ic)::value;  // (k-ks)&3 when FAST
    const int rel = k - ks;
    // ring slots (element offsets / indices) of plane k+d
    auto sF8 = [&](int d) -> unsigned {
      if (FAST) { const int s = I + d; return (s >= 0 && s < 4) ? fLo + s * PLH : ((s >= 4) ? fHi + (s - 4) * PLH : fHi + (s + 4) * PLH); }
      return (unsigned)(((rel + d) & 7) * PLH);
    };
    auto s4 = [&](int d) -> unsigned { return FAST ? (unsigned)((I + d) & 3) : (unsigned)((rel + d) & 3); };
    auto s2 = [&](int d) -> unsigned { return FAST ? (unsigned)((I + d) & 1) : (unsigned)((rel + d) & 1); };

    // c̄(k+1) for this step's dilation: the read is issued before the barrier, so its latency overlaps the wait
    int cb1[CPT];
    int cbh1 = 0;
#pragma unroll
    for (int j = 0; j < CPT; ++j) cb1[j] = 0;
    if (!first) {
      const unsigned oc1 = FAST ? lkU + sA : pm(k + 1);
#pragma unroll
      for (int j = 0; j < CPT; ++j) cb1[j] = (int)P.cbar[go[j] + oc1];
      if (MOM && hU) cbh1 = (int)P.cbar[gh + oc1];
    }
    cp_async_wait_all();
    __syncthreads();  // S1: the copies issued during step k-1 have landed; every read / write of step k-1 is done

    // block-uniform boundary rules of this step (constants in FAST steps)
    const bool dq = FAST ? false : dirAf(k + 2), dp = FAST ? false : dirAf(k + 1), dpm = FAST ? false : dirAf(k);
    const bool Lvar = FAST ? false : (!perA && k + 1 == 2), Rvar = FAST ? false : (!perA && k + 1 == nA);
    const bool needn = FAST ? true : (k + 2 <= nA && (perA || k + 2 >= 2));      // face k+2 carries a flux
    const bool ghU = FAST ? false : (!perA && (k + 1 < 2 || k + 1 > nA - 1));    // cell k+1 is a ghost cell on a non-periodic side
    const bool ghD = FAST ? false : (!perA && (k + 2 < 2 || k + 2 > nA - 1));    // cell k+2
    const bool store = k >= k0;

    // A. next planes' copies: f(k+3), u_a/u⁰_a(k+3), ρu(k+3), uOld(k+1), c̄(k+2).  One 64-bit base per array; component offsets
    //    and the plane lookahead are part of the 32-bit element offsets (block-uniform part + go[j]).
    {
      unsigned oM3, oO3, oM1;  // offsets of plane k+3 (mapped / as stored), k+1, k+2
      if (FAST) { oM3 = oO3 = lkU + 3 * sA; oM1 = lkU + sA; }  // lkU = (k-1)*sA = offset of plane k
      else { oM3 = pm(k + 3); oO3 = po(k + 3); oM1 = pm(k + 1); }
      const unsigned dF = (OF + sF8(3)) * SZ, dU = (OU + s4(3) * PLH) * SZ, dU0 = (OU0 + s4(3) * PLH) * SZ;
      const unsigned dR = (OR + s4(3) * 3 * NC) * SZ, dO = (OO + s2(1) * 3 * NC) * SZ;
      const unsigned oA3 = oO3 + cA, oC3 = oM3 + cC, oA1 = oM1 + cA, oC1 = oM1 + cC;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        cp_async_s(se0 + dF + j * (TR * WX * SZ), P.f_in + (go[j] + oM3));
        cp_async_s(se0 + dU + j * (TR * WX * SZ), ubase + (go[j] + oA3));
        if (!SAMEU) cp_async_s(se0 + dU0 + j * (TR * WX * SZ), u0base + (go[j] + oA3));
        if (MOM) {
          cp_async_s(st0 + dR + j * NT * SZ, rsrc + (go[j] + oA3));
          cp_async_s(st0 + dR + (NC + j * NT) * SZ, rsrc + (go[j] + oM3));
          cp_async_s(st0 + dR + (2 * NC + j * NT) * SZ, rsrc + (go[j] + oC3));
          if (!fused) {
            cp_async_s(st0 + dO + j * NT * SZ, P.uOld + (go[j] + oA1));
            cp_async_s(st0 + dO + (NC + j * NT) * SZ, P.uOld + (go[j] + oM1));
            cp_async_s(st0 + dO + (2 * NC + j * NT) * SZ, P.uOld + (go[j] + oC1));
          }
        }
      }
      if (hasH) cp_async_s(seh + dF, P.f_in + (gh + oM3));
      if (MOM && hU) {
        cp_async_s(seh + dU, ubase + (gh + oA3));
        if (!SAMEU) cp_async_s(seh + dU0, u0base + (gh + oA3));
      }
      cp_async_commit();
    }

    // P. lane-dense PLIC reconstruction of the interface faces k+1 marked in step k-1 (general branch of getVOFFlux!, advection.jl:131-134)
    {
      const int cnt = sCnt[rel & 3];  // block-uniform
      if (cnt > 0) {
        const T* Up = sm + OU + ((rel + 1) & 3) * PLH;
        const T* U0p = SAMEU ? Up : sm + OU0 + ((rel + 1) & 3) * PLH;
        T* Mp = sm + OM + ((rel + 1) & 1) * PLH;
        for (int i = tid; i < cnt; i += NT) {
          const int e = sList[i];
          const T dl = P.hdt * (Up[e] + U0p[e]);
          const int pr = (dl > T(0)) ? rel : rel + 1;  // upwind cell: plane k or k+1
          A2Box<T, J, PLH, WX> B{sm + OF, e, pr};
          // inlined: an ABI call here would force the whole rolling register state of the march through the stack
          const T ff = plic_face_flux_inl<T, 3>(P.scheme, B, sm[OF + (pr & 7) * PLH + e], J, dl);
          T m = dl * lr + omlr * ff;
          if (MOM) m = m * P.idt;
          sm[OFX + e] = ff;
          Mp[e] = m;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          if (mk[j]) { FFhi[j] = sm[OFX + e0 + j * TR * WX]; Mhi[j] = Mp[e0 + j * TR * WX]; }
      }
      if (tid == 0) sCnt[(rel - 1) & 3] = 0;
    }

    const unsigned qF = OF + sF8(2);                                  // f(k+2)
    const unsigned qU = OU + s4(2) * PLH, qU0 = OU0 + s4(2) * PLH;    // u_a, u⁰_a at face k+2
    const unsigned wM = OM + s2(2) * PLH, rM = OM + s2(1) * PLH;      // mass flux of face k+2 (written), of face k+1 (read)
    const unsigned wD = ODIL + s2(1) * PLH, rD = ODIL + s2(0) * PLH;  // dilation of plane k+1 (written), of plane k (read)
    T h2[3][CPT], FFn[CPT], Mn[CPT], dv1[CPT], dil1[CPT], f2[CPT], u2[CPT], u02[CPT];
    bool mkn[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int e = e0 + j * TR * WX;
      // B. u★ of plane q = k+2 (flow.jl:197): BC!(ρu/ρ(f̄))
      f2[j] = sm[qF + e];
      h2[0][j] = h2[1][j] = h2[2][j] = T(1);
      if (MOM) {
        const T* R = sm + OR + s4(2) * 3 * NC + tid + j * NT;
        h2[0][j] = rho_face(f2[j], f1[j], lr, omlr);
        h2[1][j] = rho_face(f2[j], sm[qF + e - 1], lr, omlr);
        h2[2][j] = rho_face(f2[j], sm[qF + e - WX], lr, omlr);
        // fused: ρu = u*ρ (u2ρu!, VOFutil.jl:208-211) and straight back to u★ = ρu/ρ, rounding as the two passes would
        const T ra = t_div(fused ? R[0] * h2[0][j] : R[0], h2[0][j]);
        const T rx = t_div(fused ? R[NC] * h2[1][j] : R[NC], h2[1][j]);
        const T rc = t_div(fused ? R[2 * NC] * h2[2][j] : R[2 * NC], h2[2][j]);
        us[0][j][3] = dq ? AA : ra;  // Dirichlet planes of BC!
        us[1][j][3] = dirX ? AXv : rx;
        us[2][j][3] = dirC[j] ? ACv : rc;
      }
      // C. VOF flux + mass flux through face k+2 (advection.jl:108-137)
      u2[j] = sm[qU + e]; u02[j] = SAMEU ? u2[j] : sm[qU0 + e];
      FFn[j] = T(0); Mn[j] = T(0); mkn[j] = false;
      if (needn) {
        T dl = P.hdt * (u2[j] + u02[j]);  // δt/2*(u+u⁰)
        dl = (dl != T(0)) ? dl : T(0);    // -0 -> +0: the zero-flux case of advection.jl:115 without a branch
        const bool up = dl > T(0);
        const T fc = up ? f1[j] : f2[j];  // upwind cell
        const bool gho = up ? ghU : ghD;
        if (dl != T(0) && !gho && !fullorempty(fc)) {
          sList[atomicAdd(&sCnt[(rel + 1) & 3], 1)] = e;  // interface face: reconstructed lane-dense at the top of the next step
          mkn[j] = true;
        } else {
          FFn[j] = fc * dl;
          Mn[j] = dl * lr + omlr * FFn[j];  // fᶠ2ρuf, VOFutil.jl:218
          if (MOM) Mn[j] = Mn[j] * P.idt;   // rmul!(ρuf, inv(δt)), flow.jl:207
        }
      }
      sm[wM + e] = Mn[j];
      // D. dilation of plane k+1 (flow.jl:216)
      const T div1 = (u2[j] - u1[j]) + (u02[j] - u01[j]);  // ∂(d,I,u)+∂(d,I,u⁰)
      if (first) cb1[j] = (f1[j] < T(0.5)) ? 0 : 1;  // flow.jl:172 (c̄ from the incoming f)
      dv1[j] = ((cb1[j] ? div1 : T(0)) * dt) / T(2);   // c̄[I]*(∂u+∂u⁰)*δt/2 of advection.jl:83 for the update of cell k+1
      dil1[j] = T(0);
      if (MOM) {
        dil1[j] = ((cb1[j] ? lam1 : lr) * div1) / T(2);
        sm[wD + e] = dil1[j];
      }
    }
    // C/D for the halo entry (mass flux and dilation that the x-1 / c-1 neighbours of the tile edge need)
    if (MOM && hU) {
      const unsigned pF = OF + sF8(1), pU = OU + s4(1) * PLH, pU0 = OU0 + s4(1) * PLH;
      const T uh2 = sm[qU + eh], u0h2 = SAMEU ? uh2 : sm[qU0 + eh], fh1 = sm[pF + eh];
      T m = T(0);
      if (needn) {
        T dl = P.hdt * (uh2 + u0h2);
        dl = (dl != T(0)) ? dl : T(0);
        const bool up = dl > T(0);
        const T fc = up ? fh1 : sm[qF + eh];
        const bool gho = up ? ghU : ghD;
        if (dl != T(0) && !gho && !fullorempty(fc)) sList[atomicAdd(&sCnt[(rel + 1) & 3], 1)] = eh;
        else {
          m = dl * lr + omlr * (fc * dl);
          m = m * P.idt;
        }
      }
      sm[wM + eh] = m;
      const T uh1 = sm[pU + eh];
      const T dh = (uh2 - uh1) + (u0h2 - (SAMEU ? uh1 : sm[pU0 + eh]));
      const int ch = first ? ((fh1 < T(0.5)) ? 0 : 1) : cbh1;
      sm[wD + eh] = ((ch ? lam1 : lr) * dh) / T(2);
    }

    const unsigned lk0 = lkU;
    lkU += sA;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int e = e0 + j * TR * WX;
      // G. SynDRoM momentum flux through face p = k+1 of the three momentum cells of this column (flow.jl:20-57,223)
      T Fhi[3] = {T(0), T(0), T(0)};
      if (MOM) {
        const T Mc = dp ? AA : Mhi[j];  // velocity BC! on ρuf (flow.jl:207)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          T Mo;
          if (r == 0) Mo = dpm ? AA : Mlo[j];
          else if (r == 1) Mo = dp ? AA : sm[rM + e - 1];
          else Mo = dp ? AA : sm[rM + e - WX];
          const T Psi = (Mc + Mo) / T(2);
          const T* u4 = us[r][j];
          const bool pos = Psi > T(0);
          T uu, cc, dd;
          if (Lvar) {  // ϕuL
            if (pos) { uu = T(2) * u4[1] - u4[2]; cc = u4[1]; dd = u4[2]; }
            else { uu = u4[3]; cc = u4[2]; dd = u4[1]; }
          } else if (Rvar) {  // ϕuR
            if (Psi < T(0)) { uu = T(2) * u4[2] - u4[1]; cc = u4[2]; dd = u4[1]; }
            else { uu = u4[0]; cc = u4[1]; dd = u4[2]; }
          } else {  // ϕu
            uu = pos ? u4[0] : u4[3];
            cc = pos ? u4[1] : u4[2];
            dd = pos ? u4[2] : u4[1];
          }
          // density of the donor momentum cell (plane k for Ψ>0, else k+1): linInterpProp of its face-centred old f
          // (dρ after f2face!+BCv!, flow.jl:205) -- the same ρ(f̄) that u★ was formed with
          T mOld = pos ? h0[r][j] : h1[r][j];
          if (r == 0) {
            if (Lvar && pos) mOld = h2[0][j];  // donor index 1: BCv! copies plane 3 = (f(3)+f(2))/2
            if (Rvar && !pos) mOld = lin_interp(__ldg(P.drho + (cA + (unsigned)(nA - 1) * sA + go[j])), lr, omlr);  // donor index nA: never written by f2face!
          }
          Fhi[r] = syndrom_flux_t<KOREN>(P.lim, Psi, uu, cc, dd, mOld, dt);
        }
      }
      // H. update of cell k
      if (store && valid[j]) {
        const unsigned lk = lk0 + go[j];
        if (first) P.cbar[lk] = (int8_t)((f0[j] < T(0.5)) ? 0 : 1);
        T fn = f0[j] + ((FFlo[j] - FFhi[j]) + dv0[j]);  // advection.jl:83
        rmax = max_nan(rmax, fn);
        rmin = t_min(rmin, fn);
        if (fn > T(1) || fn < T(0)) {  // only cells outside [0,1] can be reported (reportFillError, advection.jl:145-189)
          if (fn >= rmax) amax = lk;
          if (fn <= rmin) amin = lk;
        }
        fn = (fn < P.tol) ? T(0) : ((fn > P.onemtol) ? T(1) : fn);  // cleanWisp!
        P.f_out[lk] = fn;
        if (!MOM && P.rhouf_j != nullptr) {
          P.rhouf_j[lk] = Mlo[j];
          if (!FAST && k == nA - 1) P.rhouf_j[lk + sA] = Mhi[j];  // inside_uWB includes the upper boundary face
        }
        if (MOM) {
          const T* R = sm + OR + s4(0) * 3 * NC + tid + j * NT;
          const T* O = fused ? R : sm + OO + s2(0) * 3 * NC + tid + j * NT;
          const T dNa = (!FAST && !perA && k == 2) ? dil0[j] : dilm1[j];  // BCf! (Neumann) on ρ̄∂ⱼuⱼ along the sweep direction
          T qA = R[0], qX = R[NC], qC = R[2 * NC];  // ρu before the sweep
          if (fused) {  // u2ρu! + BC!(ρu,uBC): Dirichlet plane 2 of the normal component holds uBC
            qA = dpm ? AA : qA * h0[0][j];
            qX = dirX ? AXv : qX * h0[1][j];
            qC = dirC[j] ? ACv : qC * h0[2][j];
          }
          // r = Φ[I] - Φ[I+δj] + uOld*ϕ(i,I,ρ̄∂ⱼuⱼ);  ρu += δt*r          flow.jl:223-231
          const T rA = (Flo[0][j] - Fhi[0]) + O[0] * ((dil0[j] + dNa) / T(2));
          const T rX = (Flo[1][j] - Fhi[1]) + O[NC] * ((dil0[j] + sm[rD + e - 1]) / T(2));
          const T rC = (Flo[2][j] - Fhi[2]) + O[2 * NC] * ((dil0[j] + sm[rD + e - WX]) / T(2));
          P.rhou_out[lk + cA] = qA + dt * rA;
          P.rhou_out[lk] = qX + dt * rX;
          P.rhou_out[lk + cC] = qC + dt * rC;
        }
      }
      // I. roll the register pipeline
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        Flo[r][j] = Fhi[r]; h0[r][j] = h1[r][j]; h1[r][j] = h2[r][j];
#pragma unroll
        for (int i = 0; i < 3; ++i) us[r][j][i] = us[r][j][i + 1];
      }
      FFlo[j] = FFhi[j]; Mlo[j] = Mhi[j]; FFhi[j] = FFn[j]; Mhi[j] = Mn[j]; mk[j] = mkn[j];
      dilm1[j] = dil0[j]; dil0[j] = dil1[j]; dv0[j] = dv1[j];
      f0[j] = f1[j]; f1[j] = f2[j]; u1[j] = u2[j]; u01[j] = u02[j];
    }
  };

  // ---- march: groups of four FAST steps wherever every plane a step touches is a plain interior plane ----------------------------
  {
    int k = ks;
    while (k < k1) {
      if (((k - ks) & 3) == 0 && k >= 3 && k + 3 <= nA - 4 && k + 3 < k1) {
        step(BoolC<true>{}, IntC<0>{}, k);
        step(BoolC<true>{}, IntC<1>{}, k + 1);
        step(BoolC<true>{}, IntC<2>{}, k + 2);
        step(BoolC<true>{}, IntC<3>{}, k + 3);
        k += 4;
        const unsigned t = fLo; fLo = fHi; fHi = t;
      } else {
        step(BoolC<false>{}, IntC<0>{}, k);
        ++k;
        if (((k - ks) & 3) == 0) { const unsigned t = fLo; fLo = fHi; fHi = t; }
      }
    }
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
      red_commit<T>(P.red, rmax, rmin, amax, amin, rnan);
    }
  }
}

}  // namespace ifadv
