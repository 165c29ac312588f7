// ifadv_march.cuh -- the fused directional sweep for 3-D grids (sm_100a): plane marching with an asynchronous,
// double-buffered shared-memory pipeline.
//
// A CTA owns a tile of TA x TB cells in the plane spanned by the sweep direction a = J and one cross direction b,
// and marches through a chunk of planes along the remaining direction c.  x (the contiguous dimension) is always
// one of the two in-plane dimensions, so every global access is a coalesced row segment:
//     J = 0 : a = x, b = y, march c = z        J = 1 : a = y, b = x, march c = z        J = 2 : a = z, b = x, march c = y
//
// Every shared plane has the same shape, (TA+5) x (TB+1) entries (halo -3..+2 along a, -1 along b), and thread t owns
// entries t, t+NT, t+2NT.. of it.  Everything that depends only on the entry -- the global offsets with the ghost
// rules of BCf!/BC! folded in (clamp = Neumann, wrap = periodic), and a bit mask of which stage regions and which
// one-sided / Dirichlet rules apply -- is computed ONCE before the march, so the per-plane work is pure data flow:
//   wait   cp.async copies of plane k (issued one plane earlier; f two planes earlier) have landed in shared memory
//   issue  cp.async copies of f(k+2) and u_a, u⁰_a, ρu (3 components) of plane k+1 -> other ring slots (HBM latency hidden
//          behind the arithmetic of plane k; the row pitch 4(N+2) B is not a multiple of 16 B, so TMA tiled copies
//          cannot be used on the reference's array layout -- LDGSTS 4/8-byte copies can)
//   P2     VOF face flux fᶠ + mass flux (advection.jl:108-137), dilation ρ̄∂ⱼuⱼ (flow.jl:216), u★ = ρu/ρ (flow.jl:197);
//          faces whose upwind cell holds an interface are only MARKED here and then reconstructed lane-dense from the
//          shared f planes (3^3 box: planes k-1,k,k+1) -- a divergent in-line PLIC would cost a full warp per interface cell
//   P3     SynDRoM momentum flux through every a-face of the tile, once per face (flow.jl:20-57,223)
//   P4     cell update: f (advection.jl:83, cleanWisp!), ρu (flow.jl:224-231), fill-error extrema
// Quantities of the previous plane that the stencil reaches along c (f, mass flux, dilation) stay in shared-memory
// rings; none of the reference's intermediates (u★, Φ, n̂, α, fᶠ, ρuf, dρ, ρ̄∂ⱼuⱼ, r) ever reaches HBM.
#pragma once
#include "ifadv_sweep.cuh"

namespace ifadv {

template <int J, int TA, int TB, int NT> struct MTile {
  static constexpr bool AX = (J == 0);  // is the sweep direction the contiguous one?
  // halo: -3..+2 along a (u★ line + upwind cell), -2..+1 along b (3^3 PLIC box of a cell at b = -1 / TB-1)
  static constexpr int HAm = 3, HAp = 2, HBm = 2, HBp = 1;
  static constexpr int WA = TA + HAm + HAp, WB = TB + HBm + HBp;
  static constexpr int PL = WA * WB;                      // entries of one shared plane
  static constexpr int SA = AX ? 1 : WB, SB = AX ? WA : 1;  // x fastest; entry index == shared index
  // Entry enumeration: in the main trips a warp owns whole rows of the 32 tile columns along x (all lanes active in
  // every stage); the few halo columns are enumerated separately in one extra trip, so no stage pays for idle lanes.
  static constexpr int NW = NT / 32;
  static constexpr int ROWS = AX ? WB : WA;                 // rows = the non-contiguous in-plane direction
  static constexpr int WX = AX ? WA : WB;                   // row length (x)
  static constexpr int HXm = AX ? HAm : HBm;                // halo columns left of the tile
  static constexpr int NHC = WX - 32;                       // halo columns per row
  static constexpr int TM = (ROWS + NW - 1) / NW;           // main trips
  static constexpr int NH = ROWS * NHC;                     // halo entries
  static constexpr int TH = (NH + NT - 1) / NT;             // halo trips
  static constexpr int TRIPS = TM + TH;
  static constexpr int ESTR = NW * WX;                      // entry stride between main trips
  static_assert((AX ? TA : TB) == 32, "the tile must be 32 cells wide along x");
  // CMOM planes: F x4, U x2, U0 x2, M x2, FF, RU x6, Us x3, Dil x2, Fl x3 = 25 ; pure VOF: F x4, U x2, U0 x2, M, FF, list = 11
  static constexpr int NPLANES_MOM = 25, NPLANES_VOF = 11;
  template <class T> static constexpr size_t smem_bytes(bool mom) { return sizeof(T) * (size_t)PL * (mom ? NPLANES_MOM : NPLANES_VOF) + 16; }
};

enum : unsigned {
  MF_U = 1u << 0,       // entry is a face of the velocity / flux region      la in [-1,TA], lb >= -1
  MF_NEEDM = 1u << 1,   // VOF face flux has to be evaluated here
  MF_DIL = 1u << 2,     // dilation region                                     la in [-1,TA-1]
  MF_US = 1u << 3,      // u★ region                                           la in [-2,TA+1], lb >= 0
  MF_FL = 1u << 4,      // momentum-flux faces                                 la in [0,TA],    lb >= 0
  MF_CELL = 1u << 5,    // owned cell (inside the domain)                      la in [0,TA-1],  lb >= 0
  MF_DIRA = 1u << 6,    // Dirichlet plane of component a at this index (BC!)  va in {1,2,nA}, a not periodic
  MF_DIRAM = 1u << 7,   // ... at va-1
  MF_DIRB = 1u << 8,    // Dirichlet plane of component b                      vb in {2,nB}, b not periodic
  MF_DILSH = 1u << 9,   // Neumann ghost of ρ̄∂ⱼuⱼ along a: use the cell at la+1 (va == 1)
  MF_LVAR = 1u << 10,   // ϕuL face (va == 2, a not periodic)
  MF_RVAR = 1u << 11,   // ϕuR face (va == nA, a not periodic)
  MF_GHLO = 1u << 12,   // cell la-1 is a ghost cell on a non-periodic side (no PLIC reconstruction)
  MF_GHHI = 1u << 13,   // cell la itself is one
  MF_TOPA = 1u << 14,
  MF_VALID = 1u << 15,  // entry exists (trip slot in use)   // owned cell with va == nA-1 (writes the boundary face of ρuf in the pure-VOF path)
};

template <class T> IFADV_DI void cp_async(T* smem_dst, const T* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gsrc));
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
template <class T> IFADV_DI void cp_async_s(unsigned saddr, const T* gsrc) {
  if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(saddr), "l"(gsrc));
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(gsrc));
}
// index maps for (block-uniform) plane indices; the warm-up of a chunk may reach several periods below a tiny periodic box
IFADV_DI int wrap1(int v, int n) { return wrapc(v, n); }
IFADV_DI int map1(int v, int n, bool per) { return per ? wrap1(v, n) : min(max(v, 2), n - 1); }
IFADV_DI int own1(int v, int n, bool per) { return per ? wrap1(v, n) : min(max(v, 1), n); }

IFADV_DI void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
IFADV_DI void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// 3^3 box accessor of the normal schemes reading the shared f planes: (dx,dy,dz) in global axes -> (a,b,c) offsets
template <class T, int J, int PL, int SA, int SB> struct SBox {
  const T* sF;  // 4-slot ring of f planes
  int e;        // entry of the box centre
  int vc;       // plane of the box centre
  IFADV_DI T operator()(int dx, int dy, int dz) const {
    const int da = (J == 0) ? dx : ((J == 1) ? dy : dz);
    const int db = (J == 0) ? dy : dx;
    const int dc = (J == 2) ? dy : dz;
    return sF[((vc + dc) & 3) * PL + e + da * SA + db * SB];
  }
};

template <class T, int J, int TA, int TB, bool MOM, bool FUSED, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) march_kernel(const SweepP<T> P, const int chunk) {
  using TL = MTile<J, TA, TB, NT>;
  constexpr bool AX = TL::AX;
  constexpr int DB = (J == 0) ? 1 : 0;  // global dimension of b
  constexpr int DC = (J == 2) ? 1 : 2;  // global dimension of the march direction c
  constexpr int PL = TL::PL, SA = TL::SA, SB = TL::SB, TR = TL::TRIPS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);

  const Geo& g = P.g;
  const int tid = threadIdx.x;
  const int nA = g.n[J], nB = g.n[DB], nC = g.n[DC];
  const long long st3[3] = {1, g.s1, g.s2};
  const long long sA = st3[J], sB = st3[DB], sC = st3[DC];
  const bool perA = (g.per >> J) & 1u, perB = (g.per >> DB) & 1u, perC = (g.per >> DC) & 1u;
  const long long cA = P.coff[J], cB = P.coff[DB], cC = P.coff[DC];

  // blockIdx.x tiles x, blockIdx.y tiles the other in-plane dimension, blockIdx.z chunks of the march direction
  const int ox = 2 + blockIdx.x * (AX ? TA : TB), oo = 2 + blockIdx.y * (AX ? TB : TA);
  const int oa = AX ? ox : oo, ob = AX ? oo : ox;
  const int k0 = 2 + blockIdx.z * chunk, k1 = min(k0 + chunk, nC);

  // shared planes
  T* sF = sm;                                   // 4 slots: plane k&3 (loaded two planes ahead: the PLIC box reaches k+1)
  T* sU = sm + 4 * PL;                          // 2 slots: plane k&1
  T* sU0 = sm + 6 * PL;                         // 2 slots
  T* sM = sm + 8 * PL;                          // CMOM: 2 slots; pure VOF: 1
  T* sFF = MOM ? sm + 10 * PL : sm + 9 * PL;
  T* sRU = sm + 11 * PL;                        // [slot][role] : (k&1)*3 + r
  T* sUs = sm + 17 * PL;                        // [role]
  T* sDil = sm + 20 * PL;                       // 2 slots
  T* sFl = sm + 22 * PL;                        // [role]
  // interface faces of the current plane, compacted so that the expensive reconstruction runs lane-dense;
  // the list shares storage with the momentum-flux planes (CMOM) that are written only after it is consumed
  int* sList = reinterpret_cast<int*>(MOM ? sFl : sm + 10 * PL);
  int* sCnt = reinterpret_cast<int*>(sm + (size_t)PL * (MOM ? TL::NPLANES_MOM : TL::NPLANES_VOF));
  if (tid == 0) *sCnt = 0;

  // ---- per-entry constants (hoisted out of the march) ----------------------------------------------------------------------
  unsigned flg[TR];
  int gmm[TR];  // offset within a c-plane, both in-plane indices mapped to the interior (f, c̄, tangential components)
  int gom[TR];  // a as stored (component a / u_a faces), b mapped
  const int lane = tid & 31, warp = tid >> 5;
  const int e0 = warp * TL::WX + TL::HXm + lane;  // main-trip entry of trip 0; trip t adds t*ESTR
  int eh[TL::TH > 0 ? TL::TH : 1];
#pragma unroll
  for (int h = 0; h < TL::TH; ++h) {
    const int q = tid + h * NT, row = q / TL::NHC, c = q % TL::NHC;
    eh[h] = row * TL::WX + (c < TL::HXm ? c : c + 32);
  }
  auto ent = [&](int t) -> int { return (t < TL::TM) ? e0 + t * TL::ESTR : eh[(t < TL::TM) ? 0 : t - TL::TM]; };
#pragma unroll
  for (int t = 0; t < TR; ++t) {
    const int e = ent(t);
    const bool valid = (t < TL::TM) ? (warp + t * TL::NW < TL::ROWS) : (tid + (t - TL::TM) * NT < TL::NH);
    const int ia = AX ? e % TL::WA : e / TL::WB, ib = AX ? e / TL::WA : e % TL::WB;
    const int la = ia - TL::HAm, lb = ib - TL::HBm;
    const int va = oa + la, vb = ob + lb;
    unsigned f = 0;
    if (valid) {
      f |= MF_VALID;
      const bool inB = lb >= -1 && lb <= TB - 1;  // b-range of every computed region (lb = -2 and lb = TB only feed the PLIC box)
      const bool inU = inB && la >= -1 && la <= TA;
      if (inU) f |= MF_U;
      if (inU && !(la < 0 && lb < 0) && va <= nA && (perA || va >= 2)) f |= MF_NEEDM;
      if (inB && la >= -1 && la <= TA - 1) f |= MF_DIL;
      if (inB && la >= -2 && la <= TA + 1 && lb >= 0) f |= MF_US;
      if (inB && la >= 0 && la <= TA && lb >= 0 && va <= nA) f |= MF_FL;
      if (inB && la >= 0 && la <= TA - 1 && lb >= 0 && va <= nA - 1 && vb <= nB - 1) f |= MF_CELL;
      if (!perA) {
        if (va == 1 || va == 2 || va == nA) f |= MF_DIRA;
        if (va - 1 == 1 || va - 1 == 2 || va - 1 == nA) f |= MF_DIRAM;
        if (va == 1) f |= MF_DILSH;
        if (va == 2) f |= MF_LVAR;
        if (va == nA) f |= MF_RVAR;
        if (va - 1 < 2 || va - 1 > nA - 1) f |= MF_GHLO;
        if (va < 2 || va > nA - 1) f |= MF_GHHI;
      }
      if (!perB && (vb == 2 || vb == nB)) f |= MF_DIRB;
      if (va == nA - 1) f |= MF_TOPA;
    }
    flg[t] = f;
    const int ma = mapc(va, nA, perA), mb = mapc(vb, nB, perB);
    const int wa = perA ? wrapc(va, nA) : min(max(va, 1), nA);
    gmm[t] = (int)((ma - 1) * sA + (mb - 1) * sB);
    gom[t] = (int)((wa - 1) * sA + (mb - 1) * sB);  // (component b "as stored" == mapped: its entries all have vb interior)
  }

  const T lr = P.lr, omlr = P.omlr, dt = P.dt;
  const T AA = P.A[J], AB = P.A[DB], AC = P.A[DC];

  // ---- asynchronous plane loads ---------------------------------------------------------------------------------------------------
  constexpr unsigned SZ = sizeof(T);
  const unsigned sb = (unsigned)__cvta_generic_to_shared(sm);
  auto issue_f = [&](int vc) {
    const T* fp = P.f_in + (long long)(map1(vc, nC, perC) - 1) * sC;
    const unsigned dF = sb + (unsigned)(vc & 3) * (PL * SZ);
#pragma unroll
    for (int t = 0; t < TR; ++t) {
      if (flg[t] & MF_VALID) cp_async_s(dF + ent(t) * SZ, fp + gmm[t]);
    }
  };
  auto issue_rest = [&](int vc, bool full) {
    const long long pm = (long long)(map1(vc, nC, perC) - 1) * sC;
    const long long po = (long long)(own1(vc, nC, perC) - 1) * sC;
    const unsigned s2 = (unsigned)(vc & 1);
    const T* up = P.uj + pm;
    const T* u0p = P.u0j + pm;
    const unsigned dU = sb + (unsigned)((sU - sm) + s2 * PL) * SZ;
    const unsigned dU0 = sb + (unsigned)((sU0 - sm) + s2 * PL) * SZ;
    const unsigned dR = sb + (unsigned)((sRU - sm) + s2 * 3 * PL) * SZ;
    const T* rsrc = (MOM && FUSED) ? P.uOld : P.rhou_in;  // fused sweep 1: the ρu planes carry uOld
    const T* ra = rsrc + cA + pm;
    const T* rb = rsrc + cB + pm;
    const T* rc = rsrc + cC + po;
#pragma unroll
    for (int t = 0; t < TR; ++t) {
      const unsigned e4 = ent(t) * SZ;
      if (flg[t] & MF_U) {
        cp_async_s(dU + e4, up + gom[t]);
        cp_async_s(dU0 + e4, u0p + gom[t]);
      }
      if (MOM && full && !FUSED && (flg[t] & MF_CELL)) {  // uOld of the next plane: pull the lines into L2/L1 ahead of the update stage
        const T* uo = P.uOld + (long long)(vc - 1) * sC + gmm[t];
        asm volatile("prefetch.global.L1 [%0];" ::"l"(uo + cA));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(uo + cB));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(uo + cC));
      }
      if (MOM && full && (flg[t] & MF_US)) {
        cp_async_s(dR + e4, ra + gom[t]);
        cp_async_s(dR + e4 + PL * SZ, rb + gmm[t]);
        cp_async_s(dR + e4 + 2 * PL * SZ, rc + gmm[t]);
      }
    }
  };

  // c̄ of the entries of plane vc (non-first sweeps: read one plane ahead into registers)
  int cbn[TR];
  auto load_cbar = [&](int vc) {
    if (MOM && !P.first) {
      const long long pm = (long long)(map1(vc, nC, perC) - 1) * sC;
#pragma unroll
      for (int t = 0; t < TR; ++t) cbn[t] = (flg[t] & MF_DIL) ? (int)P.cbar[pm + gmm[t]] : 0;
    }
  };

  // ---- P2a: VOF face flux + mass flux (+ dilation) of plane vc ---------------------------------------------------------------------
  int cbk[TR];
  auto flux_stage = [&](int vc) {
    const int s2 = vc & 1;
    const T* cF = sF + (vc & 3) * PL;
    const T* cU = sU + s2 * PL;
    const T* cU0 = sU0 + s2 * PL;
    T* cM = sM + (MOM ? s2 * PL : 0);
#pragma unroll
    for (int t = 0; t < TR; ++t) {
      const int e = ent(t);
      const unsigned fl = flg[t];
      if (fl & MF_U) {
        T ff = T(0), m = T(0);
        if (fl & MF_NEEDM) {
          const T dl = P.hdt * (cU[e] + cU0[e]);  // δt/2*(u+u⁰), advection.jl:110
          if (dl != T(0)) {                       // advection.jl:115
            const bool up = dl > T(0);            // upwind cell la-1, advection.jl:120
            const T fc = cF[up ? e - SA : e];
            const bool ghost = (fl & (up ? MF_GHLO : MF_GHHI)) != 0;
            if (ghost || fullorempty(fc)) {       // advection.jl:125-126
              ff = fc * dl;
              m = dl * lr + omlr * ff;            // fᶠ2ρuf, VOFutil.jl:218
              if (MOM) m = m * P.idt;             // rmul!(ρuf, inv(δt)), flow.jl:207
            } else sList[atomicAdd(sCnt, 1)] = e;  // interface face: reconstructed lane-dense in plic_stage
          }
        }
        sFF[e] = ff;
        cM[e] = m;
      }
      if (MOM && (fl & MF_DIL)) {
        const int e2 = (fl & MF_DILSH) ? e + SA : e;  // BCf! (Neumann) on ρ̄∂ⱼuⱼ along the sweep direction, flow.jl:217
        const T div = (cU[e2 + SA] - cU[e2]) + (cU0[e2 + SA] - cU0[e2]);  // ∂(d,I,u)+∂(d,I,u⁰)
        const int cb = P.first ? ((cF[e] < T(0.5)) ? 0 : 1) : cbn[t];  // flow.jl:172 (c̄ from the incoming f)
        cbk[t] = cb;
        sDil[s2 * PL + e] = (lin_interp(T(cb), lr, omlr) * div) / T(2);  // flow.jl:216
      }
    }
  };
  // PLIC reconstruction + flux of the compacted interface faces (general branch of getVOFFlux!, advection.jl:131-134)
  auto plic_stage = [&](int vc, int cnt) {
    const int s2 = vc & 1;
    const T* cF = sF + (vc & 3) * PL;
    const T* cU = sU + s2 * PL;
    const T* cU0 = sU0 + s2 * PL;
    T* cM = sM + (MOM ? s2 * PL : 0);
    for (int i = tid; i < cnt; i += NT) {
      const int e = sList[i];
      const T dl = P.hdt * (cU[e] + cU0[e]);
      const int eu = (dl > T(0)) ? e - SA : e;
      SBox<T, J, PL, SA, SB> B{sF, eu, vc};
      const T ff = plic_face_flux<T, 3>(P.scheme, B, cF[eu], J, dl);
      T m = dl * lr + omlr * ff;
      if (MOM) m = m * P.idt;
      sFF[e] = ff;
      cM[e] = m;
    }
  };

  // ---- P2b: u★ = BC!(ρu/ρ(f̄)) of plane vc (flow.jl:197, VOFutil.jl:198-201) -----------------------------------------------------------
  auto ustar_stage = [&](int vc) {
    const int s2 = vc & 1;
    const T* cF = sF + (vc & 3) * PL;
    const T* pF = sF + ((vc - 1) & 3) * PL;
    const T* cR = sRU + (s2 * 3) * PL;
    const bool dirC = !perC && (vc == 2 || vc == nC);
#pragma unroll
    for (int t = 0; t < TR; ++t) {
      const int e = ent(t);
      const unsigned fl = flg[t];
      if (fl & MF_US) {
        const T fc = cF[e];
        const T ha = rho_face(fc, cF[e - SA], lr, omlr), hb = rho_face(fc, cF[e - SB], lr, omlr),
                hc = rho_face(fc, pF[e], lr, omlr);
        constexpr bool fu = MOM && FUSED;  // ρu = u*ρ (u2ρu!) formed on the fly, then u★ = ρu/ρ
        const T ra = t_div(fu ? cR[e] * ha : cR[e], ha);
        const T rb = t_div(fu ? cR[PL + e] * hb : cR[PL + e], hb);
        const T rc = t_div(fu ? cR[2 * PL + e] * hc : cR[2 * PL + e], hc);
        sUs[e] = (fl & MF_DIRA) ? AA : ra;  // Dirichlet planes of BC!
        sUs[PL + e] = (fl & MF_DIRB) ? AB : rb;
        sUs[2 * PL + e] = dirC ? AC : rc;
      }
    }
  };

  // ---- P3: SynDRoM momentum flux through the lower a-face of every momentum cell of the tile -----------------------------------------------
  auto mom_flux_stage = [&](int vc) {
    const int s2 = vc & 1;
    const T* cF = sF + (vc & 3) * PL;
    const T* pF = sF + ((vc - 1) & 3) * PL;
    const T* cM = sM + s2 * PL;
    const T* pM = sM + (s2 ^ 1) * PL;
#pragma unroll
    for (int t = 0; t < TR; ++t) {
      const int e = ent(t);
      const unsigned fl = flg[t];
      if (fl & MF_FL) {
        // BC-aware mass flux (velocity BC! on ρuf: Dirichlet planes of component a), flow.jl:207
        const bool dir0 = fl & MF_DIRA;
        const T Mc = dir0 ? AA : cM[e];
        const bool Lvar = fl & MF_LVAR, Rvar = fl & MF_RVAR;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          T Mo;  // second sample of Ψ = ϕ(i,CI(I,j),ρuf)
          if (r == 0) Mo = (fl & MF_DIRAM) ? AA : cM[e - SA];
          else if (r == 1) Mo = dir0 ? AA : cM[e - SB];
          else Mo = dir0 ? AA : pM[e];
          const T Psi = (Mc + Mo) / T(2);
          const T* su = sUs + r * PL;
          const T um1 = su[e - SA], uc = su[e];
          const bool pos = Psi > T(0);
          T uu, cc, dd;
          if (Lvar) {  // ϕuL, flow.jl:28-31
            if (pos) { uu = T(2) * um1 - uc; cc = um1; dd = uc; }
            else { uu = su[e + SA]; cc = uc; dd = um1; }
          } else if (Rvar) {  // ϕuR, flow.jl:32-35
            if (Psi < T(0)) { uu = T(2) * uc - um1; cc = uc; dd = um1; }
            else { uu = su[e - 2 * SA]; cc = um1; dd = uc; }
          } else {  // ϕu, flow.jl:20-23 (ϕuP on a periodic boundary is the same stencil through the wrap)
            uu = pos ? su[e - 2 * SA] : su[e + SA];
            cc = pos ? um1 : uc;
            dd = pos ? uc : um1;
          }
          // donor momentum cell (flow.jl:42) and its face-centred old f (dρ after f2face!+BCv!, flow.jl:205)
          int ed = pos ? e - SA : e;
          T fo;
          if (r == 0) {
            if (Lvar && pos) ed += 2 * SA;  // donor index 1 on a non-periodic side: BCv! copies plane 3
            fo = (cF[ed] + cF[ed - SA]) / T(2);
            if (Rvar && !pos) {             // donor index nA: the plane f2face! never writes
              const int ib = AX ? e / TL::WA : e % TL::WB;
              fo = __ldg(P.drho + cA + (long long)(nA - 1) * sA + (long long)(mapc(ob + ib - TL::HBm, nB, perB) - 1) * sB +
                         (long long)(mapc(vc, nC, perC) - 1) * sC);
            }
          } else if (r == 1) fo = (cF[ed] + cF[ed - SB]) / T(2);
          else fo = (cF[ed] + pF[ed]) / T(2);
          sFl[r * PL + e] = syndrom_flux(P.lim, Psi, uu, cc, dd, lin_interp(fo, lr, omlr), dt);
        }
      }
    }
  };

  // ---- P4: cell update -------------------------------------------------------------------------------------------------------------------------
  T rmax = -INFINITY, rmin = INFINITY;
  unsigned int amax = 0, amin = 0;
  int rnan = 0;
  auto update_stage = [&](int vc) {
    const int s2 = vc & 1;
    const T* cF = sF + (vc & 3) * PL;
    const T* cU = sU + s2 * PL;
    const T* cU0 = sU0 + s2 * PL;
    const T* cM = sM + (MOM ? s2 * PL : 0);
    const long long pc = (long long)(vc - 1) * sC;
#pragma unroll
    for (int t = 0; t < TR; ++t) {
      const int e = ent(t);
      const unsigned fl = flg[t];
      if (fl & MF_CELL) {
        const long long lk = pc + gmm[t];  // owned cells are interior: the mapped offset is the cell itself
        const T fK = cF[e];
        int cb;
        if (MOM) cb = cbk[t];
        else cb = P.first ? ((fK < T(0.5)) ? 0 : 1) : (int)P.cbar[lk];
        const T div = (cU[e + SA] - cU[e]) + (cU0[e + SA] - cU0[e]);
        if (P.first) P.cbar[lk] = (int8_t)cb;
        // f[I] += fᶠ[I]-fᶠ[I+δ] + c̄[I]*(∂u+∂u⁰)*δt/2          advection.jl:83
        T fn = fK + ((sFF[e] - sFF[e + SA]) + ((T(cb) * div) * dt) / T(2));
        if (fn != fn) rnan = 1;
        if (fn > rmax) { rmax = fn; amax = (unsigned int)lk; }
        if (fn < rmin) { rmin = fn; amin = (unsigned int)lk; }
        fn = (fn < P.tol) ? T(0) : ((fn > P.onemtol) ? T(1) : fn);  // cleanWisp!
        P.f_out[lk] = fn;
        if (!MOM && P.rhouf_j != nullptr) {
          P.rhouf_j[lk] = cM[e];
          if (fl & MF_TOPA) P.rhouf_j[lk + sA] = cM[e + SA];  // inside_uWB includes the upper boundary face
        }
        if (MOM) {
          const T dK = sDil[s2 * PL + e];
          const T* cR = sRU + (s2 * 3) * PL;
          const long long co[3] = {cA, cB, cC};
          constexpr bool fu = MOM && FUSED;
          const T* pF = sF + ((vc - 1) & 3) * PL;
          const bool dirCc = !perC && (vc == 2 || vc == nC);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            T dN;
            if (r == 0) dN = sDil[s2 * PL + e - SA];
            else if (r == 1) dN = sDil[s2 * PL + e - SB];
            else dN = sDil[(s2 ^ 1) * PL + e];
            const T src = cR[r * PL + e];  // ρu before the sweep, or uOld in the fused first sweep
            T q = src, uo;
            if (fu) {  // u2ρu! + BC!(ρu,uBC): Dirichlet plane 2 of the normal component holds uBC
              const T fn2 = (r == 0) ? cF[e - SA] : ((r == 1) ? cF[e - SB] : pF[e]);
              const bool dir = (r == 0) ? ((fl & MF_LVAR) != 0) : ((r == 1) ? ((fl & MF_DIRB) != 0) : dirCc);
              const T Ar = (r == 0) ? AA : ((r == 1) ? AB : AC);
              q = dir ? Ar : src * rho_face(fK, fn2, lr, omlr);
              uo = src;
            } else uo = __ldg(P.uOld + co[r] + lk);
            // r = Φ[I] - Φ[I+δj] + uOld*ϕ(i,I,ρ̄∂ⱼuⱼ);  ρu += δt*r          flow.jl:223-231
            const T rr = (sFl[r * PL + e] - sFl[r * PL + e + SA]) + uo * ((dK + dN) / T(2));
            P.rhou_out[co[r] + lk] = q + dt * rr;
          }
        }
      }
    }
  };

  // ---- march -----------------------------------------------------------------------------------------------------------------------------------
  // one directional VOF flux pass over plane vc: mark + (if any interface face) lane-dense reconstruction
  auto vof_flux = [&](int vc) {
    flux_stage(vc);
  };
  if (MOM) {
    // prologue: plane k0-1 supplies f, mass flux and dilation of the previous plane (its PLIC box reaches k0-2 .. k0)
    issue_f(k0 - 2); issue_f(k0 - 1); issue_f(k0);
    issue_rest(k0 - 1, false);
    cp_async_commit();
    load_cbar(k0 - 1);
    cp_async_wait_all();
    __syncthreads();
    issue_f(k0 + 1);
    issue_rest(k0, true);
    cp_async_commit();
    vof_flux(k0 - 1);
    __syncthreads();
    {
      const int cnt = *sCnt;
      if (cnt > 0) plic_stage(k0 - 1, cnt);
    }
    load_cbar(k0);
    __syncthreads();
    if (tid == 0) *sCnt = 0;
  } else {
    issue_f(k0 - 1); issue_f(k0); issue_f(k0 + 1);
    issue_rest(k0, true);
    cp_async_commit();
  }
  for (int k = k0; k < k1; ++k) {
    cp_async_wait_all();
    __syncthreads();  // S1: plane k (and f of plane k+1) landed for every thread
    if (k + 1 < k1) {
      issue_f(k + 2);
      issue_rest(k + 1, true);
      cp_async_commit();
    }
    vof_flux(k);
    if (MOM) {
      if (k + 1 < k1) load_cbar(k + 1);
      ustar_stage(k);
    }
    __syncthreads();  // S2
    {
      const int cnt = *sCnt;  // block-uniform
      if (cnt > 0) {
        plic_stage(k, cnt);
        __syncthreads();
      }
    }
    if (MOM) {
      mom_flux_stage(k);
      __syncthreads();  // S3
    }
    update_stage(k);
    if (tid == 0) *sCnt = 0;
  }

  // ---- fill-error reduction: warp shuffles, then one atomic per warp (replaces findmax/findmin + host sync) -----------------------------------
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
