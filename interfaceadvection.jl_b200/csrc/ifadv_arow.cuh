// ifadv_arow.cuh -- fused directional sweep along y or z (J = 1, 2) for 3-D grids: WARP-AUTONOMOUS columns, register marching.
//
// The counterpart of ifadv_xrow.cuh for the two sweep directions that are NOT the contiguous one.  A WARP owns one strip of
// 62 (x) cells of one cross-direction column and marches ALONG the sweep direction a through a chunk of planes on its own: no CTA
// barrier anywhere in the march, no shared planes for neighbour exchange.
//   * a thread owns TWO x-adjacent cells (a, b): every global access is one aligned 8-byte (Float32) / 16-byte (Float64) vector;
//     lanes 1..31 produce output (62 cells), lane 0 only supplies the x-1 neighbour of lane 1 (tiles overlap by two cells);
//   * what the stencil reaches ALONG the sweep -- the 4-point u★ line of each component, the VOF / mass / momentum flux through
//     the lower face, the previous dilation, the face densities ρ(f̄) -- rolls through REGISTERS exactly as in ifadv_along2.cuh;
//   * the x-1 neighbour's mass flux and dilation come by __shfl_up (inside a thread cell a is cell b's x-1); the c-1 neighbour's are
//     RECOMPUTED: the warp evaluates the VOF face flux and the dilation of its own column and of the column c-1;
//   * f and ρu pass through warp-private shared rings filled with cp.async (f: 8 planes x 4 columns for the 3^3 PLIC boxes;
//     ρu: 4 planes, read twice -- for u★ of plane k+2 and for the update of plane k); u_a, uOld and c̄ go straight to registers;
//   * faces whose upwind cell holds an interface are marked, compacted with a ballot and reconstructed lane-dense by the warp
//     before the step continues (general branch of getVOFFlux!, advection.jl:131-134).
// Step k: u★ of plane k+2 (flow.jl:197), VOF flux + mass flux of face k+1 and dilation of plane k (advection.jl:108-137, flow.jl:216),
// SynDRoM fluxes of face k+1 (flow.jl:20-57,223), update of cell k (advection.jl:83, cleanWisp!, flow.jl:224-231).  Arithmetic
// (expression by expression) and boundary rules are those of ifadv_along2.cuh; reference lines cited there.
#pragma once
#include "ifadv_xrow.cuh"

namespace ifadv {

struct ARTile {
  static constexpr int NW = 8;          // warps per CTA = cross-direction columns per CTA (independent of each other)
  static constexpr int TX = 62;         // cells a warp owns along x: lanes 1..31, two each
  static constexpr int FC = 4;          // columns of the warp's f ring: c-2 .. c+1
  static constexpr int FP = 64;         // row pitch: a0, b0, ..., a31, b31
  static constexpr int PLF = FC * FP;   // one f plane
  static constexpr int NF = 128;        // faces a warp evaluates per step (own column + column c-1)
  // per warp: f ring x8, ρu ring 4 planes x 3 components, fᶠ and mass flux of reconstructed faces (2 columns), δl of listed faces, the list
  static constexpr int WELEMS = 8 * PLF + 4 * 3 * FP + 2 * 2 * FP + NF;
  template <class T> struct Bytes {
    static constexpr size_t warp = (sizeof(T) * (size_t)WELEMS + sizeof(int) * (size_t)NF + 15) / 16 * 16;
    static constexpr size_t cta = NW * warp;
  };
};

template <class T, int J> struct ARBox {  // 3^3 box on the warp's f ring: x fastest, columns along c, ring along the sweep direction
  const T* sF;
  int e, k;  // entry (column * FP + x) of the box centre within a plane, plane index
  IFADV_DI T operator()(int dx, int dy, int dz) const {
    const int da = (J == 1) ? dy : dz, dc = (J == 1) ? dz : dy;
    return sF[((k + da) & 7) * ARTile::PLF + e + dx + dc * ARTile::FP];
  }
};

template <class T, int J, bool MOM, bool FUSED, bool KOREN, bool SAMEU, bool EDGE>
IFADV_DI void arow_body(const SweepP<T>& P, const int chunk, unsigned char* smem_raw) {
  using TL = ARTile;
  using T2 = typename V2<T>::type;
  static_assert(J == 1 || J == 2, "sweeps along x use ifadv_xrow.cuh");
  constexpr int DCC = (J == 1) ? 2 : 1;  // global dimension of the cross direction c
  constexpr int FP = TL::FP, PLF = TL::PLF;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned SZ = sizeof(T);
  constexpr bool fused = MOM && FUSED;
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  T* sF = reinterpret_cast<T*>(smem_raw + (size_t)wq * TL::template Bytes<T>::warp);  // [8][4][FP]
  T* sR = sF + 8 * PLF;                                                               // [4][3][FP]  ρu ring: roles A, X, C
  T* sFX = sR + 4 * 3 * FP;                                                           // [2][FP]  fᶠ of reconstructed faces (columns c-1, c)
  T* sMX = sFX + 2 * FP;                                                              // [2][FP]  their mass flux
  T* sDl = sMX + 2 * FP;                                                              // [NF] δl of the listed faces
  int* sList = reinterpret_cast<int*>(sDl + TL::NF);                                  // [NF] face entries

  const Geo& g = P.g;
  const int nA = g.n[J], nX = g.n[0], nCc = g.n[DCC];
  const unsigned sA = (unsigned)((J == 1) ? g.s1 : g.s2), sCc = (unsigned)((DCC == 1) ? g.s1 : g.s2);
  const bool perA = (g.per >> J) & 1u, perX = g.per & 1u, perC = (g.per >> DCC) & 1u;
  const unsigned cA = (unsigned)P.coff[J], cC = (unsigned)P.coff[DCC];
  // dimension 3 is restricted to the planes [kz0, kz1) (z-slabs): the march range for J == 2, the columns for J == 1
  const int vcc = ((J == 1) ? P.kz0 : 2) + (int)blockIdx.y * TL::NW + wq;  // the warp's column
  if (vcc >= ((J == 1) ? P.kz1 : nCc)) return;                             // ragged tile (no CTA barrier exists in this kernel)
  const int k0 = ((J == 2) ? P.kz0 : 2) + (int)blockIdx.z * chunk, k1 = min(k0 + chunk, (J == 2) ? P.kz1 : nA);
  const int ks = k0 - 4;                         // four warm-up planes fill the register pipeline
  const int ea0 = (int)blockIdx.x * TL::TX - 2;  // 0-based element of lane 0's cell a (even: every pair is 2-element aligned)
  const int va = ea0 + 2 * lane + 1;             // 1-based x index of cell a; cell b = va + 1
  const T lr = P.lr, omlr = P.omlr, dt = P.dt;
  const T lam1 = lin_interp(T(1), lr, omlr);
  const T AA = P.A[J], AXv = P.A[0], ACv = P.A[DCC];
  const bool first = fused ? true : (P.first != 0);  // the fused sweep is always sweep 1
  const T* const rsrc = fused ? P.uOld : P.rhou_in;  // fused sweep 1: the ρu ring carries uOld, ρu = BC!(uOld*ρ(f̄)) on the fly

  // ---- per-thread constants ------------------------------------------------------------------------------------------------
  unsigned xm[2];
  bool okc[2], dirX[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int v = va + c;
    if (EDGE) {
      xm[c] = (unsigned)(mapc(v, nX, perX) - 1);
      okc[c] = lane >= 1 && v >= 2 && v <= nX - 1;
      dirX[c] = !perX && (v == 2 || v == nX);
    } else {
      xm[c] = (unsigned)(v - 1);
      okc[c] = lane >= 1;
      dirX[c] = false;
    }
  }
  // cells whose mass flux / dilation somebody uses (their own update, or as the x-1 neighbour): b of lane 0 onwards, inside the row
  const bool needA = lane >= 1 && (!EDGE || va <= nX - 1);
  const bool needB = !EDGE || va + 1 <= nX - 1;
  const bool dirC = !perC && (vcc == 2 || vcc == nCc);
  // column offsets (warp-uniform): columns c-2 .. c+1 of the f ring, mapped; own = [2], c-1 = [1]
  unsigned colm[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) colm[q] = (unsigned)(mapc(vcc - 2 + q, nCc, perC) - 1) * sCc;
  auto pm = [&](int v) -> unsigned { return (unsigned)(map1(v, nA, perA) - 1) * sA; };
  auto po = [&](int v) -> unsigned { return (unsigned)(own1(v, nA, perA) - 1) * sA; };
  auto dirAf = [&](int v) -> bool { return !perA && (v == 1 || v == 2 || v == nA); };
  const unsigned sFa = (unsigned)__cvta_generic_to_shared(sF), sRa = (unsigned)__cvta_generic_to_shared(sR);

  // f columns c-2..c+1 of plane v -> f ring slot v & 7;  ρu of plane v (own column) -> ρu ring slot v & 3
  auto ld_ring = [&](const int v) {
    const unsigned pv = pm(v);
    const unsigned sd = sFa + (unsigned)((v & 7) * PLF + 2 * lane) * SZ;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (!MOM && q == 0) continue;  // pure VOF has no column c-1 to evaluate: its boxes reach c-1 .. c+1 only
      const T* src = P.f_in + (pv + colm[q]);
      if (!EDGE) cp_async_pair<T>(sd + (unsigned)(q * FP) * SZ, src + xm[0]);
      else {
        cp_async_s(sd + (unsigned)(q * FP) * SZ, src + xm[0]);
        cp_async_s(sd + (unsigned)(q * FP + 1) * SZ, src + xm[1]);
      }
    }
    if (MOM) {
      const unsigned sr = sRa + (unsigned)((v & 3) * 3 * FP + 2 * lane) * SZ;
      const unsigned oa = po(v) + cA + colm[2], ox = pv + colm[2], oc = pv + cC + colm[2];
      if (!EDGE) {
        cp_async_pair<T>(sr, rsrc + (oa + xm[0]));
        cp_async_pair<T>(sr + FP * SZ, rsrc + (ox + xm[0]));
        cp_async_pair<T>(sr + 2 * FP * SZ, rsrc + (oc + xm[0]));
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          cp_async_s(sr + c * SZ, rsrc + (oa + xm[c]));
          cp_async_s(sr + (FP + c) * SZ, rsrc + (ox + xm[c]));
          cp_async_s(sr + (2 * FP + c) * SZ, rsrc + (oc + xm[c]));
        }
      }
    }
    cp_async_commit();
  };
  // face velocities u_a (u⁰_a) of face v and c̄ of plane vcb, own column [1] and column c-1 [0] -> registers
  auto ld_u = [&](const int v, const int vcb, T2 (&un)[2], T2 (&u0n)[2], int (&cbn)[2]) {
    const unsigned ov = po(v) + cA, ocb = pm(vcb);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!MOM && h == 0) { un[0] = un[1]; u0n[0] = u0n[1]; cbn[0] = 0; continue; }
      const unsigned col = colm[1 + h];
      if (!EDGE) {
        un[h] = ldg2(P.u + (ov + col + xm[0]));
        if (!SAMEU) u0n[h] = ldg2(P.u0 + (ov + col + xm[0]));
        if (!first) cbn[h] = (int)__ldg(reinterpret_cast<const unsigned short*>(P.cbar + (ocb + col + xm[0])));
      } else {
        un[h].x = __ldg(P.u + (ov + col + xm[0])); un[h].y = __ldg(P.u + (ov + col + xm[1]));
        if (!SAMEU) { u0n[h].x = __ldg(P.u0 + (ov + col + xm[0])); u0n[h].y = __ldg(P.u0 + (ov + col + xm[1])); }
        if (!first) cbn[h] = (int)(unsigned char)P.cbar[ocb + col + xm[0]] | ((int)(unsigned char)P.cbar[ocb + col + xm[1]] << 8);
      }
      if (SAMEU) u0n[h] = un[h];
    }
    if (!MOM) { un[0] = un[1]; u0n[0] = u0n[1]; }
  };

  // ---- rolling register state (values entering step k) ---------------------------------------------------------------------------
  T us[2][3][4];            // u★ of planes k-1, k, k+1, (k+2): [cell][A, X, C]
  T Flo[2][3];              // SynDRoM flux through face k
  T FFlo[2], Mlo[2];        // VOF / mass flux through face k
  T dilm1[2];               // dilation of plane k-1
  T f0[2], f1[2];           // f(k), f(k+1)
  T h0[2][3], h1[2][3];     // ρ at the lower a / x / c faces of cells k, k+1
  T2 uak[2], u0ak[2];       // u_a (u⁰_a) at face k: own column [1], column c-1 [0]
#pragma unroll
  for (int c = 0; c < 2; ++c) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int i = 0; i < 4; ++i) us[c][r][i] = T(0);
      Flo[c][r] = T(0); h0[c][r] = T(1); h1[c][r] = T(1);
    }
    FFlo[c] = Mlo[c] = dilm1[c] = T(0);
  }
  T rmax = -INFINITY, rmin = INFINITY;
  unsigned int amax = 0, amin = 0;

  // ---- prologue: f planes ks-1 .. ks+2, ρu planes ks .. ks+2 (one commit group), face velocities of faces ks, ks+1 -----------------
  T2 un[2], u0n[2];
  int cbn[2] = {0, 0};
  {
    // plane ks-1: f only (the box of an upwind cell in plane ks); written through ld_ring's ρu slot (ks-1)&3 = (ks+3)&3, refilled later
    ld_ring(ks - 1); ld_ring(ks); ld_ring(ks + 1); ld_ring(ks + 2);
    int cb0[2];
    ld_u(ks, ks, uak, u0ak, cb0);
    ld_u(ks + 1, ks, un, u0n, cbn);
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      f0[c] = sF[(ks & 7) * PLF + 2 * FP + 2 * lane + c];
      f1[c] = sF[((ks + 1) & 7) * PLF + 2 * FP + 2 * lane + c];
    }
  }

  unsigned lkU = (unsigned)(ks - 1) * sA + colm[2];  // offset of (plane k, own column) when k is a plain interior plane

  for (int k = ks; k < k1; ++k, lkU += sA) {
    // block-uniform boundary rules of this step
    const bool dq = dirAf(k + 2), dp = dirAf(k + 1), dpm = dirAf(k);
    const bool Lvar = !perA && k + 1 == 2, Rvar = !perA && k + 1 == nA;
    const bool needn = k + 1 <= nA && (perA || k + 1 >= 2);       // face k+1 carries a flux
    const bool ghU = !perA && (k < 2 || k > nA - 1);              // cell k is a ghost cell on a non-periodic side
    const bool ghD = !perA && (k + 1 < 2 || k + 1 > nA - 1);      // cell k+1
    const bool store = k >= k0;

    // ---- loads: rings of plane k+3, uOld of plane k (this step), u_a of face k+2 and c̄ of plane k+1 (next step) ----------------
    T2 uo[3];
    if (MOM && !fused && store) {
      const unsigned ob = pm(k) + colm[2];
      if (!EDGE) {
        uo[0] = ldg2(P.uOld + (ob + cA + xm[0]));
        uo[1] = ldg2(P.uOld + (ob + xm[0]));
        uo[2] = ldg2(P.uOld + (ob + cC + xm[0]));
      } else {
        uo[0].x = __ldg(P.uOld + (ob + cA + xm[0])); uo[0].y = __ldg(P.uOld + (ob + cA + xm[1]));
        uo[1].x = __ldg(P.uOld + (ob + xm[0])); uo[1].y = __ldg(P.uOld + (ob + xm[1]));
        uo[2].x = __ldg(P.uOld + (ob + cC + xm[0])); uo[2].y = __ldg(P.uOld + (ob + cC + xm[1]));
      }
    }
    T2 u1[2], u01[2];  // face k+1
    int cbk[2];        // c̄ of plane k
#pragma unroll
    for (int h = 0; h < 2; ++h) { u1[h] = un[h]; u01[h] = u0n[h]; cbk[h] = cbn[h]; }
    ld_ring(k + 3);
    ld_u(k + 2, k + 1, un, u0n, cbn);

    const T* F0 = sF + (k & 7) * PLF;
    const T* F1 = sF + ((k + 1) & 7) * PLF;
    const T* F2 = sF + ((k + 2) & 7) * PLF;

    // ---- A. u★ of plane k+2 (flow.jl:197): BC!(ρu/ρ(f̄)) --------------------------------------------------------------------------
    const T2 f2p = lds2<T>(F2 + 2 * FP + 2 * lane);
    T f2[2] = {f2p.x, f2p.y};
    T h2[2][3];
    if (MOM) {
      const T f2xm = (lane > 0) ? F2[2 * FP + 2 * lane - 1] : f2p.x;  // lane 0's cell a has no x-1 in the ring (and is never used)
      const T2 f2c = lds2<T>(F2 + 1 * FP + 2 * lane);
      const T* R = sR + ((k + 2) & 3) * 3 * FP + 2 * lane;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        h2[c][0] = rho_face(f2[c], f1[c], lr, omlr);
        h2[c][1] = rho_face(f2[c], c ? f2[0] : f2xm, lr, omlr);
        h2[c][2] = rho_face(f2[c], c ? f2c.y : f2c.x, lr, omlr);
        // fused: ρu = u*ρ (u2ρu!, VOFutil.jl:208-211) and straight back to u★ = ρu/ρ, rounding as the two passes would
        const T ra = t_div(fused ? R[c] * h2[c][0] : R[c], h2[c][0]);
        const T rx = t_div(fused ? R[FP + c] * h2[c][1] : R[FP + c], h2[c][1]);
        const T rc = t_div(fused ? R[2 * FP + c] * h2[c][2] : R[2 * FP + c], h2[c][2]);
        us[c][0][3] = dq ? AA : ra;  // Dirichlet planes of BC!
        us[c][1][3] = dirX[c] ? AXv : rx;
        us[c][2][3] = dirC ? ACv : rc;
      }
    }

    // ---- B. VOF flux + mass flux through face k+1 and dilation of plane k: own column [h = 1] and column c-1 [h = 0] -----------
    T FFhi[2], Mhi[2], Mhh[2], dil0[2], dilh[2], dv0[2];
    unsigned marks = 0;
#pragma unroll
    for (int h = MOM ? 0 : 1; h < 2; ++h) {
      T fl[2], fu[2];  // f(k), f(k+1) of this column
      if (h == 1) { fl[0] = f0[0]; fl[1] = f0[1]; fu[0] = f1[0]; fu[1] = f1[1]; }
      else {
        const T2 a = lds2<T>(F0 + 1 * FP + 2 * lane), b = lds2<T>(F1 + 1 * FP + 2 * lane);
        fl[0] = a.x; fl[1] = a.y; fu[0] = b.x; fu[1] = b.y;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const T uf = c ? u1[h].y : u1[h].x, u0f = c ? u01[h].y : u01[h].x;
        const T ul = c ? uak[h].y : uak[h].x, u0l = c ? u0ak[h].y : u0ak[h].x;
        T FFo = T(0), Mo = T(0);
        if (needn) {
          T dl = P.hdt * (uf + u0f);      // δt/2*(u+u⁰), advection.jl:110
          dl = (dl != T(0)) ? dl : T(0);  // -0 -> +0: the zero-flux case of advection.jl:115 without a branch
          const bool up = dl > T(0);
          const T fc = up ? fl[c] : fu[c];  // upwind cell
          const bool gho = up ? ghU : ghD;
          if (dl != T(0) && !gho && !fullorempty(fc) && (c ? needB : needA)) {
            marks |= 1u << (2 * h + c);  // interface face: reconstructed lane-dense below
            FFo = dl;                    // parked here until then
          } else {
            FFo = fc * dl;
            Mo = dl * lr + omlr * FFo;   // fᶠ2ρuf, VOFutil.jl:218
            if (MOM) Mo = Mo * P.idt;    // rmul!(ρuf, inv(δt)), flow.jl:207
          }
        }
        // dilation of plane k (flow.jl:216) and c̄[I]*(∂u+∂u⁰)*δt/2 of advection.jl:83
        const T div = (uf - ul) + (u0f - u0l);  // ∂(d,I,u)+∂(d,I,u⁰)
        const int cb = first ? ((fl[c] < T(0.5)) ? 0 : 1) : ((cbk[h] >> (8 * c)) & 0xff);  // flow.jl:172 (c̄ from the incoming f)
        const T dvv = ((cb ? div : T(0)) * dt) / T(2);
        const T dil = ((cb ? lam1 : lr) * div) / T(2);
        if (h == 1) { FFhi[c] = FFo; Mhi[c] = Mo; dv0[c] = dvv; dil0[c] = dil; }
        else { Mhh[c] = Mo; dilh[c] = dil; if (marks & (1u << c)) Mhh[c] = FFo; }
      }
    }
    if (__any_sync(FULL, marks != 0u)) {
      int cnt = 0;
      const unsigned lt = (1u << lane) - 1u;
#pragma unroll
      for (int h = MOM ? 0 : 1; h < 2; ++h) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const bool mk = (marks >> (2 * h + c)) & 1u;
          const unsigned bal = __ballot_sync(FULL, mk);
          if (mk) {
            const int p = cnt + __popc(bal & lt);
            sList[p] = (1 + h) * FP + 2 * lane + c;  // ring column of the face: c-1 -> 1, own -> 2
            sDl[p] = h ? FFhi[c] : Mhh[c];
          }
          cnt += __popc(bal);
        }
      }
      __syncwarp();
      for (int i = lane; i < cnt; i += 32) {
        const int e = sList[i];
        const T dl = sDl[i];
        const int pr = (dl > T(0)) ? k : k + 1;  // upwind cell: plane k or k+1
        ARBox<T, J> B{sF, e, pr};
        // inlined: an ABI call here would force the whole rolling register state of the march through the stack
        const T ff = plic_face_flux_inl<T, 3>(P.scheme, B, sF[(pr & 7) * PLF + e], J, dl);
        T m = dl * lr + omlr * ff;
        if (MOM) m = m * P.idt;
        sFX[e - FP] = ff;
        sMX[e - FP] = m;
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (marks & (4u << c)) { FFhi[c] = sFX[FP + 2 * lane + c]; Mhi[c] = sMX[FP + 2 * lane + c]; }
        if (MOM && (marks & (1u << c))) Mhh[c] = sMX[2 * lane + c];
      }
      __syncwarp();
    }

    // ---- C. x-1 neighbours by shuffle ----------------------------------------------------------------------------------------------
    T MhiL = T(0), dil0L = T(0);
    if (MOM) {
      MhiL = __shfl_up_sync(FULL, Mhi[1], 1);
      dil0L = __shfl_up_sync(FULL, dil0[1], 1);
    }

    // ---- D. SynDRoM momentum flux through face k+1 of the three momentum cells of each cell (flow.jl:20-57,223) ------------------
    T Fhi[2][3];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      Fhi[c][0] = Fhi[c][1] = Fhi[c][2] = T(0);
      if (MOM) {
        const T Mc = dp ? AA : Mhi[c];  // velocity BC! on ρuf (flow.jl:207)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          T Mo;
          if (r == 0) Mo = dpm ? AA : Mlo[c];
          else if (r == 1) Mo = dp ? AA : (c ? Mhi[0] : MhiL);
          else Mo = dp ? AA : Mhh[c];
          const T Psi = (Mc + Mo) / T(2);
          const T* u4 = us[c][r];
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
          T mOld = pos ? h0[c][r] : h1[c][r];
          if (r == 0) {
            if (Lvar && pos) mOld = h2[c][0];  // donor index 1: BCv! copies plane 3 = (f(3)+f(2))/2
            if (Rvar && !pos) mOld = lin_interp(__ldg(P.drho + (cA + (unsigned)(nA - 1) * sA + colm[2] + xm[c])), lr, omlr);  // donor index nA
          }
          Fhi[c][r] = syndrom_flux_t<KOREN>(P.lim, Psi, uu, cc, dd, mOld, dt);
        }
      }
    }

    // ---- E. update of cell k --------------------------------------------------------------------------------------------------------
    if (store) {
      T fn[2], qn[2][3];
      const T* R = sR + (k & 3) * 3 * FP + 2 * lane;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        T v = f0[c] + ((FFlo[c] - FFhi[c]) + dv0[c]);  // advection.jl:83
        if (okc[c]) {
          rmax = max_nan(rmax, v);
          rmin = t_min(rmin, v);
          if (v > T(1) || v < T(0)) {  // only cells outside [0,1] can be reported (reportFillError, advection.jl:145-189)
            const unsigned lk = lkU + xm[c];
            if (v >= rmax) amax = lk;
            if (v <= rmin) amin = lk;
          }
        }
        fn[c] = (v < P.tol) ? T(0) : ((v > P.onemtol) ? T(1) : v);  // cleanWisp!
        if (MOM) {
          const T dNa = (!perA && k == 2) ? dil0[c] : dilm1[c];  // BCf! (Neumann) on ρ̄∂ⱼuⱼ along the sweep direction
          T qA = R[c], qX = R[FP + c], qC = R[2 * FP + c];  // ρu before the sweep
          T oA, oX, oC;
          if (fused) {  // u2ρu! + BC!(ρu,uBC): Dirichlet plane 2 of the normal component holds uBC
            oA = qA; oX = qX; oC = qC;
            qA = dpm ? AA : qA * h0[c][0];
            qX = dirX[c] ? AXv : qX * h0[c][1];
            qC = dirC ? ACv : qC * h0[c][2];
          } else {
            oA = c ? uo[0].y : uo[0].x; oX = c ? uo[1].y : uo[1].x; oC = c ? uo[2].y : uo[2].x;
          }
          // r = Φ[I] - Φ[I+δj] + uOld*ϕ(i,I,ρ̄∂ⱼuⱼ);  ρu += δt*r          flow.jl:223-231
          const T rA = (Flo[c][0] - Fhi[c][0]) + oA * ((dil0[c] + dNa) / T(2));
          const T rX = (Flo[c][1] - Fhi[c][1]) + oX * ((dil0[c] + (c ? dil0[0] : dil0L)) / T(2));
          const T rC = (Flo[c][2] - Fhi[c][2]) + oC * ((dil0[c] + dilh[c]) / T(2));
          qn[c][0] = qA + dt * rA;
          qn[c][1] = qX + dt * rX;
          qn[c][2] = qC + dt * rC;
        }
      }
      if (!EDGE) {
        if (okc[0]) {
          const unsigned l0 = lkU + xm[0];
          T2 w;
          w.x = fn[0]; w.y = fn[1];
          *reinterpret_cast<T2*>(P.f_out + l0) = w;
          if (first) *reinterpret_cast<unsigned short*>(P.cbar + l0) = (unsigned short)(((f0[0] < T(0.5)) ? 0 : 1) | (((f0[1] < T(0.5)) ? 0 : 1) << 8));
          if (MOM) {
            w.x = qn[0][0]; w.y = qn[1][0];
            *reinterpret_cast<T2*>(P.rhou_out + (l0 + cA)) = w;
            w.x = qn[0][1]; w.y = qn[1][1];
            *reinterpret_cast<T2*>(P.rhou_out + l0) = w;
            w.x = qn[0][2]; w.y = qn[1][2];
            *reinterpret_cast<T2*>(P.rhou_out + (l0 + cC)) = w;
          } else if (P.rhouf_j != nullptr) {
            w.x = Mlo[0]; w.y = Mlo[1];
            *reinterpret_cast<T2*>(P.rhouf_j + l0) = w;
            if (k == nA - 1) { w.x = Mhi[0]; w.y = Mhi[1]; *reinterpret_cast<T2*>(P.rhouf_j + (l0 + sA)) = w; }  // inside_uWB: the upper boundary face
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (okc[c]) {
            const unsigned l0 = lkU + xm[c];
            P.f_out[l0] = fn[c];
            if (first) P.cbar[l0] = (int8_t)((f0[c] < T(0.5)) ? 0 : 1);
            if (MOM) {
              P.rhou_out[l0 + cA] = qn[c][0];
              P.rhou_out[l0] = qn[c][1];
              P.rhou_out[l0 + cC] = qn[c][2];
            } else if (P.rhouf_j != nullptr) {
              P.rhouf_j[l0] = Mlo[c];
              if (k == nA - 1) P.rhouf_j[l0 + sA] = Mhi[c];
            }
          }
        }
      }
    }

    // ---- F. roll the register pipeline ---------------------------------------------------------------------------------------------
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        Flo[c][r] = Fhi[c][r]; h0[c][r] = h1[c][r]; h1[c][r] = h2[c][r];
#pragma unroll
        for (int i = 0; i < 3; ++i) us[c][r][i] = us[c][r][i + 1];
      }
      FFlo[c] = FFhi[c]; Mlo[c] = Mhi[c]; dilm1[c] = dil0[c];
      f0[c] = f1[c]; f1[c] = f2[c];
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) { uak[h] = u1[h]; u0ak[h] = u01[h]; }
    cp_async_wait_all();  // the rings of plane k+3 have landed (issued a whole step ago)
    __syncwarp();
  }

  // ---- fill-error reduction ------------------------------------------------------------------------------------------------------------
  if (P.red != nullptr) {
    int rnan = 0;
    if (rmax != rmax) { rnan = 1; rmax = -INFINITY; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const T omax = __shfl_xor_sync(FULL, rmax, off), omin = __shfl_xor_sync(FULL, rmin, off);
      const unsigned int oamax = __shfl_xor_sync(FULL, amax, off), oamin = __shfl_xor_sync(FULL, amin, off);
      const int onan = __shfl_xor_sync(FULL, rnan, off);
      if (omax > rmax) { rmax = omax; amax = oamax; }
      if (omin < rmin) { rmin = omin; amin = oamin; }
      rnan |= onan;
    }
    if (lane == 0) {
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

template <class T, int J, bool MOM, bool FUSED, bool KOREN, int MINB, bool SAMEU>
__global__ void __launch_bounds__(256, MINB) arow_kernel(const SweepP<T> P, const int chunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nX = P.g.n[0];
  const int ea0 = (int)blockIdx.x * ARTile::TX - 2;
  // a tile is interior when every cell its lanes touch (elements ea0 .. ea0+63) is an interior cell: no index map, no Dirichlet column
  const bool edge = ea0 < 1 || ea0 + 63 > nX - 2;
  if (!edge) arow_body<T, J, MOM, FUSED, KOREN, SAMEU, false>(P, chunk, smem_raw);
  else arow_body<T, J, MOM, FUSED, KOREN, SAMEU, true>(P, chunk, smem_raw);
}

}  // namespace ifadv
