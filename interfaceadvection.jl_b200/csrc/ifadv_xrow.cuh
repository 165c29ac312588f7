// ifadv_xrow.cuh -- fused CMOM directional sweep along x (J = 0) for 3-D grids: WARP-AUTONOMOUS rows, warp-shuffle neighbour exchange.
//
// x is the contiguous dimension, so the sweep direction runs across the lanes of a warp.  A WARP owns a strip of 60 (x) x R (y)
// cells and marches along z on its own: there is no CTA-wide barrier anywhere in the march.
//   * a thread owns TWO x-adjacent cells (a, b) of every row, so every global access is one aligned 8-byte (Float32) / 16-byte
//     (Float64) vector load or store; lanes 1..30 produce output (60 cells), lanes 0 and 31 only supply the x-2 / x-1 / x+1
//     neighbours.  Tiles overlap by four cells along x (6 % redundant lanes) instead of exchanging a halo;
//   * what the stencil reaches ALONG x (the 4-point u★ line, the donor cell's face densities, mass flux and dilation of x-1, the
//     SynDRoM and VOF fluxes of the face x+1) moves between lanes with __shfl_up/down; inside a thread cell a is cell b's x-1;
//   * what it reaches along y (mass flux and dilation of the row y-1) is recomputed: a warp evaluates the VOF face flux and the
//     dilation of R+1 rows (its own and the one below) -- 24 instructions per cell and extra row instead of shared planes + a barrier;
//   * what it reaches along z (f, mass flux, dilation of plane k-1) stays in the owner's registers;
//   * ρu, uOld, u_x and c̄ go straight from global memory to registers (each value has exactly one consumer); only f passes through
//     shared memory, a warp-private 4-plane ring of R+3 rows filled with cp.async, because the PLIC reconstruction of an interface
//     cell needs its 3^3 box.  Faces whose upwind cell holds an interface are marked, compacted with a ballot and reconstructed
//     lane-dense by the warp (general branch of getVOFFlux!, advection.jl:131-134).
// One plane is finished per step (no software skew): VOF flux + mass flux + dilation of the plane (advection.jl:108-137, flow.jl:216),
// u★ = BC!(ρu/ρ(f̄)) (flow.jl:197), SynDRoM momentum fluxes (flow.jl:20-57,223), update of f (advection.jl:83, cleanWisp!) and of ρu
// (flow.jl:224-231), fill-error extrema.  Arithmetic (expression by expression) and boundary rules are those of ifadv_xsweep.cuh /
// ifadv_march.cuh<J=0>; tiles that touch a ghost column run the EDGE instantiation (index maps, scalar accesses; WALL adds the
// Dirichlet planes of BC!, ϕuL/ϕuR and the ghost upwind cells of a non-periodic x boundary).
#pragma once
#include "ifadv_xsweep.cuh"

namespace ifadv {

template <class T> struct V2;
template <> struct V2<float> { using type = float2; };
template <> struct V2<double> { using type = double2; };

template <int R> struct XRTile {
  static constexpr int NW = 8;          // warps per CTA (independent of each other)
  static constexpr int TX = 60;         // cells a warp owns along x: lanes 1..30, two each
  static constexpr int TY = NW * R;     // rows per CTA
  static constexpr int FR = R + 3;      // rows of the warp's f ring: y-2 .. y+R
  static constexpr int FP = 66;         // row pitch: [pad, b(-1), a0, b0, ..., a31, b31]
  static constexpr int PLF = FR * FP;   // one f plane
  static constexpr int NF = (R + 1) * 64;  // faces a warp evaluates per plane
  // per warp: f ring x4, fᶠ and mass flux of reconstructed faces, δl of listed faces, the list
  static constexpr int WELEMS = 4 * PLF + 2 * (R + 1) * FP + NF;
  template <class T> struct Bytes {
    static constexpr size_t warp = (sizeof(T) * (size_t)WELEMS + sizeof(int) * (size_t)NF + 15) / 16 * 16;
    static constexpr size_t cta = NW * warp;
  };
};

template <class T, int PLF, int FP> struct RBox {  // 3^3 box on the warp's f ring (x fastest, rows along y, ring along z)
  const T* sF;
  int e, k;  // entry of the box centre within a plane, plane index
  IFADV_DI T operator()(int dx, int dy, int dz) const { return sF[((k + dz) & 3) * PLF + e + dx + dy * FP]; }
};

template <class T> IFADV_DI void cp_async_pair(unsigned saddr, const T* gsrc) {
  if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(gsrc));
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gsrc));
}
template <class T> IFADV_DI typename V2<T>::type ldg2(const T* p) { return __ldg(reinterpret_cast<const typename V2<T>::type*>(p)); }
template <class T> IFADV_DI typename V2<T>::type lds2(const T* p) { return *reinterpret_cast<const typename V2<T>::type*>(p); }

template <class T, int R, bool MOM, bool FUSED, bool KOREN, bool SAMEU, bool EDGE>
IFADV_DI void xrow_body(const SweepP<T>& P, const int chunk, unsigned char* smem_raw) {
  using TL = XRTile<R>;
  using T2 = typename V2<T>::type;
  constexpr int FP = TL::FP, PLF = TL::PLF, FR = TL::FR;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned SZ = sizeof(T);
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  T* sF = reinterpret_cast<T*>(smem_raw + (size_t)wq * TL::template Bytes<T>::warp);  // [4][FR][FP]
  T* sFX = sF + 4 * PLF;                                                                // [R+1][FP] fᶠ of reconstructed faces
  T* sMX = sFX + (R + 1) * FP;                                                          // [R+1][FP] their mass flux
  T* sDl = sMX + (R + 1) * FP;                                                          // [NF] δl of the listed faces
  int* sList = reinterpret_cast<int*>(sDl + TL::NF);                                    // [NF] face entries

  const Geo& g = P.g;
  const int nA = g.n[0], nB = g.n[1], nC = g.n[2];
  const unsigned s1 = (unsigned)g.s1, s2 = (unsigned)g.s2;
  const bool perA = g.per & 1u, perB = (g.per >> 1) & 1u, perC = (g.per >> 2) & 1u;
  const bool WALL = EDGE && !perA;  // block-uniform: the rules of a non-periodic x boundary apply (flag bits; none are set otherwise)
  const unsigned cB = (unsigned)P.coff[1], cC = (unsigned)P.coff[2];
  const int ea0 = (int)blockIdx.x * TL::TX - 2;  // 0-based element of lane 0's cell a (even: every pair is 2-element aligned)
  const int va = ea0 + 2 * lane + 1;             // 1-based x index of cell a; cell b = va + 1
  const int vy0 = 2 + (int)blockIdx.y * TL::TY + wq * R;  // the warp's first row
  if (vy0 > nB - 1) return;                      // ragged tile: nothing to own (no CTA barrier exists in this kernel)
  const int k0 = P.kz0 + (int)blockIdx.z * chunk, k1 = min(k0 + chunk, P.kz1);  // planes [k0, k1) of this CTA
  const T lr = P.lr, omlr = P.omlr, dt = P.dt;
  const T lam1 = lin_interp(T(1), lr, omlr);
  const T AA = P.A[0], AB = P.A[1], AC = P.A[2];
  const bool first = FUSED ? true : (P.first != 0);  // the fused sweep is always sweep 1
  const T* const rsrc = FUSED ? P.uOld : P.rhou_in;  // fused sweep 1: ρu = BC!(uOld*ρ(f̄)) is formed on the fly
  constexpr int R0 = MOM ? 0 : 1;                    // first S1 row: the row y-1 only feeds the momentum fluxes

  // ---- per-thread constants ------------------------------------------------------------------------------------------------
  // x offsets of the two cells: mapped (f, c̄, ρu_y, ρu_z, uOld: ghost -> interior-equivalent cell) and as stored (u_x faces, ρu_x)
  unsigned xm[2], xs[2], flg[2];
  bool okc[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int v = va + c;
    if (EDGE) {
      xm[c] = (unsigned)(mapc(v, nA, perA) - 1);
      xs[c] = (unsigned)((perA ? wrapc(v, nA) : min(max(v, 1), nA)) - 1);
      flg[c] = WALL ? xflags(v, nA, perA, P.uexit != nullptr) : (unsigned)XF_NEEDM;
      okc[c] = lane >= 1 && lane <= 30 && v >= 2 && v <= nA - 1;
    } else {
      xm[c] = xs[c] = (unsigned)(v - 1);
      flg[c] = XF_NEEDM;
      okc[c] = lane >= 1 && lane <= 30;
    }
  }
  // faces whose flux somebody uses: b of lane 0 .. a of lane 31, inside the row
  const bool needA = lane >= 1 && (!EDGE || va <= nA), needB = lane <= 30 && (!EDGE || va + 1 <= nA);
  // row offsets (warp-uniform): rows y-2 .. y+R of the f ring, mapped
  unsigned rowm[FR];
  bool rv[R], dirB[R];
#pragma unroll
  for (int r = 0; r < FR; ++r) rowm[r] = (unsigned)(mapc(vy0 - 2 + r, nB, perB) - 1) * s1;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    rv[j] = vy0 + j <= nB - 1;
    dirB[j] = !perB && (vy0 + j == 2 || vy0 + j == nB);
  }
  // the extra element b(-1) of row `lane` of the f ring (the 3^3 box of lane 0's cell a)
  const unsigned xrow = (lane < FR) ? (unsigned)(mapc(vy0 - 2 + lane, nB, perB) - 1) * s1 + (unsigned)(mapc(ea0, nA, perA) - 1) : 0u;
  auto pm = [&](int v) -> unsigned { return (unsigned)(map1(v, nC, perC) - 1) * s2; };
  const unsigned sFa = (unsigned)__cvta_generic_to_shared(sF);

  // f rows of plane v -> ring slot v & 3
  auto ld_f = [&](const int v) {
    const unsigned pv = pm(v);
    const unsigned sd = sFa + (unsigned)((v & 3) * PLF + 2 + 2 * lane) * SZ;
#pragma unroll
    for (int r = 0; r < FR; ++r) {
      const T* src = P.f_in + (pv + rowm[r]);
      if (!EDGE) cp_async_pair<T>(sd + (unsigned)(r * FP) * SZ, src + xm[0]);
      else {
        cp_async_s(sd + (unsigned)(r * FP) * SZ, src + xm[0]);
        cp_async_s(sd + (unsigned)(r * FP + 1) * SZ, src + xm[1]);
      }
    }
    if (lane < FR) cp_async_s(sFa + (unsigned)((v & 3) * PLF + lane * FP + 1) * SZ, P.f_in + (pv + xrow));
    cp_async_commit();
  };
  // u_x (and u⁰_x, c̄) of plane v, rows y-1 .. y+R-1 -> registers
  auto ld_u = [&](const int v, T2 (&un)[R + 1], T2 (&u0n)[R + 1], int (&cbn)[R + 1]) {
    const unsigned pv = pm(v);
#pragma unroll
    for (int r = 0; r <= R; ++r) {
      const unsigned o = pv + rowm[r + 1];
      if (!EDGE) {
        un[r] = ldg2(P.u + (o + xs[0]));
        if (!SAMEU) u0n[r] = ldg2(P.u0 + (o + xs[0]));
        if (!first) cbn[r] = (int)__ldg(reinterpret_cast<const unsigned short*>(P.cbar + (o + xm[0])));
      } else {
        un[r].x = __ldg(P.u + (o + xs[0])); un[r].y = __ldg(P.u + (o + xs[1]));
        if (!SAMEU) { u0n[r].x = __ldg(P.u0 + (o + xs[0])); u0n[r].y = __ldg(P.u0 + (o + xs[1])); }
        if (!first) cbn[r] = (int)(unsigned char)P.cbar[o + xm[0]] | ((int)(unsigned char)P.cbar[o + xm[1]] << 8);
      }
      if (SAMEU) u0n[r] = un[r];
    }
  };

  // ---- rolling register state along z (values of plane k-1 entering step k) ---------------------------------------------------
  T fz[R][2], Mz[R][2], dilz[R][2];
  T rmax = -INFINITY, rmin = INFINITY;
  unsigned int amax = 0, amin = 0;

  // S1 of one plane: VOF face flux, mass flux, dilation of rows jlo-1.. (r = j+1) from the f ring and the face velocities uc/u0c.
  // Marked interface faces are reconstructed lane-dense by the warp before the function returns.
  auto s1_plane = [&](const int k, const bool warm, const T2 (&uc)[R + 1], const T2 (&u0c)[R + 1], const int (&cbc)[R + 1],
                      T (&FF)[R + 1][2], T (&M)[R + 1][2], T (&dil)[R + 1][2], T (&dv)[R + 1][2]) {
    unsigned marks = 0;
    const T* Fk = sF + (k & 3) * PLF;
#pragma unroll
    for (int r = 0; r <= R; ++r) {
      if ((warm || !MOM) && r == 0) {  // the warm-up plane only feeds the z-1 terms of the warp's own rows; pure VOF needs no row y-1
        FF[0][0] = FF[0][1] = M[0][0] = M[0][1] = dil[0][0] = dil[0][1] = dv[0][0] = dv[0][1] = T(0);
        continue;
      }
      const T* Frow = Fk + (r + 1) * FP;
      const T2 fo = lds2<T>(Frow + 2 + 2 * lane);
      const T fxm = Frow[1 + 2 * lane];  // f(x-1) of cell a
      const T ua = uc[r].x, ub = uc[r].y, u0a = u0c[r].x, u0b = u0c[r].y;
      const T uan = __shfl_down_sync(FULL, ua, 1);  // face x+1 of cell b
      const T u0an = SAMEU ? uan : __shfl_down_sync(FULL, u0a, 1);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const T uf = c ? ub : ua, u0f = c ? u0b : u0a;
        const T fup = c ? fo.x : fxm, fown = c ? fo.y : fo.x;
        T FFo = T(0), Mo = T(0);
        if (!EDGE || (flg[c] & XF_NEEDM)) {
          T dl = P.hdt * (uf + u0f);        // δt/2*(u+u⁰), advection.jl:110
          dl = (dl != T(0)) ? dl : T(0);    // -0 -> +0: the zero-flux case of advection.jl:115 without a branch
          const bool up = dl > T(0);
          const T fc = up ? fup : fown;     // upwind cell x-1 / x, advection.jl:120
          const bool gh = EDGE && (flg[c] & (up ? XF_GHLO : XF_GHHI));
          if (dl != T(0) && !gh && !fullorempty(fc) && (c ? needB : needA)) {
            marks |= 1u << (2 * r + c);     // interface face: reconstructed lane-dense below
            FFo = dl;                       // parked here until then
          } else {
            FFo = fc * dl;                            // advection.jl:125-126
            Mo = dl * lr + omlr * FFo;                // fᶠ2ρuf (VOFutil.jl:218)
            if (MOM) Mo = Mo * P.idt;                 // rmul!(ρuf, inv(δt)) (flow.jl:207)
          }
        }
        FF[r][c] = FFo; M[r][c] = Mo;
        // dilation of the cell (flow.jl:216) and c̄[I]*(∂u+∂u⁰)*δt/2 of advection.jl:83
        const T div = c ? ((uan - ub) + (u0an - u0b)) : ((ub - ua) + (u0b - u0a));  // ∂(d,I,u)+∂(d,I,u⁰)
        const int cb = first ? ((fown < T(0.5)) ? 0 : 1) : ((cbc[r] >> (8 * c)) & 0xff);  // flow.jl:172 (c̄ from the incoming f)
        dv[r][c] = ((cb ? div : T(0)) * dt) / T(2);
        dil[r][c] = ((cb ? lam1 : lr) * div) / T(2);
      }
      if (EDGE && (flg[0] & XF_DILSH)) dil[r][0] = dil[r][1];  // BCf! (Neumann) on ρ̄∂ⱼuⱼ along the sweep direction, flow.jl:217
    }
    if (__any_sync(FULL, marks != 0u)) {
      int cnt = 0;
      const unsigned lt = (1u << lane) - 1u;
#pragma unroll
      for (int r = 0; r <= R; ++r) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const bool mk = (marks >> (2 * r + c)) & 1u;
          const unsigned bal = __ballot_sync(FULL, mk);
          if (mk) {
            const int p = cnt + __popc(bal & lt);
            sList[p] = r * FP + 2 + 2 * lane + c;
            sDl[p] = FF[r][c];
          }
          cnt += __popc(bal);
        }
      }
      __syncwarp();
      for (int i = lane; i < cnt; i += 32) {
        const int e = sList[i];
        const T dl = sDl[i];
        const int eu = ((dl > T(0)) ? e - 1 : e) + FP;  // upwind cell; list row r <-> ring row r+1
        RBox<T, PLF, FP> B{sF, eu, k};
        const T ff = plic_face_flux_inl<T, 3>(P.scheme, B, Fk[eu], 0, dl);
        sFX[e] = ff;
        T m = dl * lr + omlr * ff;
        if (MOM) m = m * P.idt;
        sMX[e] = m;
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r <= R; ++r) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if ((marks >> (2 * r + c)) & 1u) { FF[r][c] = sFX[r * FP + 2 + 2 * lane + c]; M[r][c] = sMX[r * FP + 2 + 2 * lane + c]; }
      }
      __syncwarp();
    }
  };

  // ---- prologue: f planes k0-2 .. k0 and the face velocities of the warm-up plane k0-1 ------------------------------------------
  T2 un[R + 1], u0n[R + 1];
  int cbn[R + 1];
#pragma unroll
  for (int r = 0; r <= R; ++r) cbn[r] = 0;
  if (MOM) { ld_f(k0 - 2); ld_f(k0 - 1); ld_f(k0); }
  else { ld_f(k0 - 1); ld_f(k0); ld_f(k0 + 1); }
  ld_u(MOM ? k0 - 1 : k0, un, u0n, cbn);
  cp_async_wait_all();
  __syncwarp();

  // The march starts one plane early: step k0-1 only evaluates S1 (it feeds the z-1 terms -- f, mass flux, dilation -- of plane k0).
  const int kstart = MOM ? k0 - 1 : k0;          // pure VOF has no z-1 terms: no warm-up plane
  unsigned pk = (unsigned)(kstart - 1) * s2;  // offset of plane k (owned planes are interior: no map)
  for (int k = kstart; k < k1; ++k, pk += s2) {
    const bool warm = k < k0;
    // ---- loads of this step: f(k+2) -> ring, ρu / uOld of plane k and u_x / c̄ of plane k+1 -> registers -------------------
    T2 q[R][3], o[R][3];
    if (MOM && !warm) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const unsigned ob = pk + rowm[j + 2];
        if (!EDGE) {
          q[j][0] = ldg2(rsrc + (ob + xs[0]));
          q[j][1] = ldg2(rsrc + (ob + cB + xm[0]));
          q[j][2] = ldg2(rsrc + (ob + cC + xm[0]));
          if (!FUSED) {
            o[j][0] = ldg2(P.uOld + (ob + xm[0]));
            o[j][1] = ldg2(P.uOld + (ob + cB + xm[0]));
            o[j][2] = ldg2(P.uOld + (ob + cC + xm[0]));
          }
        } else {
          // exitBC: the cell at plane nA takes the saved exit value of u★ (component x) instead of ρu/ρ
          q[j][0].x = __ldg(((flg[0] & XF_EXIT) ? P.uexit : rsrc) + (ob + xs[0])); q[j][0].y = __ldg(((flg[1] & XF_EXIT) ? P.uexit : rsrc) + (ob + xs[1]));
          q[j][1].x = __ldg(rsrc + (ob + cB + xm[0])); q[j][1].y = __ldg(rsrc + (ob + cB + xm[1]));
          q[j][2].x = __ldg(rsrc + (ob + cC + xm[0])); q[j][2].y = __ldg(rsrc + (ob + cC + xm[1]));
          if (!FUSED) {
            o[j][0].x = __ldg(P.uOld + (ob + xm[0])); o[j][0].y = __ldg(P.uOld + (ob + xm[1]));
            o[j][1].x = __ldg(P.uOld + (ob + cB + xm[0])); o[j][1].y = __ldg(P.uOld + (ob + cB + xm[1]));
            o[j][2].x = __ldg(P.uOld + (ob + cC + xm[0])); o[j][2].y = __ldg(P.uOld + (ob + cC + xm[1]));
          }
        }
      }
    }
    T2 uc[R + 1], u0c[R + 1];
    int cbc[R + 1];
#pragma unroll
    for (int r = 0; r <= R; ++r) { uc[r] = un[r]; u0c[r] = u0n[r]; cbc[r] = cbn[r]; }
    ld_f(k + 2);
    ld_u(k + 1, un, u0n, cbn);

    // ---- S1: VOF flux, mass flux, dilation of rows y-1 .. y+R-1 -----------------------------------------------------------------
    T FF[R + 1][2], M[R + 1][2], dil[R + 1][2], dv[R + 1][2];
    s1_plane(k, warm, uc, u0c, cbc, FF, M, dil, dv);

    // ---- S2: u★, SynDRoM fluxes, update -- row by row ------------------------------------------------------------------------------
    const bool dirC = !perC && (k == 2 || k == nC);
    const T* Fk = sF + (k & 3) * PLF;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int r = j + 1;
      const T* Frow = Fk + (r + 1) * FP;
      const T2 fo = lds2<T>(Frow + 2 + 2 * lane);
      if (warm) {  // roll only
        fz[j][0] = fo.x; fz[j][1] = fo.y;
        Mz[j][0] = M[r][0]; Mz[j][1] = M[r][1];
        dilz[j][0] = dil[r][0]; dilz[j][1] = dil[r][1];
        continue;
      }
      if (!MOM) {  // pure VOF (advectVOF!, advection.jl:34-78): the f update and the optional ρuf[·,x] output -- nothing else
        const T FFRa = __shfl_down_sync(FULL, FF[r][0], 1);
        const T MRa = __shfl_down_sync(FULL, M[r][0], 1);
        T fn[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const T f0 = c ? fo.y : fo.x;
          T v = f0 + ((FF[r][c] - (c ? FFRa : FF[r][1])) + dv[r][c]);  // advection.jl:67
          if (okc[c] && rv[j]) {
            rmax = max_nan(rmax, v);
            rmin = t_min(rmin, v);
            if (v > T(1) || v < T(0)) {
              const unsigned lk = pk + rowm[r + 1] + xm[c];
              if (v >= rmax) amax = lk;
              if (v <= rmin) amin = lk;
            }
          }
          fn[c] = (v < P.tol) ? T(0) : ((v > P.onemtol) ? T(1) : v);  // cleanWisp!
        }
        if (rv[j]) {
          const unsigned lk = pk + rowm[r + 1];
          if (!EDGE) {
            if (okc[0]) {
              const unsigned l0 = lk + xm[0];
              T2 w;
              w.x = fn[0]; w.y = fn[1];
              *reinterpret_cast<T2*>(P.f_out + l0) = w;
              if (first) *reinterpret_cast<unsigned short*>(P.cbar + l0) = (unsigned short)(((fo.x < T(0.5)) ? 0 : 1) | (((fo.y < T(0.5)) ? 0 : 1) << 8));
              if (P.rhouf_j != nullptr) { w.x = M[r][0]; w.y = M[r][1]; *reinterpret_cast<T2*>(P.rhouf_j + l0) = w; }
            }
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (okc[c]) {
                const unsigned l0 = lk + xm[c];
                P.f_out[l0] = fn[c];
                if (first) P.cbar[l0] = (int8_t)(((c ? fo.y : fo.x) < T(0.5)) ? 0 : 1);
                if (P.rhouf_j != nullptr) {
                  P.rhouf_j[l0] = M[r][c];
                  if (va + c == nA - 1) P.rhouf_j[l0 + 1] = c ? MRa : M[r][1];  // inside_uWB includes the upper boundary face
                }
              }
            }
          }
        }
        continue;
      }
      const T fxm = Frow[1 + 2 * lane];
      const T2 fy = lds2<T>(Frow - FP + 2 + 2 * lane);
      // ρ at the lower x / y / z faces of the two cells (the ρ(f̄) u★ is formed with, and the SynDRoM donor density)
      T h[2][3];
      h[0][0] = rho_face(fo.x, fxm, lr, omlr);  h[0][1] = rho_face(fo.x, fy.x, lr, omlr);  h[0][2] = rho_face(fo.x, fz[j][0], lr, omlr);
      h[1][0] = rho_face(fo.y, fo.x, lr, omlr); h[1][1] = rho_face(fo.y, fy.y, lr, omlr);  h[1][2] = rho_face(fo.y, fz[j][1], lr, omlr);
      // u★ = BC!(ρu/ρ(f̄)) (flow.jl:197, VOFutil.jl:198-201); fused: ρu = u*ρ (u2ρu!) formed on the fly, rounding as the two passes would
      T us[2][3];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const T qa = c ? q[j][0].y : q[j][0].x, qb = c ? q[j][1].y : q[j][1].x, qc = c ? q[j][2].y : q[j][2].x;
        const T ra = t_div(FUSED ? qa * h[c][0] : qa, h[c][0]);
        const T rb = t_div(FUSED ? qb * h[c][1] : qb, h[c][1]);
        const T rc = t_div(FUSED ? qc * h[c][2] : qc, h[c][2]);
        us[c][0] = (EDGE && (flg[c] & XF_DIRA)) ? AA : ((EDGE && (flg[c] & XF_EXIT)) ? qa : ra);  // Dirichlet planes of BC! / saved exit plane
        us[c][1] = dirB[j] ? AB : rb;
        us[c][2] = dirC ? AC : rc;
      }
      // neighbours along x: lane-1 supplies its cells a', b' (x-2, x-1 of cell a), lane+1 its cell a'' (x+1 of cell b)
      T uLa[3], uLb[3], hLb[3], uRa[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        uLa[i] = __shfl_up_sync(FULL, us[0][i], 1);
        uLb[i] = __shfl_up_sync(FULL, us[1][i], 1);
        hLb[i] = __shfl_up_sync(FULL, h[1][i], 1);
        uRa[i] = __shfl_down_sync(FULL, us[0][i], 1);
      }
      const T MLb = __shfl_up_sync(FULL, M[r][1], 1);
      const T dilLb = __shfl_up_sync(FULL, dil[r][1], 1);
      const T hRa0 = EDGE ? __shfl_down_sync(FULL, h[0][0], 1) : T(0);  // ρ at the x-face of cell x+1 (ϕuL donor rule)

      // SynDRoM fluxes through the lower x-face of cell c: stencil (x-2, x-1, x, x+1)
      T Fl[2][3];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const unsigned fg = flg[c];
        const bool dira = EDGE && (fg & XF_DIRA), Lvar = EDGE && (fg & XF_LVAR), Rvar = EDGE && (fg & XF_RVAR);
        const T Mc = dira ? AA : M[r][c];  // velocity BC! on ρuf: Dirichlet planes of component x (flow.jl:207)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          T Mo;
          if (i == 0) Mo = (EDGE && (fg & XF_DIRAM)) ? AA : (c ? M[r][0] : MLb);
          else if (i == 1) Mo = dira ? AA : M[r - 1][c];
          else Mo = dira ? AA : Mz[j][c];
          const T Psi = (Mc + Mo) / T(2);
          const T um2 = c ? uLb[i] : uLa[i], um1 = c ? us[0][i] : uLb[i], ucr = us[c][i], up1 = c ? uRa[i] : us[1][i];
          const bool pos = Psi > T(0);
          T uu, cc, dd;
          if (Lvar) {  // ϕuL, flow.jl:28-31
            if (pos) { uu = T(2) * um1 - ucr; cc = um1; dd = ucr; }
            else { uu = up1; cc = ucr; dd = um1; }
          } else if (Rvar) {  // ϕuR, flow.jl:32-35
            if (Psi < T(0)) { uu = T(2) * ucr - um1; cc = ucr; dd = um1; }
            else { uu = um2; cc = um1; dd = ucr; }
          } else {  // ϕu, flow.jl:20-23
            uu = pos ? um2 : up1;
            cc = pos ? um1 : ucr;
            dd = pos ? ucr : um1;
          }
          // density of the donor momentum cell (x-1 for Ψ>0, else x): the same ρ(f̄) its u★ was formed with
          T mOld = pos ? (c ? h[0][i] : hLb[i]) : h[c][i];
          if (EDGE && i == 0) {
            if (Lvar && pos) mOld = c ? hRa0 : h[1][0];  // donor index 1: BCv! copies plane 3 = (f(3)+f(2))/2
            if (Rvar && !pos)                            // donor index nA: the plane f2face! never writes
              mOld = lin_interp(__ldg(P.drho + ((unsigned)(nA - 1) + rowm[r + 1] + pk)), lr, omlr);
          }
          Fl[c][i] = syndrom_flux_t<KOREN>(P.lim, Psi, uu, cc, dd, mOld, dt);
        }
      }
      T FRa[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) FRa[i] = __shfl_down_sync(FULL, Fl[0][i], 1);  // fluxes through the face x+1 of cell b
      const T FFRa = __shfl_down_sync(FULL, FF[r][0], 1);

      // update of the two cells
      T fn[2], qn[2][3];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const T f0 = c ? fo.y : fo.x;
        const T FFhi = c ? FFRa : FF[r][1];
        T v = f0 + ((FF[r][c] - FFhi) + dv[r][c]);  // advection.jl:83
        if (okc[c] && rv[j]) {
          rmax = max_nan(rmax, v);
          rmin = t_min(rmin, v);
          if (v > T(1) || v < T(0)) {  // only cells outside [0,1] can be reported (reportFillError, advection.jl:145-189)
            const unsigned lk = pk + rowm[r + 1] + xm[c];
            if (v >= rmax) amax = lk;
            if (v <= rmin) amin = lk;
          }
        }
        fn[c] = (v < P.tol) ? T(0) : ((v > P.onemtol) ? T(1) : v);  // cleanWisp!
        const T dK = dil[r][c];
        T qa = c ? q[j][0].y : q[j][0].x, qb = c ? q[j][1].y : q[j][1].x, qc = c ? q[j][2].y : q[j][2].x;
        T oa, ob_, oc;
        if (FUSED) {  // u2ρu! + BC!(ρu,uBC): Dirichlet plane 2 of the normal component holds uBC
          oa = qa; ob_ = qb; oc = qc;
          qa = (EDGE && (flg[c] & XF_LVAR)) ? AA : qa * h[c][0];
          qb = dirB[j] ? AB : qb * h[c][1];
          qc = dirC ? AC : qc * h[c][2];
        } else {
          oa = c ? o[j][0].y : o[j][0].x; ob_ = c ? o[j][1].y : o[j][1].x; oc = c ? o[j][2].y : o[j][2].x;
        }
        // r = Φ[I] - Φ[I+δj] + uOld*ϕ(i,I,ρ̄∂ⱼuⱼ);  ρu += δt*r          flow.jl:223-231
        const T dxm = c ? dil[r][0] : dilLb;
        const T ra = (Fl[c][0] - (c ? FRa[0] : Fl[1][0])) + oa * ((dK + dxm) / T(2));
        const T rb = (Fl[c][1] - (c ? FRa[1] : Fl[1][1])) + ob_ * ((dK + dil[r - 1][c]) / T(2));
        const T rc = (Fl[c][2] - (c ? FRa[2] : Fl[1][2])) + oc * ((dK + dilz[j][c]) / T(2));
        qn[c][0] = qa + dt * ra;
        qn[c][1] = qb + dt * rb;
        qn[c][2] = qc + dt * rc;
      }
      if (rv[j]) {
        const unsigned lk = pk + rowm[r + 1];  // owned rows / planes are interior: the mapped offset is the cell itself
        if (!EDGE) {
          if (okc[0]) {
            const unsigned l0 = lk + xm[0];
            T2 w;
            w.x = fn[0]; w.y = fn[1];
            *reinterpret_cast<T2*>(P.f_out + l0) = w;
            w.x = qn[0][0]; w.y = qn[1][0];
            *reinterpret_cast<T2*>(P.rhou_out + l0) = w;
            w.x = qn[0][1]; w.y = qn[1][1];
            *reinterpret_cast<T2*>(P.rhou_out + (l0 + cB)) = w;
            w.x = qn[0][2]; w.y = qn[1][2];
            *reinterpret_cast<T2*>(P.rhou_out + (l0 + cC)) = w;
            if (first) {
              const unsigned short cbw = (unsigned short)(((fo.x < T(0.5)) ? 0 : 1) | (((fo.y < T(0.5)) ? 0 : 1) << 8));
              *reinterpret_cast<unsigned short*>(P.cbar + l0) = cbw;
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (okc[c]) {
              const unsigned l0 = lk + xm[c];
              P.f_out[l0] = fn[c];
              P.rhou_out[l0] = qn[c][0];
              P.rhou_out[l0 + cB] = qn[c][1];
              P.rhou_out[l0 + cC] = qn[c][2];
              if (first) P.cbar[l0] = (int8_t)(((c ? fo.y : fo.x) < T(0.5)) ? 0 : 1);
            }
          }
        }
      }
      // roll the z pipeline
      fz[j][0] = fo.x; fz[j][1] = fo.y;
      Mz[j][0] = M[r][0]; Mz[j][1] = M[r][1];
      dilz[j][0] = dil[r][0]; dilz[j][1] = dil[r][1];
    }
    cp_async_wait_all();  // f(k+2) has landed (issued a whole step ago)
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
      red_commit<T>(P.red, rmax, rmin, amax, amin, rnan);
    }
  }
}

template <class T, int R, bool MOM, bool FUSED, bool KOREN, int MINB, bool SAMEU>
__global__ void __launch_bounds__(256, MINB) xrow_kernel(const SweepP<T> P, const int chunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nA = P.g.n[0];
  const int ea0 = (int)blockIdx.x * XRTile<R>::TX - 2;
  // a tile is interior when every cell its lanes touch (elements ea0-1 .. ea0+63) is an interior cell: no index map, no boundary rule
  const bool edge = ea0 - 1 < 1 || ea0 + 63 > nA - 2;
  if (!edge) xrow_body<T, R, MOM, FUSED, KOREN, SAMEU, false>(P, chunk, smem_raw);
  else xrow_body<T, R, MOM, FUSED, KOREN, SAMEU, true>(P, chunk, smem_raw);
}

}  // namespace ifadv
