// ifadv_sweep_inst.cu -- instantiates the fused sweep kernels for ONE (T, D, MOM, family) combination, chosen with
// -DIFADV_T=float|double -DIFADV_D=2|3 -DIFADV_MOM=0|1 -DIFADV_FAM=0..6, so that the instantiations compile in parallel:
//   FAM 0  dispatcher (launch_sweep_dim) + v1 tile kernel        FAM 1  plane-marching kernel (march)
//   FAM 3  lean register marching (along2), y / z sweeps   (FAM 2, its first generation `along`, is retired)
//   FAM 4  lean plane marching along x (xsweep, CMOM only)       FAM 5  warp-autonomous rows along x (xrow, CMOM only)
//   FAM 6  cell-parallel pure-VOF sweep (vofcell, advect! only)
//   families 1-6 exist for 3-D grids only
#include <algorithm>
#include <cstdlib>
#include <limits>

#include <cooperative_groups.h>

#include "ifadv_ctx.hpp"
#ifndef IFADV_FAM
#error "compile with -DIFADV_FAM=0..5"
#endif
#if IFADV_FAM == 0
#include "ifadv_sweep.cuh"
#if IFADV_D == 2 && IFADV_MOM == 0
#include "ifadv_vofcell.cuh"
#endif
#elif IFADV_FAM == 1
#include "ifadv_march.cuh"
#elif IFADV_FAM == 3
#include "ifadv_along2.cuh"
#elif IFADV_FAM == 4
#include "ifadv_xsweep.cuh"
#elif IFADV_FAM == 5
#include "ifadv_xrow.cuh"
#else
#include "ifadv_vofcell.cuh"
#endif

namespace ifadv {

#if IFADV_FAM == 0
template <class T, int D, int J, int TX, int TY, int TZ, bool MOM>
static int launch_sweep_t(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  constexpr int NT = 256;
  using TL = Tile<D, J, TX, TY, TZ>;
  SweepP<T> P;
  P.f_in = q.f_in; P.f_out = q.f_out;
  P.u = q.u; P.u0 = q.u0;
  P.uj = q.u + (long long)J * c->g.S; P.u0j = q.u0 + (long long)J * c->g.S;
  P.cbar = q.cbar;
  P.rhou_in = q.rhou_in; P.rhou_out = q.rhou_out; P.uOld = q.uOld; P.drho = q.drho; P.uexit = q.uexit;
  P.rhouf_j = q.rhouf ? q.rhouf + (long long)J * c->g.S : nullptr;
  P.dt = (T)q.dt; P.hdt = P.dt / T(2); P.idt = T(1) / P.dt; P.lr = (T)q.lr; P.omlr = T(1) - P.lr;
  P.tol = T(10) * std::numeric_limits<T>::epsilon(); P.onemtol = T(1) - P.tol;
  for (int i = 0; i < 3; ++i) P.A[i] = (T)q.A[i];
  P.g = c->g; P.scheme = q.scheme; P.lim = q.lim; P.first = q.first; P.red = q.red;
  const size_t smem = TL::template smem_bytes<T>(MOM);
  auto kern = sweep_kernel<T, D, J, TX, TY, TZ, MOM, NT>;
  // per device (one context per device): opt in to the large dynamic shared-memory carve-out once
  static unsigned long long attr_devs = 0ull;
  if (!((attr_devs >> (c->device & 63)) & 1ull)) {
    CU_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_devs |= 1ull << (c->device & 63);
  }
  dim3 grid((unsigned)((c->g.n[0] - 2 + TL::T0 - 1) / TL::T0), (unsigned)((c->g.n[1] - 2 + TL::T1 - 1) / TL::T1),
            (unsigned)(D == 3 ? (c->g.n[2] - 2 + TL::T2 - 1) / TL::T2 : 1));
  const bool prof = c->prof_on && c->prof_ev && c->prof_n < IFADV_PROF_MAX;
  if (prof) cudaEventRecord(c->prof_ev[2 * c->prof_n], st);
  kern<<<grid, NT, smem, st>>>(P);
  if (prof) { cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], st); c->prof_tag[c->prof_n] = (unsigned char)((q.fused ? 1 : 0) | ((2 * q.j + (q.fused ? 1 : 0)) << 1)); c->prof_n++; }
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

#endif

template <class T> static void fill_params(ifadv_ctx* c, const SweepCfg<T>& q, int J, SweepP<T>& P) {
  P.f_in = q.f_in; P.f_out = q.f_out;
  P.u = q.u; P.u0 = q.u0;
  P.uj = q.u + (long long)J * c->g.S; P.u0j = q.u0 + (long long)J * c->g.S;
  P.cbar = q.cbar;
  P.rhou_in = q.rhou_in; P.rhou_out = q.rhou_out; P.uOld = q.uOld; P.drho = q.drho; P.uexit = q.uexit;
  P.rhouf_j = q.rhouf ? q.rhouf + (long long)J * c->g.S : nullptr;
  P.dt = (T)q.dt; P.hdt = P.dt / T(2); P.idt = T(1) / P.dt; P.lr = (T)q.lr; P.omlr = T(1) - P.lr;
  P.tol = T(10) * std::numeric_limits<T>::epsilon(); P.onemtol = T(1) - P.tol;
  for (int i = 0; i < 3; ++i) P.A[i] = (T)q.A[i];
  P.g = c->g; P.scheme = q.scheme; P.lim = q.lim; P.first = q.first; P.red = q.red; P.fused = q.fused;
  for (int i = 0; i < 3; ++i) P.coff[i] = (long long)i * c->g.S;
  P.kz0 = c->kz0; P.kz1 = c->kz1;
}

#if IFADV_FAM == 1
// v2: plane-marching kernel (3-D only)
template <class T, int J, int TA, int TB, bool MOM, bool FUSED, int MINB>
static int launch_march_t(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  constexpr int NT = 256;
  using TL = MTile<J, TA, TB, NT>;
  SweepP<T> P;
  fill_params<T>(c, q, J, P);
  const size_t smem = TL::template smem_bytes<T>(MOM);
  auto kern = march_kernel<T, J, TA, TB, MOM, FUSED, NT, MINB>;
  // per device (one context per device): opt in to the large dynamic shared-memory carve-out once
  static unsigned long long attr_devs = 0ull;
  if (!((attr_devs >> (c->device & 63)) & 1ull)) {
    CU_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_devs |= 1ull << (c->device & 63);
  }
  constexpr int DB = (J == 0) ? 1 : 0, DC = (J == 2) ? 1 : 2;
  const int nx = c->g.n[0] - 2, no = c->g.n[(J == 2) ? 2 : 1] - 2, nc = c->g.n[DC] - 2;
  (void)DB;
  const int tx = TL::AX ? TA : TB, to = TL::AX ? TB : TA;
  // chunk the march direction so that the grid holds several waves of CTAs
  const long long tiles = (long long)((nx + tx - 1) / tx) * ((no + to - 1) / to);
  int chunk = 64;
  while (chunk > 8 && tiles * ((nc + chunk - 1) / chunk) < 148 * 6) chunk >>= 1;
  dim3 grid((unsigned)((nx + tx - 1) / tx), (unsigned)((no + to - 1) / to), (unsigned)((nc + chunk - 1) / chunk));
  const bool prof = c->prof_on && c->prof_ev && c->prof_n < IFADV_PROF_MAX;
  if (prof) cudaEventRecord(c->prof_ev[2 * c->prof_n], st);
  kern<<<grid, NT, smem, st>>>(P, chunk);
  if (prof) { cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], st); c->prof_tag[c->prof_n] = (unsigned char)((q.fused ? 1 : 0) | ((2 * q.j + (q.fused ? 1 : 0)) << 1)); c->prof_n++; }
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

#endif

#if IFADV_FAM == 3
// v4: lean register-marching kernel (static ring slots, one barrier per plane) for sweeps along y / z (3-D only)
template <class T, int J, int CPT, bool MOM, bool FUSED, bool KOREN, int MINB, bool SAMEU = false>
static int launch_along2_t(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  // Float64 (196 registers at 256 threads = 8 warps per SM): the fused first sweep fits 168 registers without spills, so it runs 384
  // threads per CTA (12 warps: 256^3 y 1.16 -> 0.98 ms, z 1.13 -> 0.93 ms); the standard sweep spills ~12 doubles at 168 (1.14 -> 1.21 ms)
  // and stays at 256 threads, like the fused sweep with a limiter other than Koren (9..12 warps per SM all cap at 168 registers: 3 warps
  // per scheduler) -- tools/runs/gpurun_run64.sh
#ifndef IFADV_XP_A2NT64
#define IFADV_XP_A2NT64 256
#endif
#ifndef IFADV_XP_A2NT64F
#define IFADV_XP_A2NT64F 384
#endif
  constexpr int NT = (sizeof(T) == 8) ? ((MOM && FUSED && KOREN) ? IFADV_XP_A2NT64F : IFADV_XP_A2NT64) : 256, TC = (NT / 32) * CPT;
  using TL = ATile<TC>;
  SweepP<T> P;
  fill_params<T>(c, q, J, P);
  if ((unsigned long long)c->g.S * 3ull >= 0xffffffffull) { c->err = "grid too large for 32-bit element offsets"; return -2; }
  const size_t smem = TL::template smem_bytes<T>(MOM);
  auto kern = along2_kernel<T, J, CPT, MOM, FUSED, KOREN, NT, MINB, SAMEU>;
  static unsigned long long attr_devs = 0ull;
  if (!((attr_devs >> (c->device & 63)) & 1ull)) {
    CU_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_devs |= 1ull << (c->device & 63);
  }
  constexpr int DCC = (J == 1) ? 2 : 1;
  const int nzo = c->kz1 - c->kz0;  // planes of dimension 3 to update (all of them on one GPU, the owned ones of a z-slab)
  const int nx = c->g.n[0] - 2, ncc = (DCC == 2) ? nzo : c->g.n[DCC] - 2, na = (J == 2) ? nzo : c->g.n[J] - 2;
  const long long tiles = (long long)((nx + 31) / 32) * ((ncc + TC - 1) / TC);
  // march length (a multiple of 4; 4 warm-up planes per chunk): the largest candidate that still gives >= 8 CTAs per SM.  Not a power of
  // two: at 512³ chunks of 96..120 planes run 2 % faster than 128 (equal-length CTAs that start in lockstep keep alternating between their
  // load and arithmetic phases together; a shorter last chunk desynchronises them) -- tools/runs/gpurun_run39.sh, gpurun_run40.sh
  int chunk = 16;
  if (na <= 128 && tiles >= 148 * 8) chunk = (na + 3) / 4 * 4;  // thin slabs with plenty of tiles: one chunk
  else {
    static const int cand[] = {120, 88, 56, 40, 24, 16};
    for (int cc : cand)
      if (tiles * ((na + cc - 1) / cc) >= 148 * 8) { chunk = cc; break; }
    const int n = (na + chunk - 1) / chunk, last = na - (n - 1) * chunk;
    if (n > 1 && last * 4 < chunk) chunk = ((na + n - 1) / n + 3) / 4 * 4;  // no sliver at the end
  }
  if (const char* e = getenv("IFADV_CHUNK")) chunk = std::max(16, (atoi(e) / 4) * 4);  // measurement override
  {  // per kernel form: IFADV_CHUNK_Y / _YF / _Z / _ZF (F = fused first sweep)
    static const char* names[4] = {"IFADV_CHUNK_Y", "IFADV_CHUNK_YF", "IFADV_CHUNK_Z", "IFADV_CHUNK_ZF"};
    if (const char* e = getenv(names[(J - 1) * 2 + (FUSED ? 1 : 0)])) chunk = std::max(16, (atoi(e) / 4) * 4);
  }
  dim3 grid((unsigned)((nx + 31) / 32), (unsigned)((ncc + TC - 1) / TC), (unsigned)((na + chunk - 1) / chunk));
  const bool prof = c->prof_on && c->prof_ev && c->prof_n < IFADV_PROF_MAX;
  if (prof) cudaEventRecord(c->prof_ev[2 * c->prof_n], st);
  kern<<<grid, NT, smem, st>>>(P, chunk);
  if (prof) { cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], st); c->prof_tag[c->prof_n] = (unsigned char)((q.fused ? 1 : 0) | ((2 * q.j + (q.fused ? 1 : 0)) << 1)); c->prof_n++; }
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

#endif

#if IFADV_FAM == 4
// v4: lean plane-marching kernel for CMOM sweeps along x (3-D only)
template <class T, int CPT, bool FUSED, bool KOREN, int MINB, bool SAMEU = false>
static int launch_xsweep_t(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  constexpr int NT = 256, TY = (NT / 32) * CPT;
  using TL = XTile<TY>;
  SweepP<T> P;
  fill_params<T>(c, q, 0, P);
  if ((unsigned long long)c->g.S * 3ull >= 0xffffffffull) { c->err = "grid too large for 32-bit element offsets"; return -2; }
  const size_t smem = TL::template smem_bytes<T>();
  auto kern = xsweep_kernel<T, CPT, FUSED, KOREN, NT, MINB, SAMEU>;
  static unsigned long long attr_devs = 0ull;
  if (!((attr_devs >> (c->device & 63)) & 1ull)) {
    CU_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_devs |= 1ull << (c->device & 63);
  }
  const int nx = c->g.n[0] - 2, ny = c->g.n[1] - 2, nz = c->kz1 - c->kz0;
  const long long tiles = (long long)((nx + 31) / 32) * ((ny + TY - 1) / TY);
  int chunk = 128;  // 3-6 warm-up planes per chunk
  while (chunk > 16 && tiles * ((nz + chunk - 1) / chunk) < 148 * 8) chunk >>= 1;
  if (const char* e = getenv("IFADV_CHUNK")) chunk = std::max(16, (atoi(e) / 4) * 4);  // measurement override
  dim3 grid((unsigned)((nx + 31) / 32), (unsigned)((ny + TY - 1) / TY), (unsigned)((nz + chunk - 1) / chunk));
  const bool prof = c->prof_on && c->prof_ev && c->prof_n < IFADV_PROF_MAX;
  if (prof) cudaEventRecord(c->prof_ev[2 * c->prof_n], st);
  kern<<<grid, NT, smem, st>>>(P, chunk);
  if (prof) { cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], st); c->prof_tag[c->prof_n] = (unsigned char)((q.fused ? 1 : 0) | ((2 * q.j + (q.fused ? 1 : 0)) << 1)); c->prof_n++; }
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

#endif

#if IFADV_FAM == 5
#ifndef IFADV_XP_XROW_R
#define IFADV_XP_XROW_R 2
#endif
#ifndef IFADV_XP_XROW_MB
#define IFADV_XP_XROW_MB 2
#endif
// v5: warp-autonomous row kernel for CMOM sweeps along x (3-D only, even row pitch, vector-aligned arrays)
template <class T, int R, bool MOM, bool FUSED, bool KOREN, int MINB, bool SAMEU>
static int launch_xrow_t(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  using TL = XRTile<R>;
  SweepP<T> P;
  fill_params<T>(c, q, 0, P);
  if ((unsigned long long)c->g.S * 3ull >= 0xffffffffull) { c->err = "grid too large for 32-bit element offsets"; return -2; }
  const size_t smem = TL::template Bytes<T>::cta;
  auto kern = xrow_kernel<T, R, MOM, FUSED, KOREN, MINB, SAMEU>;
  static unsigned long long attr_devs = 0ull;
  if (!((attr_devs >> (c->device & 63)) & 1ull)) {
    CU_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_devs |= 1ull << (c->device & 63);
  }
  const int nx = c->g.n[0] - 2, ny = c->g.n[1] - 2, nz = c->kz1 - c->kz0;
  const unsigned gx = (unsigned)((nx + TL::TX) / TL::TX), gy = (unsigned)((ny + TL::TY - 1) / TL::TY);  // elements 1..nx in tiles [60b, 60b+59]
  // march length (one warm-up plane per chunk): 20..48 planes run 4 % faster than 64 or 128 at 512³, 12..20 are the best at
  // 512x256x256 (tools/runs/gpurun_run39.sh, gpurun_run40.sh)
  int chunk = 12;
  {
    static const int cand[] = {20, 12};
    for (int cc : cand)
      if ((long long)gx * gy * ((nz + cc - 1) / cc) >= 148 * 8) { chunk = cc; break; }
  }
  if (const char* e = getenv("IFADV_CHUNK")) chunk = std::max(4, atoi(e));  // measurement override
  if (const char* e = getenv(FUSED ? "IFADV_CHUNK_XF" : "IFADV_CHUNK_X")) chunk = std::max(4, atoi(e));
  dim3 grid(gx, gy, (unsigned)((nz + chunk - 1) / chunk));
  const bool prof = c->prof_on && c->prof_ev && c->prof_n < IFADV_PROF_MAX;
  if (prof) cudaEventRecord(c->prof_ev[2 * c->prof_n], st);
  kern<<<grid, 256, smem, st>>>(P, chunk);
  if (prof) { cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], st); c->prof_tag[c->prof_n] = (unsigned char)((q.fused ? 1 : 0) | ((2 * q.j + (q.fused ? 1 : 0)) << 1)); c->prof_n++; }
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}
#endif

#if IFADV_FAM == 6
// v6: cell-parallel pure-VOF sweep (3-D only): no shared memory, no staging
template <class T, int J, bool SAMEU> static int launch_vofcell_t(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  SweepP<T> P;
  fill_params<T>(c, q, J, P);
  if ((unsigned long long)c->g.S >= 0x7fffffffull) { c->err = "grid too large for 32-bit element offsets"; return -2; }
  if (!c->st_list) {  // list of deferred (interface) cells, shared with the surface-tension kernels: an eighth of the cells per direction
    const unsigned cap = (unsigned)std::min<long long>(std::max<long long>(c->g.S / 8, 1 << 16), 1ll << 27);
    CU_CHECK(c, cudaMalloc(&c->st_list, sizeof(int) * (size_t)cap * 3));
    CU_CHECK(c, cudaMalloc(&c->st_cnt, sizeof(unsigned) * 4));
    c->st_cap = cap;
  }
  const int nx = c->g.n[0] - 2, ny = c->g.n[1] - 2, nz = c->g.n[2] - 2;
  const long long tiles = (long long)((nx + 31) / 32) * ((ny + 7) / 8);
  int chunk = 32;  // 256³: 16 / 32 / 64 / 128 planes -> 0.378 / 0.337 / 0.340 / 0.48 ms per step (Float32)
  while (chunk > 4 && tiles * ((nz + chunk - 1) / chunk) < 148 * 8) chunk >>= 1;
  if (const char* e = getenv("IFADV_CHUNK")) chunk = std::max(1, atoi(e));  // measurement override
  dim3 grid((unsigned)((nx + 31) / 32), (unsigned)((ny + 7) / 8), (unsigned)((nz + chunk - 1) / chunk));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  const bool prof = c->prof_on && c->prof_ev && c->prof_n < IFADV_PROF_MAX;
  CU_CHECK(c, cudaMemsetAsync(c->st_cnt, 0, sizeof(unsigned), st));
  if (prof) cudaEventRecord(c->prof_ev[2 * c->prof_n], st);
  // measurement switches: IFADV_VKB = planes loaded together (1, 2, 4), IFADV_VMB = resident CTAs asked of the compiler (4, 8)
  static const int vkb = getenv("IFADV_VKB") ? atoi(getenv("IFADV_VKB")) : 4, vmb = getenv("IFADV_VMB") ? atoi(getenv("IFADV_VMB")) : 4;
  auto go = [&](auto kern) { kern<<<grid, 256, 0, st>>>(P, chunk, c->st_list, c->st_cnt, c->st_cap * 3); };
  if (vkb == 1 && vmb == 8) go(vofcell_kernel<T, J, SAMEU, 1, 8>);
  else if (vkb == 2 && vmb == 8) go(vofcell_kernel<T, J, SAMEU, 2, 8>);
  else if (vkb == 4 && vmb == 8) go(vofcell_kernel<T, J, SAMEU, 4, 8>);
  else if (vkb == 1) go(vofcell_kernel<T, J, SAMEU, 1, 4>);
  else if (vkb == 2) go(vofcell_kernel<T, J, SAMEU, 2, 4>);
  else go(vofcell_kernel<T, J, SAMEU, 4, 4>);
  vofcell_fix_kernel<T, J, SAMEU><<<(unsigned)sms * 16, 128, 0, st>>>(P, c->st_list, c->st_cnt, c->st_cap * 3);
  if (prof) { cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], st); c->prof_tag[c->prof_n] = (unsigned char)((2 * q.j) << 1); c->prof_n++; }
  c->launches += 2;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}
#endif

// per-family entry points (3-D only), each defined and explicitly instantiated in its own translation unit
template <class T, bool MOM> int launch_fam_march(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q);
template <class T, bool MOM> int launch_fam_along2(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q);
template <class T, bool MOM> int launch_fam_xsweep(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q);
template <class T, bool MOM> int launch_fam_xrow(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q);
template <class T> int launch_fam_vofcell(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q);
// the row kernel moves two cells per access: rows must start vector-aligned (even row pitch) and so must every array
template <class T> static bool xrow_ok(const ifadv_ctx* c, const SweepCfg<T>& q) {
  if (c->g.n[0] & 1) return false;
  const uintptr_t m = 2 * sizeof(T) - 1;
  const uintptr_t a = (uintptr_t)q.f_in | (uintptr_t)q.f_out | (uintptr_t)q.u | (uintptr_t)q.u0 | (uintptr_t)q.rhou_in | (uintptr_t)q.rhou_out |
                      (uintptr_t)q.uOld | (uintptr_t)q.rhouf;
  return (a & m) == 0 && ((uintptr_t)q.cbar & 1) == 0;
}

#if IFADV_FAM == 0
template <class T, int D, bool MOM> int launch_sweep_dim(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  if (c->slab.nranks > 1 && !(D == 3 && MOM && c->use_march == 1 && c->use_along2)) {
    c->err = "z-slab contexts run the 3-D CMOM path on the lean kernels only";
    return -2;
  }
  if constexpr (D == 3) {
    if constexpr (!MOM) {
      if (c->use_vofcell) return launch_fam_vofcell<T>(c, st, q);
    }
    if (c->use_march == 1 && c->use_along2) {
      if (q.j != 0) return launch_fam_along2<T, MOM>(c, st, q);
      if (c->use_xrow && xrow_ok<T>(c, q)) return launch_fam_xrow<T, MOM>(c, st, q);
      if constexpr (MOM) return launch_fam_xsweep<T, MOM>(c, st, q);
    }
    if (c->use_march) return launch_fam_march<T, MOM>(c, st, q);
  }
  if constexpr (D == 2) {
    if (q.j == 0) return launch_sweep_t<T, 2, 0, 64, 8, 1, MOM>(c, st, q);
    return launch_sweep_t<T, 2, 1, 32, 16, 1, MOM>(c, st, q);
  } else {
    if (q.j == 0) return launch_sweep_t<T, 3, 0, 32, 8, 4, MOM>(c, st, q);
    if (q.j == 1) return launch_sweep_t<T, 3, 1, 32, 8, 4, MOM>(c, st, q);
    return launch_sweep_t<T, 3, 2, 32, 4, 8, MOM>(c, st, q);
  }
}
template int launch_sweep_dim<IFADV_T, IFADV_D, (IFADV_MOM != 0)>(ifadv_ctx*, cudaStream_t, const SweepCfg<IFADV_T>&);

#if IFADV_D == 2 && IFADV_MOM == 0
// The whole 2-D pure-VOF step (advectVOF!, advection.jl:34-78) as ONE cooperative launch: reduction-slot reset, fill!(ρuf,0), the two
// directional sweeps (cell-parallel + lane-dense fix-up of the interface cells, ifadv_vofcell.cuh) and the final BCf!, separated by
// grid-wide barriers.  Small 2-D grids (BASELINE config 1: 128², 16 k cells) are launch-latency bound.
namespace cg = cooperative_groups;
constexpr int V2_NT = 128;
template <class T, int JA, int JB, bool SAMEU, bool LOCAL>
__global__ void __launch_bounds__(V2_NT) vof2d_step_kernel(const SweepP<T> PA, const SweepP<T> PB, T* f_final, T* rhouf, const long long nruf,
                                                          unsigned long long* red, const unsigned per, int* list, unsigned* cnt, const unsigned cap) {
  cg::grid_group grid = cg::this_grid();
  const Geo g = PA.g;
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x, gs = (long long)gridDim.x * blockDim.x;
  if (gt < 3) {  // red_init_kernel
    unsigned long long* r = red + 8 * gt;
    if (r[4] != 0ull) atomicOr(red + 24, 1ull);
    r[0] = 0ull; r[1] = ~0ull; r[2] = 0ull; r[3] = ~0ull; r[4] = 0ull; r[5] = 0ull; r[6] = 0ull; r[7] = 0ull;
  }
  if (gt == 0) { cnt[0] = 0u; cnt[1] = 0u; }
  if (rhouf != nullptr)
    for (long long i = gt; i < nruf; i += gs) rhouf[i] = T(0);  // fill!(ρuf,0), advection.jl:37
  grid.sync();
  vof2d_cell_sweep<T, JA, SAMEU, LOCAL>(PA, grid, list, cnt, cap);
  vof2d_cell_sweep<T, JB, SAMEU, LOCAL>(PB, grid, list + cap, cnt + 1, cap);
  // BCf!(f;perdir), VOFutil.jl:64-75: every ghost cell takes the value of its interior-equivalent cell
  const long long n0 = g.n[0], n1 = g.n[1], c0 = 2 * n1, c1 = 2 * n0;
  for (long long t = gt; t < c0 + c1; t += gs) {
    int x, y;
    if (t < c0) { x = (t & 1) ? (int)n0 : 1; y = (int)(t >> 1) + 1; }
    else { const long long q = t - c0; y = (q & 1) ? (int)n1 : 1; x = (int)(q >> 1) + 1; }
    const int mx = mapc(x, g.n[0], per & 1u), my = mapc(y, g.n[1], per & 2u);
    f_final[lin3(g, x, y, 1)] = f_final[lin3(g, mx, my, 1)];
  }
}

template <class T> int launch_vof2d_step(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& qa, const SweepCfg<T>& qb, T* f_final, T* rhouf) {
  SweepP<T> PA, PB;
  fill_params<T>(c, qa, qa.j, PA);
  fill_params<T>(c, qb, qb.j, PB);
  if ((unsigned long long)c->g.S >= 0x7fffffffull) { c->err = "grid too large for 32-bit element offsets"; return -2; }
  if (!c->st_list) {  // list of deferred (interface) cells, shared with the surface-tension kernels
    const unsigned cap = (unsigned)std::min<long long>(std::max<long long>(c->g.S / 8, 1 << 16), 1ll << 27);
    CU_CHECK(c, cudaMalloc(&c->st_list, sizeof(int) * (size_t)cap * 3));
    CU_CHECK(c, cudaMalloc(&c->st_cnt, sizeof(unsigned) * 4));
    c->st_cap = cap;
  }
  const bool same = qa.u == qa.u0;
  const long long ncell = (long long)(c->g.n[0] - 2) * (c->g.n[1] - 2);
  const bool local = ncell <= (1ll << 17) && !getenv("IFADV_VOF2D_LIST");  // latency bound: in-line reconstruction, one grid barrier per sweep
  void* kern;
  if (local) kern = (qa.j == 0) ? (same ? (void*)vof2d_step_kernel<T, 0, 1, true, true> : (void*)vof2d_step_kernel<T, 0, 1, false, true>)
                                : (same ? (void*)vof2d_step_kernel<T, 1, 0, true, true> : (void*)vof2d_step_kernel<T, 1, 0, false, true>);
  else kern = (qa.j == 0) ? (same ? (void*)vof2d_step_kernel<T, 0, 1, true, false> : (void*)vof2d_step_kernel<T, 0, 1, false, false>)
                          : (same ? (void*)vof2d_step_kernel<T, 1, 0, true, false> : (void*)vof2d_step_kernel<T, 1, 0, false, false>);
  int per_sm = 0, sms = 0;
  CU_CHECK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, V2_NT, 0));
  CU_CHECK(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  // latency bound for small grids: one cell per thread, spread over as many SMs as there are CTAs; persistent grid for large ones
  const int grid = (int)std::max<long long>(1, std::min<long long>((ncell + V2_NT - 1) / V2_NT, (long long)std::min(per_sm, 8) * sms));
  T* ruf = rhouf;
  long long nruf = (long long)c->g.S * 2;
  unsigned long long* red = c->red_dev;
  unsigned per = c->g.per, cap = c->st_cap;
  int* list = c->st_list;
  unsigned* cnt = c->st_cnt;
  void* args[] = {&PA, &PB, &f_final, &ruf, &nruf, &red, &per, &list, &cnt, &cap};
  CU_CHECK(c, cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(V2_NT), args, 0, st));
  c->launches++;
  return 0;
}
template int launch_vof2d_step<IFADV_T>(ifadv_ctx*, cudaStream_t, const SweepCfg<IFADV_T>&, const SweepCfg<IFADV_T>&, IFADV_T*, IFADV_T*);
#endif

#elif IFADV_FAM == 1
template <class T, bool MOM> int launch_fam_march(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  // tile shapes sized so that 25 shared planes leave 3 (f32) / 2-3 (f64) CTAs per SM
  constexpr int TO = (sizeof(T) == 4) ? 16 : 8;
  constexpr int TBX = (sizeof(T) == 4) ? 16 : 8;
  constexpr int MB = 2;
  if (MOM && q.fused) {
    if (q.j == 0) return launch_march_t<T, 0, 32, TBX, MOM, MOM, MB>(c, st, q);
    if (q.j == 1) return launch_march_t<T, 1, TO, 32, MOM, MOM, MB>(c, st, q);
    return launch_march_t<T, 2, TO, 32, MOM, MOM, MB>(c, st, q);
  }
  if (q.j == 0) return launch_march_t<T, 0, 32, TBX, MOM, false, MB>(c, st, q);
  if (q.j == 1) return launch_march_t<T, 1, TO, 32, MOM, false, MB>(c, st, q);
  return launch_march_t<T, 2, TO, 32, MOM, false, MB>(c, st, q);
}
template int launch_fam_march<IFADV_T, (IFADV_MOM != 0)>(ifadv_ctx*, cudaStream_t, const SweepCfg<IFADV_T>&);

#elif IFADV_FAM == 3
#ifndef IFADV_XP_A2CPT
#define IFADV_XP_A2CPT 2
#define IFADV_XP_A2MB 2
#endif
#ifndef IFADV_XP_A2MB64
#define IFADV_XP_A2MB64 1  // Float64: 196 registers without spills at 1 CTA/SM beats 128 registers + 58 spilled at 2 CTAs/SM (+5 %)
#endif
template <class T, bool MOM> int launch_fam_along2(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  constexpr int CP = (sizeof(T) == 4) ? IFADV_XP_A2CPT : 1;
  constexpr int MBX = (sizeof(T) == 4) ? IFADV_XP_A2MB : IFADV_XP_A2MB64;
  const bool koren = !MOM || q.lim == 2;  // the package default limiter is compiled in; the others go through limiter_other
  if constexpr (MOM) {
    if (koren && q.u == q.u0) {  // one velocity array for u¹ and u² (MPFMomStep!): the SAMEU instantiations
      if (q.fused) {
        if (q.j == 1) return launch_along2_t<T, 1, CP, MOM, MOM, true, MBX, true>(c, st, q);
        return launch_along2_t<T, 2, CP, MOM, MOM, true, MBX, true>(c, st, q);
      }
      if (q.j == 1) return launch_along2_t<T, 1, CP, MOM, false, true, MBX, true>(c, st, q);
      return launch_along2_t<T, 2, CP, MOM, false, true, MBX, true>(c, st, q);
    }
  }
  if (MOM && q.fused) {
    if (q.j == 1) return koren ? launch_along2_t<T, 1, CP, MOM, MOM, true, MBX>(c, st, q) : launch_along2_t<T, 1, CP, MOM, MOM, !MOM, MBX>(c, st, q);
    return koren ? launch_along2_t<T, 2, CP, MOM, MOM, true, MBX>(c, st, q) : launch_along2_t<T, 2, CP, MOM, MOM, !MOM, MBX>(c, st, q);
  }
  if (q.j == 1) return koren ? launch_along2_t<T, 1, CP, MOM, false, true, MBX>(c, st, q) : launch_along2_t<T, 1, CP, MOM, false, !MOM, MBX>(c, st, q);
  return koren ? launch_along2_t<T, 2, CP, MOM, false, true, MBX>(c, st, q) : launch_along2_t<T, 2, CP, MOM, false, !MOM, MBX>(c, st, q);
}
template int launch_fam_along2<IFADV_T, (IFADV_MOM != 0)>(ifadv_ctx*, cudaStream_t, const SweepCfg<IFADV_T>&);

#elif IFADV_FAM == 4
template <class T, bool MOM> int launch_fam_xsweep(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  static_assert(MOM, "the pure-VOF x-sweep runs the march kernel");
  constexpr int CP = (sizeof(T) == 4) ? 2 : 1;
  constexpr int MB = (sizeof(T) == 4) ? 2 : 1;
  const bool koren = q.lim == 2;
  if (koren && q.u == q.u0) {  // one velocity array for u¹ and u² (MPFMomStep!): the SAMEU instantiations
    if (q.fused) return launch_xsweep_t<T, CP, true, true, MB, true>(c, st, q);
    return launch_xsweep_t<T, CP, false, true, MB, true>(c, st, q);
  }
  if (q.fused) return koren ? launch_xsweep_t<T, CP, true, true, MB>(c, st, q) : launch_xsweep_t<T, CP, true, false, MB>(c, st, q);
  return koren ? launch_xsweep_t<T, CP, false, true, MB>(c, st, q) : launch_xsweep_t<T, CP, false, false, MB>(c, st, q);
}
template int launch_fam_xsweep<IFADV_T, (IFADV_MOM != 0)>(ifadv_ctx*, cudaStream_t, const SweepCfg<IFADV_T>&);

#elif IFADV_FAM == 5
template <class T, bool MOM> int launch_fam_xrow(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  // Float32: two rows per warp at 128 registers (2 CTAs/SM); Float64: one row per warp at 222 registers without spills (1 CTA/SM) --
  // two rows spill at 128 registers (x-sweep 2.01 ms at 256^3) and need 254 at 1 CTA/SM (1.44 ms); one row: 1.28 ms
  constexpr int RR = (sizeof(T) == 4) ? IFADV_XP_XROW_R : 1, MB = (sizeof(T) == 4) ? IFADV_XP_XROW_MB : 1;
  if constexpr (!MOM) {  // pure VOF (advect!): no limiter, no momentum streams; little work per plane, so more resident warps pay
#ifndef IFADV_XP_XROW_MB_VOF
#define IFADV_XP_XROW_MB_VOF 3
#endif
    constexpr int MBV = (sizeof(T) == 4) ? IFADV_XP_XROW_MB_VOF : MB;
    if (q.u == q.u0) return launch_xrow_t<T, RR, false, false, true, MBV, true>(c, st, q);
    return launch_xrow_t<T, RR, false, false, true, MBV, false>(c, st, q);
  } else {
    const bool koren = q.lim == 2;
    if (koren && q.u == q.u0) {  // one velocity array for u¹ and u² (MPFMomStep!): the SAMEU instantiations
      if (q.fused) return launch_xrow_t<T, RR, true, true, true, MB, true>(c, st, q);
      return launch_xrow_t<T, RR, true, false, true, MB, true>(c, st, q);
    }
#ifdef IFADV_XROW_DEV  // development builds: only the two hot instantiations, everything else runs the plane-marching kernel
    return launch_fam_xsweep<T, MOM>(c, st, q);
#else
    if (q.fused) return koren ? launch_xrow_t<T, RR, true, true, true, MB, false>(c, st, q) : launch_xrow_t<T, RR, true, true, false, MB, false>(c, st, q);
    return koren ? launch_xrow_t<T, RR, true, false, true, MB, false>(c, st, q) : launch_xrow_t<T, RR, true, false, false, MB, false>(c, st, q);
#endif
  }
}
template int launch_fam_xrow<IFADV_T, (IFADV_MOM != 0)>(ifadv_ctx*, cudaStream_t, const SweepCfg<IFADV_T>&);

#elif IFADV_FAM == 6
template <class T> int launch_fam_vofcell(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  const bool same = q.u == q.u0;
  switch (q.j) {
    case 0: return same ? launch_vofcell_t<T, 0, true>(c, st, q) : launch_vofcell_t<T, 0, false>(c, st, q);
    case 1: return same ? launch_vofcell_t<T, 1, true>(c, st, q) : launch_vofcell_t<T, 1, false>(c, st, q);
    default: return same ? launch_vofcell_t<T, 2, true>(c, st, q) : launch_vofcell_t<T, 2, false>(c, st, q);
  }
}
template int launch_fam_vofcell<IFADV_T>(ifadv_ctx*, cudaStream_t, const SweepCfg<IFADV_T>&);

#endif

}  // namespace ifadv
