// ifadv_mlpoisson.cu -- C ABI of WaterLily's MultiLevelPoisson on the B200 path (include/ifadv.h: ifadv_ml_*); kernels in
// ifadv_mlpoisson.cuh.  Compiled -fmad=false with IEEE division in both precisions.
//
// Level 1 lives on the caller's x, L, z (Flow.p, Flow.μ₀, Flow.σ); its D, iD, ϵ, r and every array of the coarser levels belong to the
// handle, as in WaterLily (Poisson(x,L,z) allocates D, iD, ϵ, r; restrictML allocates a level).  One solver cycle -- Vcycle!(ml);
// smooth!(p); r2 = L2(p) -- is a fixed sequence of ~35 launches per level whose early exits are taken on the device (control block per
// level), so it is captured once into a CUDA graph and replayed per cycle: the host reads one scalar (r2) per cycle, as the
// reference's loop condition does.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/ifadv.h"
#include "ifadv_ctx.hpp"
#include "ifadv_mlpoisson.cuh"

using namespace ifadv;

namespace {
struct MLLevel {
  ifadv_ctx* c;  // context of this level's shape (level 1: the caller's, borrowed)
  Geo g;
  void *L, *D, *iD, *x, *eps, *r, *z;
  PoisCtl* ctl;
  unsigned nb;   // CTAs of the row-walking kernels
  unsigned nb_mult, nb_inc;  // persistent grids of the two stencil kernels: SM count x resident CTAs of that kernel (as psolver_t does)
};
}  // namespace

struct ifadv_ml {
  ifadv_ctx* c0;  // the caller's context (level 1), borrowed; not touched by ifadv_ml_destroy
  int dtype, D, device;
  unsigned per;
  std::vector<MLLevel> lv;
  double* host_r2;          // pinned
  cudaStream_t cap_stream;  // capture happens here (the caller's stream may be the legacy default stream, which cannot capture)
  cudaGraphExec_t cycle_exec;
  int64_t cycle_launches;
  int use_graph;
  int bottom;  // first level (>= 1) of the single-CTA bottom kernel; lv.size(): none
  long long march_minS;  // levels of at least this many entries use the marching form (0 when IFADV_POIS_MARCH is set explicitly)
  int march, march_occ;  // IFADV_POIS_MARCH: planes per chunk of the marching mult kernel (0: row form), its resident CTAs per SM
};

namespace {
int mfail(ifadv_ctx* c, int code, const char* msg) {
  if (c) c->err = msg;
  return code;
}
inline unsigned lv_blocks(const MLLevel& v, int D) {
  const long long rows = (long long)(v.g.n[1] - 2) * (D == 3 ? v.g.n[2] - 2 : 1);
  return (unsigned)std::max<long long>(1, std::min<long long>((rows + 7) / 8, std::min(148LL * 6, (long long)IFADV_POIS_MAXB)));
}
inline int kz1_of(const MLLevel& v, int D) { return D == 3 ? v.g.n[2] : 2; }
template <class K> unsigned lv_resident(const MLLevel& v, int D, int device, K kernel) {
  int occ = 0, sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0) != cudaSuccess || occ < 1) occ = 2;
  const long long rows = (long long)(v.g.n[1] - 2) * (D == 3 ? v.g.n[2] - 2 : 1);
  return (unsigned)std::max<long long>(1, std::min<long long>((rows + 7) / 8, std::min<long long>((long long)sms * occ, IFADV_POIS_MAXB)));
}
template <class T, int D> void lv_grids(MLLevel& v, int device) {
  v.nb_mult = lv_resident(v, D, device, ml_pcg_mult_kernel<T, D>);
  v.nb_inc = lv_resident(v, D, device, ml_increment_kernel<T, D>);
}

template <class T, int D> int perbc_lv(MLLevel& v, cudaStream_t st, T* a, unsigned per) {
  Geo g = v.g;
  g.per = per & ((1u << D) - 1u);
  if (!g.per) return 0;
  const long long n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const long long tot = ((g.per & 1u) ? 2 * n1 * n2 : 0) + ((g.per & 2u) ? 2 * n0 * n2 : 0) + ((D == 3 && (g.per & 4u)) ? 2 * n0 * n1 : 0);
  perbc_kernel<T, D><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a, g);
  v.c->launches++;
  return 0;
}

// update!(p::Poisson) of one level
template <class T, int D> void diag_lv(MLLevel& v, cudaStream_t st) {
  pois_diag_kernel<T, D><<<v.nb, 256, 0, st>>>((T*)v.D, (T*)v.iD, (const T*)v.L, v.g, 2, kz1_of(v, D));
  v.c->launches++;
}

// update!(ml): set_diag! on level 1; restrictL! + BC!(L,0,false,perdir) + set_diag! on every coarser level
template <class T, int D> int ml_update_t(ifadv_ml* m, cudaStream_t st) {
  diag_lv<T, D>(m->lv[0], st);
  const double Z[3] = {0.0, 0.0, 0.0};
  for (size_t l = 1; l < m->lv.size(); ++l) {
    MLLevel &f = m->lv[l - 1], &c = m->lv[l];
    ml_restrictL_kernel<T, D><<<c.nb, 256, 0, st>>>((T*)c.L, c.g, (const T*)f.L, f.g, 2, kz1_of(c, D));
    c.c->launches++;
    int rc = ifadv_bc_vec(c.c, (void*)st, c.L, Z, 0, m->per);
    if (rc) return mfail(m->c0, rc, "BC! of a restricted coefficient field failed");
    diag_lv<T, D>(c, st);
  }
  CU_CHECK(m->c0, cudaGetLastError());
  return 0;
}

// increment!(p): perBC!(ϵ); r -= Aϵ; x += ϵ.  (A marching form like pois_mult_march_kernel was measured and dropped: 13.48 vs 13.50 ms per
// cycle at 512^3 f32 -- the row form of this kernel already runs at 5.1 TB/s and carries two more streams per cell.)
template <class T, int D> void increment_lv(ifadv_ml* m, MLLevel& v, cudaStream_t st) {
  perbc_lv<T, D>(v, st, (T*)v.eps, m->per);
  ml_increment_kernel<T, D><<<v.nb_inc, 256, 0, st>>>((T*)v.x, (T*)v.r, (const T*)v.eps, (const T*)v.L, (const T*)v.D, v.g, 2, kz1_of(v, D));
  v.c->launches++;
}

// smooth!(p) = pcg!(p;it=6)
template <class T, int D> void smooth_lv(ifadv_ml* m, MLLevel& v, cudaStream_t st, int it = 6) {
  const int k1 = kz1_of(v, D);
  T *x = (T*)v.x, *eps = (T*)v.eps, *r = (T*)v.r, *z = (T*)v.z;
  const T *L = (const T*)v.L, *Dg = (const T*)v.D, *iD = (const T*)v.iD;
  ml_pcg_start_kernel<T, D><<<v.nb, 256, 0, st>>>(z, eps, r, iD, v.g, v.ctl, 2, k1);
  v.c->launches++;
  for (int i = 1; i <= it; ++i) {
    perbc_lv<T, D>(v, st, eps, m->per);
    if (D == 3 && m->march > 0 && v.g.S >= m->march_minS) {  // marching form of z = Aϵ on the large levels (ifadv_poisson.cuh; IFADV_POIS_MARCH=0: row form)
      const long long items = (long long)((v.g.n[0] - 2 + 31) / 32) * ((v.g.n[1] - 2 + 7) / 8) * ((k1 - 2 + m->march - 1) / m->march);
      const unsigned nbm = (unsigned)std::max<long long>(1, std::min<long long>(items, std::min<long long>(148LL * m->march_occ, IFADV_POIS_MAXB)));
      pois_mult_march_kernel<T, 1><<<nbm, 256, 0, st>>>(z, eps, L, Dg, v.g, v.ctl, 2, k1, m->march);
    } else
    ml_pcg_mult_kernel<T, D><<<v.nb_mult, 256, 0, st>>>(z, eps, L, Dg, v.g, v.ctl, 2, k1);
    ml_pcg_update_kernel<T, D><<<v.nb, 256, 0, st>>>(x, r, z, eps, iD, v.g, v.ctl, i == it ? 1 : 0, 2, k1);
    v.c->launches += 2;
    if (i == it) break;
    ml_pcg_dir_kernel<T, D><<<v.nb, 256, 0, st>>>(eps, z, v.g, v.ctl, 2, k1);
    v.c->launches++;
  }
}

// Vcycle!(ml;l)
template <class T, int D> void vcycle_t(ifadv_ml* m, cudaStream_t st, size_t l) {
  MLLevel &f = m->lv[l], &c = m->lv[l + 1];
  ml_jacobi_kernel<T, D><<<f.nb, 256, 0, st>>>((T*)f.eps, (const T*)f.r, (const T*)f.iD, f.g, 2, kz1_of(f, D));  // Jacobi!(fine)
  f.c->launches++;
  increment_lv<T, D>(m, f, st);
  ml_restrict_kernel<T, D><<<c.nb, 256, 0, st>>>((T*)c.r, (T*)c.x, c.g, (const T*)f.r, f.g, 2, kz1_of(c, D));    // restrict!; fill!(coarse.x,0)
  c.c->launches++;
  if ((int)l + 1 >= m->bottom) {  // [Vcycle!(l+1)]; smooth!(coarse) for all remaining levels in one CTA
    MLBottom<T> P{};
    P.n = (int)m->lv.size() - (int)(l + 1);
    P.per = m->per;
    for (int k = 0; k < P.n; ++k) {
      const MLLevel& v = m->lv[l + 1 + k];
      P.lv[k] = MLDev<T>{v.g, (const T*)v.L, (const T*)v.D, (const T*)v.iD, (T*)v.x, (T*)v.eps, (T*)v.r, (T*)v.z};
    }
    ml_bottom_kernel<T, D><<<1, 1024, 0, st>>>(P);
    c.c->launches++;
  } else {
    if (l + 2 < m->lv.size()) vcycle_t<T, D>(m, st, l + 1);
    smooth_lv<T, D>(m, c, st);
  }
  ml_prolongate_kernel<T, D><<<f.nb, 256, 0, st>>>((T*)f.eps, f.g, (const T*)c.x, c.g, 2, kz1_of(f, D));         // prolongate!
  f.c->launches++;
  increment_lv<T, D>(m, f, st);
}

// residual!(p) on level 1
template <class T, int D> void residual_t(ifadv_ml* m, cudaStream_t st) {
  MLLevel& v = m->lv[0];
  perbc_lv<T, D>(v, st, (T*)v.x, m->per);
  pois_residual_kernel<T, D><<<v.nb, 256, 0, st>>>((T*)v.r, (const T*)v.z, (const T*)v.x, (const T*)v.L, (const T*)v.D, (const T*)v.iD, v.g,
                                                   v.ctl, 0.0, 0, 2, kz1_of(v, D));
  ml_submean_kernel<T, D><<<v.nb, 256, 0, st>>>((T*)v.r, v.g, v.ctl, 2, kz1_of(v, D));
  v.c->launches += 2;
}

// one cycle of solver!'s loop: Vcycle!(ml); smooth!(p); r2 = L2(p)
template <class T, int D> void cycle_body(ifadv_ml* m, cudaStream_t st) {
  if (m->lv.size() > 1) vcycle_t<T, D>(m, st, 0);
  smooth_lv<T, D>(m, m->lv[0], st);
  MLLevel& v = m->lv[0];
  ml_r2_kernel<T, D><<<v.nb, 256, 0, st>>>((const T*)v.r, v.g, v.ctl, 2, kz1_of(v, D));
  v.c->launches++;
}
template <class T, int D> int cycle_t(ifadv_ml* m, cudaStream_t st) {
  if (!m->use_graph) {
    cycle_body<T, D>(m, st);
    CU_CHECK(m->c0, cudaGetLastError());
    return 0;
  }
  if (!m->cycle_exec) {
    std::vector<int64_t> before;
    for (auto& v : m->lv) before.push_back(v.c->launches);
    cudaGraph_t graph = nullptr;
    CU_CHECK(m->c0, cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
    cycle_body<T, D>(m, m->cap_stream);
    cudaError_t e = cudaStreamEndCapture(m->cap_stream, &graph);
    if (e != cudaSuccess || !graph) return mfail(m->c0, -3, "capture of the multigrid cycle failed");
    e = cudaGraphInstantiate(&m->cycle_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return mfail(m->c0, -3, "instantiation of the multigrid cycle graph failed");
    m->cycle_launches = 0;
    for (size_t l = 0; l < m->lv.size(); ++l) {
      m->cycle_launches += m->lv[l].c->launches - before[l];
      m->lv[l].c->launches = before[l];  // captured, not launched
    }
  }
  CU_CHECK(m->c0, cudaGraphLaunch(m->cycle_exec, st));
  m->c0->launches += m->cycle_launches;
  return 0;
}

// solver!(ml;tol,itmx)
template <class T, int D> int solver_t(ifadv_ml* m, cudaStream_t st, double tol, int itmx, int* cycles, double* r2_out) {
  const T tolT = tol < 0 ? T(1e-4) : (T)tol;
  if (itmx <= 0) itmx = 32;
  MLLevel& v = m->lv[0];
  residual_t<T, D>(m, st);
  CU_CHECK(m->c0, cudaGetLastError());
  int np = 0;
  double r2 = 0.0;
  while (np < itmx) {
    int rc = cycle_t<T, D>(m, st);
    if (rc) return rc;
    CU_CHECK(m->c0, cudaMemcpyAsync(m->host_r2, &v.ctl->r2, sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_CHECK(m->c0, cudaStreamSynchronize(st));
    r2 = *m->host_r2;
    ++np;
    if ((T)r2 < tolT) break;
    if (r2 != r2) break;  // NaN: the reference would spin to itmx on NaNs; report instead
  }
  perbc_lv<T, D>(v, st, (T*)v.x, m->per);
  CU_CHECK(m->c0, cudaGetLastError());
  if (cycles) *cycles = np;
  if (r2_out) *r2_out = r2;
  if (r2 != r2) return mfail(m->c0, -1, "NaN in the multigrid pressure solver");
  return 0;
}

// myproject!(a,b::MultiLevelPoisson,w), dt = T(w)·last(a.Δt): src/flow.jl:328-341 with inproject! :343-347
template <class T, int D> int ml_myproject_t(ifadv_ml* m, cudaStream_t st, T* u, double dt, int* cycles, double* r2_out) {
  MLLevel& v = m->lv[0];
  const Geo g = v.g;
  const T dtT = (T)dt;
  {
    const int bx = 128;
    const dim3 gi((unsigned)((g.n[0] + bx - 1) / bx), (unsigned)g.n[1], (unsigned)g.n[2]);
    pois_setup_kernel<T, D><<<gi, bx, 0, st>>>((T*)v.x, (T*)v.eps, (T*)v.r, (T*)v.z, u, g, dtT);   // :345 (ϵ, r cleared as in :344)
    v.c->launches++;
    CU_CHECK(m->c0, cudaGetLastError());
  }
  int rc = solver_t<T, D>(m, st, 1e-4, 200, cycles, r2_out);                                        // :346
  if (rc) return rc;
  pois_apply_kernel<T, D><<<v.nb, 256, 0, st>>>(u, (const T*)v.L, (const T*)v.x, g, 2, kz1_of(v, D));  // :331-333
  scale_kernel<T><<<(unsigned)((g.S + 255) / 256), 256, 0, st>>>((T*)v.x, T(1) / dtT, g.S);         // :334
  v.c->launches += 2;
  CU_CHECK(m->c0, cudaGetLastError());
  return 0;
}

inline bool ml_divisible(const Geo& g, int D) {  // divisible(N) = mod(N,2)==0 && N>4 on every extent of x (ghosts included)
  for (int d = 0; d < D; ++d)
    if (g.n[d] % 2 != 0 || g.n[d] <= 4) return false;
  return true;
}
}  // namespace

#define ML_DISPATCH(m, f2f, f3f, f2d, f3d)                         \
  if ((m)->dtype == IFADV_F32) return (m)->D == 2 ? f2f : f3f;     \
  return (m)->D == 2 ? f2d : f3d;

extern "C" {
int ifadv_ml_destroy(ifadv_ml* m) {
  if (!m) return 0;
  cudaSetDevice(m->device);
  for (size_t l = 0; l < m->lv.size(); ++l) {
    MLLevel& v = m->lv[l];
    cudaFree(v.D); cudaFree(v.iD); cudaFree(v.eps); cudaFree(v.r); cudaFree(v.ctl);
    if (l > 0) {
      cudaFree(v.L); cudaFree(v.x); cudaFree(v.z);
      ifadv_destroy(v.c);
    }
  }
  if (m->cycle_exec) cudaGraphExecDestroy(m->cycle_exec);
  if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
  if (m->host_r2) cudaFreeHost(m->host_r2);
  delete m;
  return 0;
}

int ifadv_ml_create(ifadv_ctx* c, ifadv_ml** out, void* stream, void* x, void* L, void* z, unsigned perdir_mask, int maxlevels) {
  if (!c || !out) return -2;
  if (!x || !L || !z) return mfail(c, -2, "null array");
  if (c->slab.nranks > 1) return mfail(c, -2, "MultiLevelPoisson is not built for z-slab contexts (use the Poisson solver, ifadv_psolver)");
  if (maxlevels <= 0) maxlevels = 10;
  CU_CHECK(c, cudaSetDevice(c->device));
  ifadv_ml* m = new ifadv_ml();
  m->c0 = c; m->dtype = c->dtype; m->D = c->D; m->device = c->device; m->per = perdir_mask & ((1u << c->D) - 1u);
  m->host_r2 = nullptr; m->cap_stream = nullptr; m->cycle_exec = nullptr; m->cycle_launches = 0;
  {
    const char* e = getenv("IFADV_ML_GRAPH");
    m->use_graph = e ? (atoi(e) != 0) : 1;
  }
  const size_t es = c->dtype == IFADV_F32 ? 4 : 8;
  auto fail = [&](int code, const char* msg) { ifadv_ml_destroy(m); return mfail(c, code, msg); };
  auto zalloc = [&](void** p, size_t bytes) { return cudaMalloc(p, bytes) == cudaSuccess && cudaMemset(*p, 0, bytes) == cudaSuccess; };
  auto add_level = [&](ifadv_ctx* lc, void* px, void* pL, void* pz) -> bool {
    MLLevel v{};
    v.c = lc; v.g = lc->g; v.g.per = 0;
    v.x = px; v.L = pL; v.z = pz;
    const size_t S = (size_t)v.g.S * es;
    m->lv.push_back(v);  // pushed first so that a failure below is cleaned up by ifadv_ml_destroy
    MLLevel& w = m->lv.back();
    if (!px && !(zalloc(&w.L, S * c->D) && zalloc(&w.x, S) && zalloc(&w.z, S))) return false;
    if (!(zalloc(&w.D, S) && zalloc(&w.iD, S) && zalloc(&w.eps, S) && zalloc(&w.r, S) && zalloc((void**)&w.ctl, sizeof(PoisCtl)))) return false;
    w.nb = lv_blocks(w, c->D);
    if (getenv("IFADV_ML_GRID888")) w.nb_mult = w.nb_inc = w.nb;  // measurement knob: the common grid for the stencil kernels too
    else if (c->dtype == IFADV_F32) { if (c->D == 2) lv_grids<float, 2>(w, c->device); else lv_grids<float, 3>(w, c->device); }
    else { if (c->D == 2) lv_grids<double, 2>(w, c->device); else lv_grids<double, 3>(w, c->device); }
    return true;
  };
  if (!add_level(c, x, L, z)) return fail(-3, "out of device memory for the multigrid levels");
  while (ml_divisible(m->lv.back().g, c->D) && (int)m->lv.size() <= maxlevels) {  // restrictML: Na = 1 + N÷2, N the interior extent (n - 2), i.e. n/2 + 1 with ghosts
    const Geo& gf = m->lv.back().g;
    int64_t Na[3] = {1 + gf.n[0] / 2, 1 + gf.n[1] / 2, c->D == 3 ? 1 + gf.n[2] / 2 : 1};
    ifadv_ctx* lc = nullptr;
    if (ifadv_create(&lc, c->D, Na, c->dtype, c->device) != 0) return fail(-3, "context of a multigrid level");
    if (!add_level(lc, nullptr, nullptr, nullptr)) return fail(-3, "out of device memory for the multigrid levels");
  }
  {  // levels of at most IFADV_ML_BOTTOM entries (default 8000, e.g. 18^3; 0: none) run in the single-CTA bottom kernel
    const char* e = getenv("IFADV_ML_BOTTOM");
    const long long cap = e ? atoll(e) : 8000;
    m->bottom = (int)m->lv.size();
    for (int l = (int)m->lv.size() - 1; l >= 1 && m->lv[l].g.S <= cap && (int)m->lv.size() - l <= IFADV_ML_BOTTOM_MAX; --l) m->bottom = l;
  }
  {
    const char* e = getenv("IFADV_POIS_MARCH");
    m->march = e ? atoi(e) : 32;
    m->march_minS = e ? 0 : 4000000;
    int occ = 0;
    cudaError_t oe = c->dtype == IFADV_F32 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pois_mult_march_kernel<float, 1>, 256, 0)
                                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pois_mult_march_kernel<double, 1>, 256, 0);
    m->march_occ = (oe == cudaSuccess && occ >= 1) ? occ : 2;
  }
  if (cudaMallocHost(&m->host_r2, sizeof(double)) != cudaSuccess) return fail(-3, "pinned memory");
  if (cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(-3, "stream");
  *out = m;
  int rc = ifadv_ml_update(m, stream);
  if (rc) { *out = nullptr; ifadv_ml_destroy(m); }
  return rc;
}

int ifadv_ml_levels(const ifadv_ml* m) { return m ? (int)m->lv.size() : -2; }

int ifadv_ml_level_array(ifadv_ml* m, int level, int which, void** dev_ptr, int64_t Ng[3]) {
  if (!m || !dev_ptr || !Ng || level < 0 || level >= (int)m->lv.size() || which < 0 || which > 6) return -2;
  MLLevel& v = m->lv[level];
  void* t[7] = {v.L, v.D, v.iD, v.x, v.eps, v.r, v.z};
  *dev_ptr = t[which];
  for (int d = 0; d < 3; ++d) Ng[d] = v.g.n[d];
  return 0;
}

int ifadv_ml_update(ifadv_ml* m, void* stream) {
  if (!m) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  ML_DISPATCH(m, (ml_update_t<float, 2>(m, st)), (ml_update_t<float, 3>(m, st)), (ml_update_t<double, 2>(m, st)), (ml_update_t<double, 3>(m, st)))
}

int ifadv_ml_residual(ifadv_ml* m, void* stream) {
  if (!m) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  if (m->dtype == IFADV_F32) { if (m->D == 2) residual_t<float, 2>(m, st); else residual_t<float, 3>(m, st); }
  else { if (m->D == 2) residual_t<double, 2>(m, st); else residual_t<double, 3>(m, st); }
  CU_CHECK(m->c0, cudaGetLastError());
  return 0;
}

int ifadv_ml_vcycle(ifadv_ml* m, void* stream) {
  if (!m) return -2;
  if (m->lv.size() < 2) return mfail(m->c0, -2, "Vcycle! needs at least two levels");
  cudaStream_t st = (cudaStream_t)stream;
  if (m->dtype == IFADV_F32) { if (m->D == 2) vcycle_t<float, 2>(m, st, 0); else vcycle_t<float, 3>(m, st, 0); }
  else { if (m->D == 2) vcycle_t<double, 2>(m, st, 0); else vcycle_t<double, 3>(m, st, 0); }
  CU_CHECK(m->c0, cudaGetLastError());
  return 0;
}

int ifadv_ml_smooth(ifadv_ml* m, void* stream, int level) {
  if (!m || level < 0 || level >= (int)m->lv.size()) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  MLLevel& v = m->lv[level];
  if (m->dtype == IFADV_F32) { if (m->D == 2) smooth_lv<float, 2>(m, v, st); else smooth_lv<float, 3>(m, v, st); }
  else { if (m->D == 2) smooth_lv<double, 2>(m, v, st); else smooth_lv<double, 3>(m, v, st); }
  CU_CHECK(m->c0, cudaGetLastError());
  return 0;
}

int ifadv_ml_solver(ifadv_ml* m, void* stream, double tol, int itmx, int* cycles, double* r2) {
  if (!m) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  ML_DISPATCH(m, (solver_t<float, 2>(m, st, tol, itmx, cycles, r2)), (solver_t<float, 3>(m, st, tol, itmx, cycles, r2)),
              (solver_t<double, 2>(m, st, tol, itmx, cycles, r2)), (solver_t<double, 3>(m, st, tol, itmx, cycles, r2)))
}

int ifadv_ml_myproject(ifadv_ml* m, void* stream, void* u, double dt, int* cycles, double* r2) {
  if (!m) return -2;
  if (!u) return mfail(m->c0, -2, "null array");
  if (!(dt != 0.0) || dt != dt) return mfail(m->c0, -2, "invalid time step");
  cudaStream_t st = (cudaStream_t)stream;
  ML_DISPATCH(m, (ml_myproject_t<float, 2>(m, st, (float*)u, dt, cycles, r2)), (ml_myproject_t<float, 3>(m, st, (float*)u, dt, cycles, r2)),
              (ml_myproject_t<double, 2>(m, st, (double*)u, dt, cycles, r2)), (ml_myproject_t<double, 3>(m, st, (double*)u, dt, cycles, r2)))
}
}  // extern "C"
