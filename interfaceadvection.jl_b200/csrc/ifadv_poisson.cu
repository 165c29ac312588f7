// ifadv_poisson.cu -- C ABI of the pressure projection (include/ifadv.h: ifadv_poisson_update, ifadv_psolver, ifadv_myproject);
// kernels in ifadv_poisson.cuh.  Compiled -fmad=false with IEEE division in both precisions.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "../../include/ifadv.h"
#include "ifadv_ctx.hpp"
#include "ifadv_poisson.cuh"

using namespace ifadv;

// ifadv_b200.cu: the z-slab transport (NCCL / CUDA-IPC copy engines) behind two small helpers
int ifadv_slab_allreduce_sum(ifadv_ctx* c, cudaStream_t st, double* dev, int n);
int ifadv_slab_exchange_scalar(ifadv_ctx* c, cudaStream_t st, void* field, size_t esz, int planes);

namespace {
int pfail(ifadv_ctx* c, int code, const char* msg) {
  if (c) c->err = msg;
  return code;
}
// control block (device), its pinned mirror for two polls in flight, and their events
int pois_alloc(ifadv_ctx* c) {
  if (c->pois_ctl) return 0;
  CU_CHECK(c, cudaMalloc(&c->pois_ctl, sizeof(PoisCtl)));
  CU_CHECK(c, cudaMemset(c->pois_ctl, 0, sizeof(PoisCtl)));
  c->pois_slab_flag = 0;
  CU_CHECK(c, cudaMallocHost(&c->pois_host, 2 * 256));
  CU_CHECK(c, cudaEventCreateWithFlags(&c->pois_ev[0], cudaEventDisableTiming));
  CU_CHECK(c, cudaEventCreateWithFlags(&c->pois_ev[1], cudaEventDisableTiming));
  return 0;
}
inline long long owned_rows(const ifadv_ctx* c) { return (long long)(c->g.n[1] - 2) * (c->D == 3 ? c->kz1 - c->kz0 : 1); }
inline unsigned row_blocks(const ifadv_ctx* c) {
  const long long rows = owned_rows(c);
  return (unsigned)std::max<long long>(1, std::min<long long>((rows + 7) / 8, std::min(148LL * 6, (long long)IFADV_POIS_MAXB)));
}
// persistent grids: as many CTAs as are resident at once for that kernel (its register count decides), never more than the rows give
template <class K> unsigned resident_blocks(const ifadv_ctx* c, K kernel) {
  int occ = 0, sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0) != cudaSuccess || occ < 1) occ = 2;
  const long long rows = owned_rows(c);
  return (unsigned)std::max<long long>(1, std::min<long long>((rows + 7) / 8, std::min<long long>((long long)sms * occ, IFADV_POIS_MAXB)));
}
template <class T, int D> int perbc_launch(ifadv_ctx* c, cudaStream_t st, T* a, unsigned per) {
  Geo g = c->g;
  g.per = per & ((1u << D) - 1u);
  if (!g.per) return 0;
  const long long n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const long long tot = ((g.per & 1u) ? 2 * n1 * n2 : 0) + ((g.per & 2u) ? 2 * n0 * n2 : 0) + ((D == 3 && (g.per & 4u)) ? 2 * n0 * n1 : 0);
  perbc_kernel<T, D><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a, g);
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}
template <class T, int D> int update_t(ifadv_ctx* c, cudaStream_t st, T* Dg, T* iD, const T* L) {
  pois_diag_kernel<T, D><<<row_blocks(c), 256, 0, st>>>(Dg, iD, L, c->g, c->kz0, c->kz1);
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

// psolver!(p;tol,itmx), src/flow.jl:300-326.  z-slab contexts: every reduction is followed by an all-reduce of the ranks' sums and a
// one-thread kernel that applies the scalar step (all ranks hold identical scalars, so they take identical decisions); the ghost
// plane of ϵ (and of x at both ends) comes from the z-neighbours.
template <class T, int D>
int psolver_t(ifadv_ctx* c, cudaStream_t st, T* x, T* eps, T* r, T* z, const T* L, const T* Dg, const T* iD, unsigned per, double tol, int itmx,
              int* iters, double* r2_out) {
  int rc = pois_alloc(c);
  if (rc) return rc;
  PoisCtl* ctl = (PoisCtl*)c->pois_ctl;
  const bool slab = c->slab.nranks > 1;
  const int kz0 = c->kz0, kz1 = c->kz1;
  const unsigned nb = row_blocks(c);
  const unsigned nb_mult = resident_blocks(c, pois_mult_kernel<T, D>), nb_upd = resident_blocks(c, pois_update_kernel<T, D>),
                 nb_dir = resident_blocks(c, pois_dir_kernel<T, D>);
  const Geo g = c->g;
  // z = Aϵ as a march along z (pois_mult_march_kernel) on large 3-D single-GPU grids: 512^3 f32 1.74 -> 1.63 ms per iteration, 256^3 f64
  // 0.449 -> 0.434; chunks of 32 planes (64: +1 %, whole columns: +14 %).  IFADV_POIS_MARCH=<planes per chunk> overrides, 0 = row form.
  int march = 0;
  unsigned nb_march = 1;
  if (D == 3 && !slab) {
    const char* e = getenv("IFADV_POIS_MARCH");
    march = e ? atoi(e) : (g.S >= 4000000 ? 32 : 0);
    if (march > 0) {
      int occ = 0, sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pois_mult_march_kernel<T, 0>, 256, 0) != cudaSuccess || occ < 1) occ = 2;
      const long long items = (long long)((g.n[0] - 2 + 31) / 32) * ((g.n[1] - 2 + 7) / 8) * ((kz1 - kz0 + march - 1) / march);
      nb_march = (unsigned)std::max<long long>(1, std::min<long long>(items, std::min<long long>((long long)sms * occ, IFADV_POIS_MAXB)));
    }
  }
  // very small grids (<= 100 k entries per field, e.g. 32^3 or BASELINE config 1's 128^2): batches of iterations as ONE cooperative
  // launch -- 16.9 vs 27.8 us per iteration at 32^3; from 64^3 on the grid-wide barriers cost what the launches cost (24.0 vs 24.5 us)
  // and at 128^3 more (53 vs 41 us), so larger grids keep the three-kernel form (IFADV_POIS_COOP=0/1 overrides)
  bool coop = false;
  unsigned nb_coop = 0;
  Geo gp = g;
  gp.per = per & ((1u << D) - 1u);
  if (!slab) {
    const char* e = getenv("IFADV_POIS_COOP");
    int can = 0;
    cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, c->device);
    coop = can && (e ? atoi(e) != 0 : g.S <= 100000);
    if (coop) {
      nb_coop = resident_blocks(c, pois_pcg_coop_kernel<T, D>);
      const char* eb = getenv("IFADV_POIS_COOP_CTAS");  // measurement knob: cap of the cooperative grid
      if (eb && atoi(eb) > 0) nb_coop = std::min<unsigned>(nb_coop, (unsigned)atoi(eb));
    }
  }
  const double tolT = tol < 0 ? (double)(T(50) * std::numeric_limits<T>::epsilon()) : (double)(T)tol;
  if (itmx <= 0) itmx = 6000;
  if (c->pois_slab_flag != (slab ? 1 : 0)) {
    const int v = slab ? 1 : 0;
    CU_CHECK(c, cudaMemcpyAsync(&ctl->slab, &v, sizeof(int), cudaMemcpyHostToDevice, st));
    CU_CHECK(c, cudaStreamSynchronize(st));
    c->pois_slab_flag = v;
  }
  auto reduce_fin = [&](int which, int nacc) -> int {  // z-slab: all-reduce the sums of the kernel just launched, then the scalar step
    if (!slab) return 0;
    int e = ifadv_slab_allreduce_sum(c, st, ctl->acc, nacc);
    if (e) return e;
    pois_fin_kernel<T><<<1, 1, 0, st>>>(ctl, which, tolT, itmx);
    c->launches++;
    return 0;
  };
  auto ghosts = [&](T* a) -> int {  // perBC! in x, y (and z on one GPU) + the z-neighbours' plane
    int e = perbc_launch<T, D>(c, st, a, per);
    if (e) return e;
    return slab ? ifadv_slab_exchange_scalar(c, st, a, sizeof(T), 1) : 0;
  };
  if ((rc = ghosts(x))) return rc;                                                                  // :301 (and residual!'s own)
  pois_residual_kernel<T, D><<<nb, 256, 0, st>>>(r, z, x, L, Dg, iD, g, ctl, tolT, itmx, kz0, kz1);  // :302
  if ((rc = reduce_fin(0, 2))) return rc;
  pois_start_kernel<T, D><<<nb, 256, 0, st>>>(r, z, eps, iD, g, ctl, kz0, kz1);                      // :302-307
  if ((rc = reduce_fin(1, 2))) return rc;
  c->launches += 2;
  CU_CHECK(c, cudaGetLastError());
  // Iterations are enqueued in batches; the control block of batch k is read back while batch k+1 is already queued, so the device
  // never waits for the host.  Kernels past convergence return immediately.
  struct Poll { double rho, zeps, beta, r2, mean, tol, r2_0; int n, itmx, done, sub_mean; };
  static_assert(sizeof(Poll) <= 256, "poll mirror");
  char* host = (char*)c->pois_host;
  auto poll = [&](int slot) -> int {
    CU_CHECK(c, cudaMemcpyAsync(host + 256 * slot, ctl, sizeof(Poll), cudaMemcpyDeviceToHost, st));
    CU_CHECK(c, cudaEventRecord(c->pois_ev[slot], st));
    return 0;
  };
  const int batch_max = slab ? 8 : 64;  // the collectives of a z-slab run are not skipped past convergence: keep the overshoot short
  int it = 0, last = 0, batch = slab ? 4 : 8;
  Poll res{};
  if ((rc = poll(last))) return rc;  // behind the start kernel
  for (;;) {
    bool queued = false;
    if (it < itmx) {
      const int end = std::min(itmx, it + batch);
      if (coop) {
        int it0 = it, it1 = end, k0 = kz0, k1 = kz1;
        void* args[] = {(void*)&x, (void*)&r, (void*)&z, (void*)&eps, (void*)&L, (void*)&Dg, (void*)&iD, (void*)&gp, (void*)&ctl, (void*)&it0,
                        (void*)&it1, (void*)&k0, (void*)&k1};
        CU_CHECK(c, cudaLaunchCooperativeKernel((void*)pois_pcg_coop_kernel<T, D>, dim3(nb_coop), dim3(256), args, 0, st));
        c->launches++;
        it = end;
      }
      for (; it < end; ++it) {
        if ((rc = ghosts(eps))) return rc;                                                           // :311
        if (march) pois_mult_march_kernel<T, 0><<<nb_march, 256, 0, st>>>(z, eps, L, Dg, g, ctl, kz0, kz1, march);
        else pois_mult_kernel<T, D><<<nb_mult, 256, 0, st>>>(z, eps, L, Dg, g, ctl, kz0, kz1);       // :312-313
        if ((rc = reduce_fin(2, 1))) return rc;
        pois_update_kernel<T, D><<<nb_upd, 256, 0, st>>>(x, r, z, eps, iD, g, ctl, kz0, kz1);        // :313-321
        if ((rc = reduce_fin(3, 2))) return rc;
        pois_dir_kernel<T, D><<<nb_dir, 256, 0, st>>>(eps, z, g, ctl, it, kz0, kz1);                 // :319
        c->launches += 3;
      }
      CU_CHECK(c, cudaGetLastError());
      if ((rc = poll(last ^ 1))) return rc;
      queued = true;
      batch = std::min(batch_max, batch * 2);
    }
    CU_CHECK(c, cudaEventSynchronize(c->pois_ev[last]));
    memcpy(&res, host + 256 * last, sizeof(Poll));
    if (res.done || !queued) break;  // converged (what is still queued returns at once), or all itmx iterations were behind this poll
    last ^= 1;
  }
  if ((rc = ghosts(x))) return rc;                                                                  // :325
  if (iters) *iters = res.n;
  if (r2_out) *r2_out = res.r2;
  if (res.r2 != res.r2) return pfail(c, -1, "NaN in the pressure solver");
  return 0;
}

// myproject!(a,b,w) with dt = T(w)·last(a.Δt), src/flow.jl:328-347
template <class T, int D>
int myproject_t(ifadv_ctx* c, cudaStream_t st, T* u, T* x, T* eps, T* r, T* z, const T* L, const T* Dg, const T* iD, double dt, unsigned per,
                int* iters, double* r2_out) {
  const Geo g = c->g;
  const T dtT = (T)dt;
  {
    const int bx = 128;
    const dim3 gi((unsigned)((g.n[0] + bx - 1) / bx), (unsigned)g.n[1], (unsigned)g.n[2]);
    pois_setup_kernel<T, D><<<gi, bx, 0, st>>>(x, eps, r, z, u, g, dtT);                            // :344-345
    c->launches++;
    CU_CHECK(c, cudaGetLastError());
  }
  int rc = psolver_t<T, D>(c, st, x, eps, r, z, L, Dg, iD, per, -1.0, 2000, iters, r2_out);         // :346
  if (rc) return rc;
  pois_apply_kernel<T, D><<<row_blocks(c), 256, 0, st>>>(u, L, x, g, c->kz0, c->kz1);               // :331-333 (owned planes)
  scale_kernel<T><<<(unsigned)((g.S + 255) / 256), 256, 0, st>>>(x, T(1) / dtT, g.S);                // :334
  c->launches += 2;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}
}  // namespace

void ifadv_poisson_free(ifadv_ctx* c) {
  if (!c || !c->pois_ctl) return;
  cudaFree(c->pois_ctl);
  cudaFreeHost(c->pois_host);
  cudaEventDestroy(c->pois_ev[0]);
  cudaEventDestroy(c->pois_ev[1]);
  c->pois_ctl = nullptr;
}

#define POIS_DISPATCH(call2f, call3f, call2d, call3d)            \
  if (c->dtype == IFADV_F32) return c->D == 2 ? call2f : call3f; \
  return c->D == 2 ? call2d : call3d;

extern "C" {
int ifadv_poisson_update(ifadv_ctx* c, void* stream, void* Dg, void* iD, const void* L) {
  if (!c) return -2;
  if (!Dg || !iD || !L) return pfail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  POIS_DISPATCH((update_t<float, 2>(c, st, (float*)Dg, (float*)iD, (const float*)L)), (update_t<float, 3>(c, st, (float*)Dg, (float*)iD, (const float*)L)),
                (update_t<double, 2>(c, st, (double*)Dg, (double*)iD, (const double*)L)),
                (update_t<double, 3>(c, st, (double*)Dg, (double*)iD, (const double*)L)))
}

int ifadv_psolver(ifadv_ctx* c, void* stream, void* x, void* eps, void* r, void* z, const void* L, const void* Dg, const void* iD,
                  unsigned perdir_mask, double tol, int itmx, int* iters, double* r2) {
  if (!c) return -2;
  if (!x || !eps || !r || !z || !L || !Dg || !iD) return pfail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
#define A_(T) (T*)x, (T*)eps, (T*)r, (T*)z, (const T*)L, (const T*)Dg, (const T*)iD, perdir_mask, tol, itmx, iters, r2
  POIS_DISPATCH((psolver_t<float, 2>(c, st, A_(float))), (psolver_t<float, 3>(c, st, A_(float))), (psolver_t<double, 2>(c, st, A_(double))),
                (psolver_t<double, 3>(c, st, A_(double))))
#undef A_
}

int ifadv_myproject(ifadv_ctx* c, void* stream, void* u, void* x, void* eps, void* r, void* z, const void* L, const void* Dg, const void* iD,
                    double dt, unsigned perdir_mask, int* iters, double* r2) {
  if (!c) return -2;
  if (!u || !x || !eps || !r || !z || !L || !Dg || !iD) return pfail(c, -2, "null array");
  if (!(dt != 0.0) || dt != dt) return pfail(c, -2, "invalid time step");
  cudaStream_t st = (cudaStream_t)stream;
#define A_(T) (T*)u, (T*)x, (T*)eps, (T*)r, (T*)z, (const T*)L, (const T*)Dg, (const T*)iD, dt, perdir_mask, iters, r2
  POIS_DISPATCH((myproject_t<float, 2>(c, st, A_(float))), (myproject_t<float, 3>(c, st, A_(float))), (myproject_t<double, 2>(c, st, A_(double))),
                (myproject_t<double, 3>(c, st, A_(double))))
#undef A_
}
}  // extern "C"
