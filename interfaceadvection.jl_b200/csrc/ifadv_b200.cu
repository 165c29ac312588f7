// ifadv_b200.cu -- libifadv_b200.so: C ABI (include/ifadv.h) + the small field kernels around the fused sweep.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared
// There is no CPU fallback: every entry point needs a CUDA device and fails with -3 otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/ifadv.h"
#include "ifadv_ctx.hpp"
#include "ifadv_forcing.cuh"
#include "ifadv_post.cuh"

using namespace ifadv;

static size_t esize(int dtype) { return dtype == IFADV_F32 ? 4 : 8; }
static int fail(ifadv_ctx* c, int code, const char* msg) {
  if (c) c->err = msg;
  return code;
}

// ------------------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------------------
// Resets the per-sweep reduction slots of a call.  Before that, a NaN count left by the PREVIOUS call (which may have run without a
// report, i.e. without anybody looking) is folded into the sticky flag red[24]: ifadv_check_nan / the next reporting call see it.
__global__ void red_init_kernel(unsigned long long* red, int nsets) {
  const int t = threadIdx.x;
  if (t < nsets) {
    unsigned long long* r = red + 8 * t;
    if (r[4] != 0ull) atomicOr(red + 24, 1ull);
    r[0] = 0ull; r[1] = ~0ull; r[2] = 0ull; r[3] = ~0ull; r[4] = 0ull; r[5] = 0ull; r[6] = 0ull; r[7] = 0ull;
  }
}

// BCf!(f;perdir), VOFutil.jl:64-75: every ghost cell takes the value of its interior-equivalent cell
// (clamp = Neumann copy, wrap = periodic).  One launch over the six (four) boundary planes.
template <class T, int D> __global__ void bcf_kernel(T* f, const Geo g) {
  const long long n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const long long c0 = 2 * n1 * n2, c1 = 2 * n0 * n2, c2 = (D == 3) ? 2 * n0 * n1 : 0;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= c0 + c1 + c2) return;
  int x, y, z;
  if (t < c0) { const long long r = t >> 1; x = (t & 1) ? (int)n0 : 1; y = (int)(r % n1) + 1; z = (int)(r / n1) + 1; }
  else if (t < c0 + c1) { const long long q = t - c0, r = q >> 1; y = (q & 1) ? (int)n1 : 1; x = (int)(r % n0) + 1; z = (int)(r / n0) + 1; }
  else { const long long q = t - c0 - c1, r = q >> 1; z = (q & 1) ? (int)n2 : 1; x = (int)(r % n0) + 1; y = (int)(r / n0) + 1; }
  const int mx = mapc(x, g.n[0], g.per & 1u), my = mapc(y, g.n[1], g.per & 2u), mz = (D == 3) ? mapc(z, g.n[2], g.per & 4u) : 1;
  f[lin3(g, x, y, z)] = f[lin3(g, mx, my, mz)];
}

// WaterLily.BC!(a,A,saveexit,perdir) for a constant tuple A: closed form of the sequential
// plane passes -- Dirichlet planes {1,2,N} of the normal component, otherwise the interior-equivalent cell.
template <class T, int D> __global__ void bcvec_kernel(T* a, const Geo g, T A0, T A1, T A2, int saveexit) {
  // threads enumerate planes {1,2,N} of every dimension
  const long long n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const long long c0 = 3 * n1 * n2, c1 = 3 * n0 * n2, c2 = (D == 3) ? 3 * n0 * n1 : 0;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= c0 + c1 + c2) return;
  int x, y, z;
  auto plane = [](int s, long long n) { return s == 0 ? 1 : (s == 1 ? 2 : (int)n); };
  if (t < c0) { const long long r = t / 3; x = plane((int)(t % 3), n0); y = (int)(r % n1) + 1; z = (int)(r / n1) + 1; }
  else if (t < c0 + c1) { const long long q = t - c0, r = q / 3; y = plane((int)(q % 3), n1); x = (int)(r % n0) + 1; z = (int)(r / n0) + 1; }
  else { const long long q = t - c0 - c1, r = q / 3; z = plane((int)(q % 3), n2); x = (int)(r % n0) + 1; y = (int)(r / n0) + 1; }
  const int v[3] = {x, y, z};
  const T A[3] = {A0, A1, A2};
  const bool ghost = (x == 1 || x == g.n[0] || y == 1 || y == g.n[1] || (D == 3 && (z == 1 || z == g.n[2])));
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const bool peri = (g.per >> i) & 1u;
    const bool dirichlet = !peri && (v[i] == 1 || v[i] == 2 || (v[i] == g.n[i] && !(saveexit && i == 0)));
    if (!ghost && !dirichlet) continue;  // plane 2 is interior for the other components
    T val;
    if (dirichlet) val = A[i];
    else {
      int m[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const bool perk = (g.per >> k) & 1u;
        if (k >= D) m[k] = 1;
        else if (k == i) m[k] = perk ? wrapc(v[k], g.n[k]) : v[k];
        else m[k] = mapc(v[k], g.n[k], perk);
      }
      val = a[(long long)i * g.S + lin3(g, m[0], m[1], m[2])];
    }
    a[(long long)i * g.S + lin3(g, x, y, z)] = val;
  }
}

// u2ρu! / ρu2u!, VOFutil.jl:198-211
template <class T, int D, bool TO_RHOU> __global__ void urhou_kernel(T* out, const T* in, const T* f, const Geo g, T lr, T omlr) {
  const int x = 2 + blockIdx.x * blockDim.x + threadIdx.x, y = 2 + blockIdx.y, z = (D == 3) ? 2 + blockIdx.z : 1;
  if (x > g.n[0] - 1) return;
  const long long l = lin3(g, x, y, z);
  const T fc = f[l];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const long long sd = (d == 0) ? 1 : ((d == 1) ? g.s1 : g.s2);
    const T rho = lin_interp((fc + f[l - sd]) / T(2), lr, omlr);
    const long long ld = (long long)d * g.S + l;
    out[ld] = TO_RHOU ? in[ld] * rho : in[ld] / rho;
  }
}

template <class T> __global__ void fill_kernel(T* out, T v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}
// out = a*x + b*y over the whole storage (midpoint f⁰, flow.jl:74): 16-byte vectors, four of them in flight per thread
template <class T> struct alignas(16) Vec16 { T v[16 / sizeof(T)]; };
template <class T, bool VEC> __global__ void __launch_bounds__(256) axpby_kernel(T* out, T a, const T* x, T b, const T* y, long long n, int same) {
  constexpr int V = VEC ? (int)(16 / sizeof(T)) : 1;
  const long long nv = n / V, stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    const Vec16<T>* xv = reinterpret_cast<const Vec16<T>*>(x);
    const Vec16<T>* yv = reinterpret_cast<const Vec16<T>*>(y);
    Vec16<T>* ov = reinterpret_cast<Vec16<T>*>(out);
    for (; i + 3 * stride < nv; i += 4 * stride) {
      Vec16<T> p[4], q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { p[k] = xv[i + k * stride]; q[k] = yv[i + k * stride]; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int e = 0; e < V; ++e) p[k].v[e] = same ? (p[k].v[e] + q[k].v[e]) * a : a * p[k].v[e] + b * q[k].v[e];
        ov[i + k * stride] = p[k];
      }
    }
    for (; i < nv; i += stride) {
      Vec16<T> p = xv[i], q = yv[i];
#pragma unroll
      for (int e = 0; e < V; ++e) p.v[e] = same ? (p.v[e] + q.v[e]) * a : a * p.v[e] + b * q.v[e];
      ov[i] = p;
    }
    const long long t = nv * V + (long long)blockIdx.x * blockDim.x + threadIdx.x;  // tail elements
    if (t < n) out[t] = same ? (x[t] + y[t]) * a : a * x[t] + b * y[t];
  } else {
    for (; i < n; i += stride) out[i] = same ? (x[i] + y[i]) * a : a * x[i] + b * y[i];
  }
}

// Field reductions over inside(f): persistent grid, a warp walks whole x-rows (coalesced), per-lane partials, one atomic per CTA.
// MPCFL's two reductions (flow.jl:267-271): max flux_out and max maxTotalFlux over inside(σ)
template <class T, int D> __global__ void __launch_bounds__(256) cfl_kernel(const T* u, const Geo g, unsigned long long* red, int kz0, int kz1) {
  const int ny = g.n[1] - 2, nz = (D == 3) ? kz1 - kz0 : 1;
  const long long rows = (long long)ny * nz;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double fo = 0.0, tf = 0.0;
  for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * wpb) {
    const int y = 2 + (int)(r % ny), z = (D == 3) ? kz0 + (int)(r / ny) : 1;
    const long long l0 = lin3(g, 0, y, z);
    for (int x = 2 + lane; x <= g.n[0] - 1; x += 32) {
      const long long l = l0 + x;
      double f1 = 0.0;
      T s2 = T(0);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const long long sd = (i == 0) ? 1 : ((i == 1) ? g.s1 : g.s2);
        const T ul = u[(long long)i * g.S + l], uh = u[(long long)i * g.S + l + sd];
        f1 += fmax(0.0, (double)uh) + fmax(0.0, -(double)ul);  // `max(0.,…)` promotes in WaterLily's flux_out
        s2 += t_max(t_abs(ul), t_abs(uh));
      }
      fo = fmax(fo, (double)(T)f1);  // stored into σ::T before maximum()
      tf = fmax(tf, (double)s2);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    fo = fmax(fo, __shfl_xor_sync(0xffffffffu, fo, off));
    tf = fmax(tf, __shfl_xor_sync(0xffffffffu, tf, off));
  }
  __shared__ double w0[8], w1[8];
  if (lane == 0) { w0[threadIdx.x >> 5] = fo; w1[threadIdx.x >> 5] = tf; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < wpb; ++w) { fo = fmax(fo, w0[w]); tf = fmax(tf, w1[w]); }
    atomicMax(red + 0, ord_key(fo));
    atomicMax(red + 1, ord_key(tf));
  }
}

template <class T, int D> __global__ void __launch_bounds__(256) sum_inside_kernel(const T* f, const Geo g, double* out, int kz0, int kz1) {
  const int ny = g.n[1] - 2, nz = (D == 3) ? kz1 - kz0 : 1;
  const long long rows = (long long)ny * nz;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double s = 0.0;
  for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * wpb) {
    const int y = 2 + (int)(r % ny), z = (D == 3) ? kz0 + (int)(r / ny) : 1;
    const long long l0 = lin3(g, 0, y, z);
    for (int x = 2 + lane; x <= g.n[0] - 1; x += 32) s += (double)f[l0 + x];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  __shared__ double ws[8];
  if (lane == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < wpb; ++w) t += ws[w];
    atomicAdd(out, t);
  }
}

// applyVOF! per-cell body + cleanWisp!, VOFutil.jl:15-37,127-136
template <class T, int D> __global__ void applyvof_kernel(T* f, T* al, T* nh, const T* sc, const T* sp, const T* sm, const Geo g, T tol) {
  const int x = 2 + blockIdx.x * blockDim.x + threadIdx.x, y = 2 + blockIdx.y, z = (D == 3) ? 2 + blockIdx.z : 1;
  if (x > g.n[0] - 1) return;
  const long long l = lin3(g, x, y, z);
  T n[3] = {T(0), T(0), T(0)}, sumN = T(0), sumN2 = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const T dd = sp[(long long)i * g.S + l] - sm[(long long)i * g.S + l];
    nh[(long long)i * g.S + l] = dd;
    n[i] = dd;
    sumN += dd;
    sumN2 += dd * dd;
  }
  const T a = sumN / T(2) - t_sqrt(sumN2) * sc[l];
  al[l] = a;
  T v = get_volume_fraction<T, D>(n, a);
  v = (v < tol) ? T(0) : ((v > T(1) - tol) ? T(1) : v);
  f[l] = v;
}

// ------------------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------------------
static inline dim3 row_grid(const Geo& g, int D, int bx) {
  return dim3((unsigned)((g.n[0] - 2 + bx - 1) / bx), (unsigned)(g.n[1] - 2), (unsigned)(D == 3 ? g.n[2] - 2 : 1));
}

template <class T, bool MOM> static int launch_sweep(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q) {
  if (c->D == 2) return launch_sweep_dim<T, 2, MOM>(c, st, q);
  return launch_sweep_dim<T, 3, MOM>(c, st, q);
}

template <class T> static int launch_bcf(ifadv_ctx* c, cudaStream_t st, T* f, unsigned per) {
  Geo g = c->g;
  g.per = per;
  const long long n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const long long tot = 2 * n1 * n2 + 2 * n0 * n2 + (c->D == 3 ? 2 * n0 * n1 : 0);
  const int bs = 256;
  const unsigned nb = (unsigned)((tot + bs - 1) / bs);
  if (c->D == 2) bcf_kernel<T, 2><<<nb, bs, 0, st>>>(f, g);
  else bcf_kernel<T, 3><<<nb, bs, 0, st>>>(f, g);
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

static int check_common(ifadv_ctx* c, int ns, const int* dirO) {
  if (!c) return -2;
  if (ns < 0 || ns > 8) return fail(c, -2, "invalid normal_scheme");
  unsigned seen = 0;
  for (int k = 0; k < c->D; ++k) {
    if (dirO[k] < 1 || dirO[k] > c->D) return fail(c, -2, "dirO entries must be in 1..D");
    seen |= 1u << (dirO[k] - 1);
  }
  if (seen != (1u << c->D) - 1u) return fail(c, -2, "dirO must be a permutation of 1..D");
  return 0;
}

// decode the per-sweep reductions into status + report (reportFillError semantics, advection.jl:145-189)
static double key_to_double(unsigned long long k) {
  long long b = (k & 0x8000000000000000ull) ? (long long)(k & 0x7fffffffffffffffull) : (long long)(~k);
  double d;
  memcpy(&d, &b, 8);
  return d;
}
static int decode_report(ifadv_ctx* c, const int* dirO, double filltol, ifadv_report* rep) {
  int status = 0;
  rep->status = 0; rep->dir = -1; rep->maxf = 0; rep->minf = 0;
  for (int k = 0; k < 3; ++k) rep->argmax[k] = rep->argmin[k] = 0;
  for (int s = 0; s < c->D; ++s) {
    const unsigned long long* r = c->red_host + 8 * s;
    const double mx = key_to_double(r[0]), mn = key_to_double(r[1]);
    int st = 0;
    if (r[4] != 0 || mx != mx || mn != mn) st = -1;
    else {
      if (mx - 1 > filltol) st |= 1;
      if (mn < -filltol) st |= 2;
    }
    if (st != 0 || rep->status == 0) {
      rep->maxf = mx; rep->minf = mn; rep->dir = dirO[s];
      const unsigned long long am = r[2] & 0xffffffffull, an = r[3] & 0xffffffffull;
      rep->argmax[0] = (int64_t)(am % c->g.n[0]) + 1; rep->argmax[1] = (int64_t)((am / c->g.n[0]) % c->g.n[1]) + 1;
      rep->argmax[2] = (int64_t)(am / ((unsigned long long)c->g.n[0] * c->g.n[1])) + 1;
      rep->argmin[0] = (int64_t)(an % c->g.n[0]) + 1; rep->argmin[1] = (int64_t)((an / c->g.n[0]) % c->g.n[1]) + 1;
      rep->argmin[2] = (int64_t)(an / ((unsigned long long)c->g.n[0] * c->g.n[1])) + 1;
    }
    if (st < 0) { rep->status = st; return st; }
    status |= st;
    rep->status = status;
  }
  return status;
}

// The reporting tail of a call (reportFillError, advection.jl:145-189): fetch the per-sweep extrema (stream sync), decode them, and for an
// over/under-filled cell evaluate |∇·u⁰| + |∇·u| there -- the reference aborts with error("divergence, …, is exploding!") when that
// exceeds 10 or is NaN (:160,180) and only prints otherwise.  A NaN left by an earlier call that ran without a report is fatal here too.
template <class T>
static int finish_report(ifadv_ctx* c, cudaStream_t st, const int* dirO, double filltol, ifadv_report* rep, const T* u, const T* u0) {
  CU_CHECK(c, cudaMemcpyAsync(c->red_host, c->red_dev, sizeof(unsigned long long) * 32, cudaMemcpyDeviceToHost, st));
  CU_CHECK(c, cudaStreamSynchronize(st));
  rep->div_u0 = rep->div_u = 0.0;
  int status = decode_report(c, dirO, filltol, rep);
  if (status >= 0 && c->red_host[24] != 0ull) {
    CU_CHECK(c, cudaMemsetAsync(c->red_dev + 24, 0, sizeof(unsigned long long), st));
    rep->status = -1;
    return fail(c, -1, "NaN in f during an earlier call that ran without a report");
  }
  if (status < 0) return fail(c, status, "NaN in f");
  for (int which = 0; which < 2 && status > 0; ++which) {  // the max cell (:150-167), then the min cell (:169-186)
    if (!(status & (1 << which))) continue;
    const int64_t* I = which == 0 ? rep->argmax : rep->argmin;
    bool edge = false;
    for (int a = 0; a < c->D; ++a) edge = edge || I[a] < 1 || I[a] >= c->g.n[a];
    if (edge) continue;
    const long long l = (long long)(I[0] - 1) + c->g.s1 * (I[1] - 1) + c->g.s2 * ((c->D == 3 ? I[2] : 1) - 1);
    T d0 = T(0), d1 = T(0);
    for (int a = 0; a < c->D; ++a) {
      const long long sa = (a == 0) ? 1 : ((a == 1) ? c->g.s1 : c->g.s2);
      T v[4];
      CU_CHECK(c, cudaMemcpyAsync(&v[0], u0 + (long long)a * c->g.S + l, sizeof(T), cudaMemcpyDeviceToHost, st));
      CU_CHECK(c, cudaMemcpyAsync(&v[1], u0 + (long long)a * c->g.S + l + sa, sizeof(T), cudaMemcpyDeviceToHost, st));
      CU_CHECK(c, cudaMemcpyAsync(&v[2], u + (long long)a * c->g.S + l, sizeof(T), cudaMemcpyDeviceToHost, st));
      CU_CHECK(c, cudaMemcpyAsync(&v[3], u + (long long)a * c->g.S + l + sa, sizeof(T), cudaMemcpyDeviceToHost, st));
      CU_CHECK(c, cudaStreamSynchronize(st));
      d0 += v[1] - v[0];  // div(I,u⁰) = Σ ∂(a,I,u⁰)
      d1 += v[3] - v[2];
    }
    rep->div_u0 = std::fabs((double)d0);
    rep->div_u = std::fabs((double)d1);
    const double s = rep->div_u0 + rep->div_u;
    if (s > 10.0 || s != s) {
      rep->status = -5;
      char msg[96];
      snprintf(msg, sizeof msg, "divergence, %g, is exploding!", s);
      return fail(c, -5, msg);
    }
  }
  return status;
}

// ------------------------------------------------------------------------------------------------------------
// typed drivers
// ------------------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------------------
// z-slab decomposition: NCCL ghost-plane exchange (new functionality, no reference counterpart: SURVEY.md §8e)
// ------------------------------------------------------------------------------------------------------------
// NCCL is resolved at run time (the copy the process has already loaded -- e.g. the one bundled with PyTorch or NCCL.jl -- else the
// system libnccl.so.2), so the single-GPU product has no link-time dependency on it.
namespace {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};
NcclApi* nccl_api() {
  static NcclApi api = [] {
    NcclApi a{};
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.GroupStart = (decltype(a.GroupStart))dlsym(h, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(h, "ncclGroupEnd");
    a.Send = (decltype(a.Send))dlsym(h, "ncclSend");
    a.Recv = (decltype(a.Recv))dlsym(h, "ncclRecv");
    a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Send && a.Recv && a.AllReduce;
    return a;
  }();
  return &api;
}
#define NCCL_CHECK(ctx, call)                                                                                  \
  do {                                                                                                         \
    ncclResult_t r_ = (call);                                                                                  \
    if (r_ != ncclSuccess) {                                                                                   \
      if (ctx) (ctx)->err = std::string(#call) + ": " + (nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r_) : "NCCL error"); \
      return -4;                                                                                               \
    }                                                                                                          \
  } while (0)

// ---- flag kernels of the peer-to-peer path -------------------------------------------------------------------------------------
__global__ void p2p_write_kernel(unsigned* a, unsigned* b, unsigned v) {
  __threadfence_system();
  if (a) *reinterpret_cast<volatile unsigned*>(a) = v;
  if (b) *reinterpret_cast<volatile unsigned*>(b) = v;
  __threadfence_system();
}
__global__ void p2p_wait_kernel(const unsigned* a, const unsigned* b, unsigned v, unsigned* err) {
  // bounded spin (about 4 s): a mismatch between the ranks must surface as an error, never as a hung GPU
  const long long t0 = clock64();
  const volatile unsigned* va = a;
  const volatile unsigned* vb = b;
  while ((va && (int)(*va - v) < 0) || (vb && (int)(*vb - v) < 0)) {
    if (clock64() - t0 > 8000000000ll) { atomicExch(err, 1u); break; }
    __nanosleep(200);
  }
  __threadfence_system();
}

struct XItem {
  void* base;
  size_t esz;
  int ncomp, nup, ndn;
};

// Ghost planes of a batch of fields (component stride S elements): my top `nup` owned planes -> the upper neighbour's lower ghost
// planes, my bottom `ndn` owned planes -> the lower neighbour's upper ghost planes.  A sweep reaches 3 planes below / 2 above a
// cell for f and 2 / 2 for ρu, so the inner calls send fewer than G planes where that suffices.
// Peer-to-peer path (default): copy engines push over NVLink into the receiver's staging buffer, flags in the receiver's memory
// order sender and receiver on the streams (no host synchronisation, no SM-resident communication kernel):
//   ready(seq) -> neighbours;  wait ready(seq) from neighbours;  push;  done(seq) -> neighbours;  wait done(seq);  unpack.
// NCCL path (IFADV_SLAB_P2P=0, or when IPC is not available): ncclSend/ncclRecv in ONE group; sends go [up, down] and receives
// [from below, from above] per component, so the pairs match when both neighbours are one peer.
int slab_exchange_batch(ifadv_ctx* c, cudaStream_t st, const XItem* items, int n) {
  const ifadv_slab& sl = c->slab;
  if (sl.nranks <= 1) return 0;
  const size_t s2 = (size_t)c->g.s2, S = (size_t)c->g.S;
  auto clampn = [&](int v) { return (v < 0 || v > sl.G) ? sl.G : v; };
  if (c->p2p.on) {
    size_t need_up = 0, need_dn = 0;
    for (int i = 0; i < n; ++i) {
      need_up += items[i].esz * s2 * (size_t)clampn(items[i].nup) * items[i].ncomp;
      need_dn += items[i].esz * s2 * (size_t)clampn(items[i].ndn) * items[i].ncomp;
    }
    if (need_up <= c->p2p.cap && need_dn <= c->p2p.cap) {
      ifadv_p2p& p = c->p2p;
      const unsigned seq = ++p.seq;
      unsigned* err = p.flags + 8;
      const bool lo = sl.lower >= 0, up = sl.upper >= 0;
      // I am the lower neighbour's UPPER neighbour (its flags [1], [3]) and the upper neighbour's LOWER neighbour (its [0], [2])
      p2p_write_kernel<<<1, 1, 0, st>>>(lo ? p.peer_flags[0] + 1 : nullptr, up ? p.peer_flags[1] + 0 : nullptr, seq);
      p2p_wait_kernel<<<1, 1, 0, st>>>(lo ? p.flags + 0 : nullptr, up ? p.flags + 1 : nullptr, seq, err);
      size_t offU = 0, offD = 0;
      for (int i = 0; i < n; ++i) {
        const size_t esz = items[i].esz, pl = esz * s2, comp = esz * S;
        const int nup = clampn(items[i].nup), ndn = clampn(items[i].ndn);
        for (int k = 0; k < items[i].ncomp; ++k) {
          const char* b = (const char*)items[i].base + comp * k;  // 0-based storage plane q holds 1-based plane q+1
          if (up) { CU_CHECK(c, cudaMemcpyAsync(p.peer_stage[1] + offU, b + pl * (size_t)(c->kz1 - 1 - nup), pl * nup, cudaMemcpyDeviceToDevice, st)); offU += pl * nup; }
          if (lo) { CU_CHECK(c, cudaMemcpyAsync(p.peer_stage[0] + offD, b + pl * (size_t)(c->kz0 - 1), pl * ndn, cudaMemcpyDeviceToDevice, st)); offD += pl * ndn; }
          c->slab.bytes_sent += (long long)pl * ((up ? nup : 0) + (lo ? ndn : 0));
        }
      }
      p2p_write_kernel<<<1, 1, 0, st>>>(lo ? p.peer_flags[0] + 3 : nullptr, up ? p.peer_flags[1] + 2 : nullptr, seq);
      p2p_wait_kernel<<<1, 1, 0, st>>>(lo ? p.flags + 2 : nullptr, up ? p.flags + 3 : nullptr, seq, err);
      size_t offL = 0, offH = 0;  // my stage[0] holds the lower neighbour's up-push (nup planes), stage[1] the upper's down-push (ndn)
      for (int i = 0; i < n; ++i) {
        const size_t esz = items[i].esz, pl = esz * s2, comp = esz * S;
        const int nup = clampn(items[i].nup), ndn = clampn(items[i].ndn);
        for (int k = 0; k < items[i].ncomp; ++k) {
          char* b = (char*)items[i].base + comp * k;
          if (lo) { CU_CHECK(c, cudaMemcpyAsync(b + pl * (size_t)(c->kz0 - 1 - nup), p.stage[0] + offL, pl * nup, cudaMemcpyDeviceToDevice, st)); offL += pl * nup; }
          if (up) { CU_CHECK(c, cudaMemcpyAsync(b + pl * (size_t)(c->kz1 - 1), p.stage[1] + offH, pl * ndn, cudaMemcpyDeviceToDevice, st)); offH += pl * ndn; }
        }
      }
      c->launches += 4;
      CU_CHECK(c, cudaGetLastError());
      return 0;
    }
  }
  NcclApi* A = nccl_api();
  if (!A->ok) return fail(c, -4, "NCCL is not available");
  ncclComm_t comm = (ncclComm_t)sl.comm;
  NCCL_CHECK(c, A->GroupStart());
  for (int i = 0; i < n; ++i) {
    const size_t esz = items[i].esz, pl = esz * s2, comp = esz * S;
    const int nup = clampn(items[i].nup), ndn = clampn(items[i].ndn);
    for (int k = 0; k < items[i].ncomp; ++k) {
      char* b = (char*)items[i].base + comp * k;
      if (sl.upper >= 0) NCCL_CHECK(c, A->Send(b + pl * (size_t)(c->kz1 - 1 - nup), pl * (size_t)nup, ncclInt8, sl.upper, comm, st));
      if (sl.lower >= 0) NCCL_CHECK(c, A->Send(b + pl * (size_t)(c->kz0 - 1), pl * (size_t)ndn, ncclInt8, sl.lower, comm, st));
      if (sl.lower >= 0) NCCL_CHECK(c, A->Recv(b + pl * (size_t)(c->kz0 - 1 - nup), pl * (size_t)nup, ncclInt8, sl.lower, comm, st));
      if (sl.upper >= 0) NCCL_CHECK(c, A->Recv(b + pl * (size_t)(c->kz1 - 1), pl * (size_t)ndn, ncclInt8, sl.upper, comm, st));
      c->slab.bytes_sent += (long long)pl * ((sl.upper >= 0 ? nup : 0) + (sl.lower >= 0 ? ndn : 0));
    }
  }
  NCCL_CHECK(c, A->GroupEnd());
  return 0;
}
int slab_exchange(ifadv_ctx* c, cudaStream_t st, void* base, size_t esz, int ncomp, int nup = -1, int ndn = -1) {
  XItem it{base, esz, ncomp, nup, ndn};
  return slab_exchange_batch(c, st, &it, 1);
}

// Set up the peer-to-peer path of a slab context: staging buffers + flag block, IPC handles swapped with both neighbours through the
// NCCL communicator, all ranks agree (all-reduce) whether every mapping succeeded; on any failure everybody stays on NCCL.
struct P2PHello {
  cudaIpcMemHandle_t h[3];  // flags, stage[0], stage[1]
  int ok;
};
int slab_p2p_setup(ifadv_ctx* c) {
  ifadv_p2p& p = c->p2p;
  const ifadv_slab& sl = c->slab;
  NcclApi* A = nccl_api();
  if (!A->ok) return 0;
  const char* e = getenv("IFADV_SLAB_P2P");
  int want = !(e && atoi(e) == 0);
  const size_t esz = esize(c->dtype);
  p.cap = (size_t)sl.G * (size_t)c->g.s2 * (4 * esz + 1);  // f + 3 components of ρu (or u) + c̄, G planes each
  P2PHello mine{};
  mine.ok = 0;
  if (want && cudaMalloc(&p.flags, 64) == cudaSuccess && cudaMemset(p.flags, 0, 64) == cudaSuccess &&
      cudaMalloc(&p.stage[0], p.cap) == cudaSuccess && cudaMalloc(&p.stage[1], p.cap) == cudaSuccess &&
      cudaIpcGetMemHandle(&mine.h[0], p.flags) == cudaSuccess && cudaIpcGetMemHandle(&mine.h[1], p.stage[0]) == cudaSuccess &&
      cudaIpcGetMemHandle(&mine.h[2], p.stage[1]) == cudaSuccess)
    mine.ok = 1;
  cudaGetLastError();
  // swap hellos with the neighbours
  P2PHello* dev = nullptr;  // [0] mine, [1] from lower, [2] from upper
  P2PHello got[3];
  memset(got, 0, sizeof got);
  cudaStream_t st = nullptr;
  CU_CHECK(c, cudaMalloc(&dev, 3 * sizeof(P2PHello)));
  CU_CHECK(c, cudaMemset(dev, 0, 3 * sizeof(P2PHello)));
  CU_CHECK(c, cudaMemcpy(dev, &mine, sizeof mine, cudaMemcpyHostToDevice));
  ncclComm_t comm = (ncclComm_t)sl.comm;
  NCCL_CHECK(c, A->GroupStart());
  if (sl.upper >= 0) NCCL_CHECK(c, A->Send(dev, sizeof(P2PHello), ncclInt8, sl.upper, comm, st));
  if (sl.lower >= 0) NCCL_CHECK(c, A->Send(dev, sizeof(P2PHello), ncclInt8, sl.lower, comm, st));
  if (sl.lower >= 0) NCCL_CHECK(c, A->Recv(dev + 1, sizeof(P2PHello), ncclInt8, sl.lower, comm, st));
  if (sl.upper >= 0) NCCL_CHECK(c, A->Recv(dev + 2, sizeof(P2PHello), ncclInt8, sl.upper, comm, st));
  NCCL_CHECK(c, A->GroupEnd());
  CU_CHECK(c, cudaStreamSynchronize(st));
  CU_CHECK(c, cudaMemcpy(got, dev, sizeof got, cudaMemcpyDeviceToHost));
  int ok = mine.ok;
  for (auto& q : p.opened) q = nullptr;
  auto open = [&](const cudaIpcMemHandle_t& h, int slot) -> void* {
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; return nullptr; }
    p.opened[slot] = ptr;
    return ptr;
  };
  if (ok && sl.lower >= 0) {
    if (!got[1].ok) ok = 0;
    else {
      p.peer_flags[0] = (unsigned*)open(got[1].h[0], 0);
      p.peer_stage[0] = (char*)open(got[1].h[2], 1);  // the lower neighbour's stage[1]: filled by ITS upper neighbour, me
    }
  }
  if (ok && sl.upper >= 0) {
    if (!got[2].ok) ok = 0;
    else if (sl.upper == sl.lower) {  // two ranks, periodic z: one peer in both roles -- a handle is opened once per process
      p.peer_flags[1] = p.peer_flags[0];
      p.peer_stage[1] = (char*)open(got[2].h[1], 2);
    } else {
      p.peer_flags[1] = (unsigned*)open(got[2].h[0], 3);
      p.peer_stage[1] = (char*)open(got[2].h[1], 4);  // the upper neighbour's stage[0]: filled by ITS lower neighbour, me
    }
  }
  // everybody or nobody
  int* okd = (int*)dev;
  CU_CHECK(c, cudaMemcpy(okd, &ok, sizeof(int), cudaMemcpyHostToDevice));
  NCCL_CHECK(c, A->AllReduce(okd, okd, 1, ncclInt32, ncclMin, comm, st));
  CU_CHECK(c, cudaStreamSynchronize(st));
  CU_CHECK(c, cudaMemcpy(&ok, okd, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(dev);
  p.on = ok;
  p.seq = 0;
  return 0;
}
void slab_p2p_free(ifadv_ctx* c) {
  ifadv_p2p& p = c->p2p;
  for (auto& q : p.opened) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
  if (p.flags) cudaFree(p.flags);
  if (p.stage[0]) cudaFree(p.stage[0]);
  if (p.stage[1]) cudaFree(p.stage[1]);
  p.flags = nullptr; p.stage[0] = p.stage[1] = nullptr; p.on = 0;
}
}  // namespace

template <class T>
static int advect_vof_t(ifadv_ctx* c, cudaStream_t st, T* f, T* ff, T* al, const T* u, const T* u0, double dt, int8_t* cbar, T* rhouf,
                        double lr, int ns, unsigned per, const int* dirO, int flags, ifadv_report* rep) {
  const int D = c->D;
  c->g.per = per;
  const bool want_rhouf = !(flags & IFADV_NO_RHOUF) && rhouf != nullptr;
  if (D == 2 && c->use_march != 0) {
    // 2-D grids (BASELINE config 1): the whole step -- slot reset, fill!(ρuf,0), both cell-parallel sweeps, BCf! -- in one cooperative launch
    SweepCfg<T> q[2];
    T* bufs2[3] = {f, ff, f};
    for (int s = 0; s < 2; ++s) {
      q[s] = SweepCfg<T>{};
      q[s].f_in = bufs2[s]; q[s].f_out = bufs2[s + 1];
      q[s].u = u; q[s].u0 = u0; q[s].cbar = cbar; q[s].rhouf = want_rhouf ? rhouf : nullptr;
      q[s].dt = dt; q[s].lr = lr; q[s].scheme = ns; q[s].lim = 0; q[s].first = (s == 0); q[s].j = dirO[s] - 1;
      q[s].red = c->red_dev + 8 * s;
    }
    int rc = launch_vof2d_step<T>(c, st, q[0], q[1], f, want_rhouf ? rhouf : nullptr);
    if (rc) return rc;
    if (rep) return finish_report<T>(c, st, dirO, 10.0 * (double)std::numeric_limits<T>::epsilon(), rep, u, u0);
    return 0;
  }
  red_init_kernel<<<1, 32, 0, st>>>(c->red_dev, 3);
  c->launches++;
  if (want_rhouf) CU_CHECK(c, cudaMemsetAsync(rhouf, 0, sizeof(T) * c->g.S * D, st));  // fill!(ρuf,0), advection.jl:37
  T* bufs[4] = {f, ff, (D == 3) ? al : f, f};  // f -> fᶠ -> α -> f (3-D);  f -> fᶠ -> f (2-D)
  for (int s = 0; s < D; ++s) {
    SweepCfg<T> q{};
    q.f_in = bufs[s]; q.f_out = bufs[s + 1];
    q.u = u; q.u0 = u0; q.cbar = cbar; q.rhouf = want_rhouf ? rhouf : nullptr;
    q.dt = dt; q.lr = lr; q.scheme = ns; q.lim = 0; q.first = (s == 0); q.j = dirO[s] - 1;
    q.red = c->red_dev + 8 * s;
    int rc = launch_sweep<T, false>(c, st, q);
    if (rc) return rc;
  }
  int rc = launch_bcf<T>(c, st, f, per);  // BCf!(f;perdir), advection.jl:72 (only the final ghosts are observable)
  if (rc) return rc;
  if (rep) return finish_report<T>(c, st, dirO, 10.0 * (double)std::numeric_limits<T>::epsilon(), rep, u, u0);  // tol, advection.jl:69
  return 0;
}

template <class T> static int u2rhou_t(ifadv_ctx* c, cudaStream_t st, T* out, const T* in, const T* f, double lr, bool to_rhou);
template <class T> static int bcvec_t(ifadv_ctx* c, cudaStream_t st, T* a, const double* A, int saveexit, unsigned per);

// f_src / fused: the fused entry (ifadv_u2rhou_advect_vof_rhouu).  In 3-D the first sweep reads f_src and forms
// ρu = BC!(uOld*ρ(f̄)) on the fly; in 2-D the three reference calls are simply issued back to back.
template <class T>
static int advect_vof_rhouu_t(ifadv_ctx* c, cudaStream_t st, T* f, T* ff, T* Phi, const T* u, const T* u0, double dt, int8_t* cbar, T* rhou,
                              T* r, T* rhouf, const T* uOld, const T* drho, double lr, int lim, int ns, const double* uBC, unsigned per,
                              const int* dirO, ifadv_report* rep, cudaEvent_t wait_f, int exitBC, const T* f_src = nullptr, int fused = 0) {
  const int D = c->D;
  c->g.per = per;
  // exitBC (BC!'s saveexit at flow.jl:197,207): plane N of component x of u★ keeps the value the caller's r array holds there and the
  // exit face keeps its computed mass flux.  The lean kernels never write a ghost plane of r, so every sweep reads the entry value.
  if (exitBC && (per & 1u)) exitBC = 0;  // no boundary planes in a periodic direction
  if (exitBC && c->use_march != 0 && D == 3 && !c->use_along2) return fail(c, -2, "exitBC needs the lean kernels or the tile kernel (IFADV_KERNEL unset or tile)");
  // wait_f (ifadv_defer_f_writes_until): one-shot event this call waits for before its first write to f
  if (fused && (D == 2 || c->use_march == 0)) {
    if (f_src && f_src != f) {
      if (wait_f) { CU_CHECK(c, cudaStreamWaitEvent(st, wait_f, 0)); wait_f = nullptr; }
      CU_CHECK(c, cudaMemcpyAsync(f, f_src, sizeof(T) * c->g.S, cudaMemcpyDeviceToDevice, st));
    }
    int rc0 = u2rhou_t<T>(c, st, rhou, uOld, f, lr, true);
    if (rc0) return rc0;
    if ((rc0 = bcvec_t<T>(c, st, rhou, uBC, exitBC, per))) return rc0;
    f_src = nullptr;
    fused = 0;
  }
  red_init_kernel<<<1, 32, 0, st>>>(c->red_dev, 3);
  c->launches++;
  T* fb[4] = {(f_src && fused) ? const_cast<T*>(f_src) : f, ff, (D == 3) ? Phi : f, f};  // f -> fᶠ -> Φ -> f
  T* rb[4] = {rhou, r, (D == 3) ? rhouf : rhou, rhou};  // ρu -> r -> ρuf -> ρu
  for (int s = 0; s < D; ++s) {
    SweepCfg<T> q{};
    q.f_in = fb[s]; q.f_out = fb[s + 1];
    q.rhou_in = rb[s]; q.rhou_out = rb[s + 1];
    q.u = u; q.u0 = u0; q.uOld = uOld; q.drho = drho; q.cbar = cbar; q.rhouf = nullptr; q.uexit = exitBC ? r : nullptr;
    q.dt = dt; q.lr = lr; q.scheme = ns; q.lim = lim; q.first = (s == 0); q.j = dirO[s] - 1;
    q.fused = (s == 0) ? fused : 0;
    for (int i = 0; i < 3; ++i) q.A[i] = (i < D) ? uBC[i] : 0.0;
    q.red = c->red_dev + 8 * s;
    if (s == D - 1 && wait_f) CU_CHECK(c, cudaStreamWaitEvent(st, wait_f, 0));  // the last sweep is the first to write f
    int rc;
    if (c->slab.nranks <= 1) {
      if ((rc = launch_sweep<T, true>(c, st, q))) return rc;
      continue;
    }
    // z-slab: the next sweep reads the neighbours' planes of what this sweep produced (stencil reach 3 below / 2 above).
    // Overlap: the G boundary planes of each slab end are swept FIRST (they need the ghost planes the previous exchange delivers and
    // produce the planes this sweep's exchange sends); the exchange then runs on the context's second stream underneath the sweep
    // of the interior planes, which reads owned planes only.
    const int G = c->slab.G, z0 = c->kz0, z1 = c->kz1;
    const bool split = c->slab.overlap && (z1 - z0) >= 2 * G + 4;
    auto sweep_planes = [&](int a, int b, bool timed) -> int {
      if (b <= a) return 0;
      const int p0 = c->prof_on;
      c->kz0 = a; c->kz1 = b;
      if (!timed) c->prof_on = 0;
      const int r = launch_sweep<T, true>(c, st, q);
      c->kz0 = z0; c->kz1 = z1; c->prof_on = p0;
      return r;
    };
    auto exchange_outputs = [&](cudaStream_t xs) -> int {
      // everything the next sweep needs from the neighbours in one batch: f 3 planes up / 2 down, ρu 2 / 2, c̄ 3 / 2 (once per call)
      XItem it[3] = {{fb[s + 1], sizeof(T), 1, 3, 2}, {rb[s + 1], sizeof(T), D, 2, 2}, {cbar, 1, 1, 3, 2}};
      return slab_exchange_batch(c, xs, it, s == 0 ? 3 : 2);
    };
    if (!split) {
      if ((rc = sweep_planes(z0, z1, true))) return rc;
      if (s < D - 1 && (rc = exchange_outputs(st))) return rc;
      continue;
    }
    if (s > 0) CU_CHECK(c, cudaStreamWaitEvent(st, c->slab_ev[1], 0));  // ghost planes of this sweep's inputs have arrived
    if ((rc = sweep_planes(z0, z0 + G, false))) return rc;
    if ((rc = sweep_planes(z1 - G, z1, false))) return rc;
    if (s < D - 1) {
      CU_CHECK(c, cudaEventRecord(c->slab_ev[0], st));
      CU_CHECK(c, cudaStreamWaitEvent(c->slab_stream, c->slab_ev[0], 0));
      if ((rc = exchange_outputs(c->slab_stream))) return rc;
      CU_CHECK(c, cudaEventRecord(c->slab_ev[1], c->slab_stream));
    }
    if ((rc = sweep_planes(z0 + G, z1 - G, true))) return rc;
  }
  int rc = launch_bcf<T>(c, st, f, per);
  if (rc) return rc;
  if (c->slab.nranks > 1 && (rc = slab_exchange(c, st, f, sizeof(T), 1))) return rc;  // the caller's f leaves with valid ghost planes
  if (rep) return finish_report<T>(c, st, dirO, 100.0 * (double)std::numeric_limits<T>::epsilon(), rep, u, u0);  // 10tol, advection.jl:85
  return 0;
}

template <class T> static int u2rhou_t(ifadv_ctx* c, cudaStream_t st, T* out, const T* in, const T* f, double lr, bool to_rhou) {
  const int bx = 128;
  dim3 grid = row_grid(c->g, c->D, bx);
  const T l = (T)lr, om = T(1) - l;
  if (c->D == 2) {
    if (to_rhou) urhou_kernel<T, 2, true><<<grid, bx, 0, st>>>(out, in, f, c->g, l, om);
    else urhou_kernel<T, 2, false><<<grid, bx, 0, st>>>(out, in, f, c->g, l, om);
  } else {
    if (to_rhou) urhou_kernel<T, 3, true><<<grid, bx, 0, st>>>(out, in, f, c->g, l, om);
    else urhou_kernel<T, 3, false><<<grid, bx, 0, st>>>(out, in, f, c->g, l, om);
  }
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

template <class T> static int bcvec_t(ifadv_ctx* c, cudaStream_t st, T* a, const double* A, int saveexit, unsigned per) {
  Geo g = c->g;
  g.per = per;
  const long long n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const long long tot = 3 * n1 * n2 + 3 * n0 * n2 + (c->D == 3 ? 3 * n0 * n1 : 0);
  const int bs = 256;
  const unsigned nb = (unsigned)((tot + bs - 1) / bs);
  if (c->D == 2) bcvec_kernel<T, 2><<<nb, bs, 0, st>>>(a, g, (T)A[0], (T)A[1], T(0), saveexit);
  else bcvec_kernel<T, 3><<<nb, bs, 0, st>>>(a, g, (T)A[0], (T)A[1], (T)A[2], saveexit);
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

template <class T> static int axpby_t(ifadv_ctx* c, cudaStream_t st, T* out, double a, const T* x, double b, const T* y) {
  const long long n = c->g.S;
  const int bs = 256;
  const bool vec = (((uintptr_t)out | (uintptr_t)x | (uintptr_t)y) & 15u) == 0;
  const long long work = vec ? n / (long long)(16 / sizeof(T)) : n;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((work + bs - 1) / bs, 148LL * 16));
  if (vec) axpby_kernel<T, true><<<grid, bs, 0, st>>>(out, (T)a, x, (T)b, y, n, a == b);
  else axpby_kernel<T, false><<<grid, bs, 0, st>>>(out, (T)a, x, (T)b, y, n, a == b);
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

// ---- explicit forcing (ifadv_forcing.cuh) ------------------------------------------------------------------------------------
static inline dim3 all_grid(const Geo& g, int D, int bx) {
  return dim3((unsigned)((g.n[0] + bx - 1) / bx), (unsigned)g.n[1], (unsigned)(D == 3 ? g.n[2] : 1));
}
template <class T>
static int visc_surften_t(ifadv_ctx* c, cudaStream_t st, T* r, const T* u, const T* f, const T* nhat, const T* fb, double lmu, double mu,
                          double lr, double eta, unsigned per) {
  Geo g = c->g;
  g.per = per;
  const int bx = 128;
  const dim3 gi = row_grid(g, c->D, bx);
  const int has_mu = mu > 0.0, has_eta = eta > 0.0;
  if (c->D == 2) visc_kernel<T, 2><<<gi, bx, 0, st>>>(r, u, f, nhat, g, (T)lmu, (T)mu, (T)lr, has_mu);
  else if (!has_mu) visc_kernel<T, 3><<<gi, bx, 0, st>>>(r, u, f, nhat, g, (T)lmu, (T)mu, (T)lr, 0);
  else {
    auto kern = visc3m_kernel<T>;
    const size_t smem = ViscM::bytes<T>();
    static unsigned long long attr_devs = 0ull;
    if (!((attr_devs >> (c->device & 63)) & 1ull)) {
      CU_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_devs |= 1ull << (c->device & 63);
    }
    const int nz = g.n[2] - 2;
    const long long tiles = (long long)((g.n[0] - 2 + ViscM::TX - 1) / ViscM::TX) * ((g.n[1] - 2 + ViscM::TY - 1) / ViscM::TY);
    int chunk = 64;  // one extra staged plane per chunk
    while (chunk > 8 && tiles * ((nz + chunk - 1) / chunk) < 148 * 16) chunk >>= 1;
    const dim3 gt((unsigned)((g.n[0] - 2 + ViscM::TX - 1) / ViscM::TX), (unsigned)((g.n[1] - 2 + ViscM::TY - 1) / ViscM::TY),
                  (unsigned)((nz + chunk - 1) / chunk));
    kern<<<gt, ViscM::NT, smem, st>>>(r, u, f, nhat, g, (T)lmu, (T)mu, (T)lr, chunk);
  }
  c->launches++;
  if (has_eta) {
    if (!c->st_list) {  // capacity per direction: an eighth of the cells (interfaces are surfaces); an overflow falls back to scanning every cell
      const unsigned cap = (unsigned)std::min<long long>(std::max<long long>(c->g.S / 8, 1 << 16), 1ll << 27);
      CU_CHECK(c, cudaMalloc(&c->st_list, sizeof(int) * (size_t)cap * 3));
      CU_CHECK(c, cudaMalloc(&c->st_cnt, sizeof(unsigned) * 4));
      c->st_cap = cap;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    CU_CHECK(c, cudaMemsetAsync(c->st_cnt, 0, sizeof(unsigned) * 4, st));
    const dim3 gs((unsigned)((g.n[0] - 2 + bx - 1) / bx), (unsigned)((g.n[1] - 2 + 3) / 4), (unsigned)(c->D == 3 ? g.n[2] - 2 : 1));
    if (c->D == 2) {
      stscan_kernel<T, 2><<<gs, bx, 0, st>>>(f, g, c->st_list, c->st_cnt, c->st_cap);
      surften_kernel<T, 2><<<(unsigned)sms * 8, 128, 0, st>>>(r, f, fb, g, (T)eta, c->st_list, c->st_cnt, c->st_cap);
    } else {
      stscan_kernel<T, 3><<<gs, bx, 0, st>>>(f, g, c->st_list, c->st_cnt, c->st_cap);
      surften_kernel<T, 3><<<(unsigned)sms * 8, 128, 0, st>>>(r, f, fb, g, (T)eta, c->st_list, c->st_cnt, c->st_cap);
    }
    c->launches += 2;
  }
  CU_CHECK(c, cudaGetLastError());
  return 0;
}
template <class T>
static int update_u_t(ifadv_ctx* c, cudaStream_t st, T* u, T* ru, const T* ru0, T* fo, double dt, const T* f, double lr, const double* grav,
                      double w) {
  const int bx = 128;
  const dim3 ga = all_grid(c->g, c->D, bx);
  const T G0 = grav ? (T)grav[0] : T(0), G1 = grav ? (T)grav[1] : T(0), G2 = (grav && c->D == 3) ? (T)grav[2] : T(0);
  if (c->D == 2) update_u_kernel<T, 2><<<ga, bx, 0, st>>>(u, ru, ru0, fo, f, c->g, (T)dt, (T)lr, (T)w, G0, G1, G2, grav != nullptr);
  else update_u_kernel<T, 3><<<ga, bx, 0, st>>>(u, ru, ru0, fo, f, c->g, (T)dt, (T)lr, (T)w, G0, G1, G2, grav != nullptr);
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}
template <class T> static int update_l_t(ifadv_ctx* c, cudaStream_t st, T* mu0, const T* f, double lr, unsigned per, int fill_one) {
  const int bx = 128;
  const dim3 gi = row_grid(c->g, c->D, bx);
  if (c->D == 2) update_l_kernel<T, 2><<<gi, bx, 0, st>>>(mu0, f, c->g, (T)lr, fill_one);
  else update_l_kernel<T, 3><<<gi, bx, 0, st>>>(mu0, f, c->g, (T)lr, fill_one);
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  const double Z[3] = {0.0, 0.0, 0.0};
  return bcvec_t<T>(c, st, mu0, Z, 0, per);  // BC!(μ₀,zeros,false,perdir), flow.jl:258
}

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
// helpers of the slab-aware pressure solver (ifadv_poisson.cu): sum all-reduce of n doubles in device memory, exchange of `planes`
// ghost planes of a scalar field with both z-neighbours
int ifadv_slab_allreduce_sum(ifadv_ctx* c, cudaStream_t st, double* dev, int n) {
  if (c->slab.nranks <= 1) return 0;
  if (!nccl_api()->ok) return fail(c, -4, "NCCL is not available");
  NCCL_CHECK(c, nccl_api()->AllReduce(dev, dev, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)c->slab.comm, st));
  return 0;
}
int ifadv_slab_exchange_scalar(ifadv_ctx* c, cudaStream_t st, void* field, size_t esz, int planes) {
  return slab_exchange(c, st, field, esz, 1, planes, planes);
}
namespace { void host_pipe_free(ifadv_ctx* c); }  // z-slab pipeline of the host-buffer entry point, defined below

extern "C" {

const char* ifadv_version(void) { return "ifadv-b200 0.1 (sm_100a)"; }

int ifadv_create(ifadv_ctx** out, int D, const int64_t Ng[3], int dtype, int device) {
  if (!out || (D != 2 && D != 3) || (dtype != IFADV_F32 && dtype != IFADV_F64)) return -2;
  for (int k = 0; k < D; ++k)
    if (Ng[k] < 3) return -2;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return -3;  // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return -3;
  ifadv_ctx* c = new ifadv_ctx();
  c->D = D; c->dtype = dtype; c->device = device; c->launches = 0;
  c->slab = ifadv_slab{nullptr, 0, 1, 0, 0, 0, -1, -1, 0, 0};
  c->slab_stream = nullptr; c->slab_ev[0] = c->slab_ev[1] = nullptr;
  memset(&c->p2p, 0, sizeof c->p2p);
  c->g.n[0] = (int)Ng[0]; c->g.n[1] = (int)Ng[1]; c->g.n[2] = (D == 3) ? (int)Ng[2] : 1;
  c->g.s1 = c->g.n[0]; c->g.s2 = (long long)c->g.n[0] * c->g.n[1];
  c->g.S = c->g.s2 * c->g.n[2];
  c->g.per = 0;
  for (int k = 0; k < 3; ++k) c->Ng[k] = c->g.n[k];
  c->kz0 = 2; c->kz1 = c->g.n[2];
  for (auto& p : c->w) p = nullptr;
  c->pin_f = c->pin_u = c->pin_ru = nullptr;
  c->own_stream = nullptr;
  c->st_list = nullptr; c->st_cnt = nullptr; c->st_cap = 0;
  c->pipe = nullptr;
  c->wait_f = nullptr;
  c->pois_ctl = c->pois_host = nullptr; c->pois_ev[0] = c->pois_ev[1] = nullptr;
  c->host_h2d = c->host_d2h = 0; c->host_slabs = 0;
  c->prof_on = 0; c->prof_n = 0; c->prof_ev = nullptr; c->prof_tag = nullptr;
  for (int k = 0; k < 8; ++k) { c->prof_dir_ms[k] = 0.0; c->prof_dir_n[k] = 0; }
  {
    const char* e = getenv("IFADV_KERNEL");
    // default: register-marching (y,z sweeps) + plane-marching (x sweep); "march": plane-marching for all; "tile": v1
    c->use_march = (e && std::string(e) == "tile") ? 0 : ((e && std::string(e) == "march") ? 2 : 1);
    c->use_along2 = 1;  // y/z sweeps: the lean register-marching kernel (its first generation, IFADV_KERNEL=along1, is retired)
    c->use_xrow = (e && std::string(e) == "xsweep") ? 0 : 1;    // "xsweep": the plane-marching kernel for CMOM x sweeps
    const char* ev = getenv("IFADV_VOF_KERNEL");
    c->use_vofcell = (e || (ev && std::string(ev) == "lean")) ? 0 : 1;  // pure VOF, 3-D: cell-parallel kernel unless an older generation is asked for
  }
  if (cudaMalloc(&c->red_dev, sizeof(unsigned long long) * 32) != cudaSuccess ||
      cudaMemset(c->red_dev, 0, sizeof(unsigned long long) * 32) != cudaSuccess ||
      cudaMallocHost(&c->red_host, sizeof(unsigned long long) * 32) != cudaSuccess ||
      cudaMalloc(&c->misc_dev, sizeof(unsigned long long) * 8) != cudaSuccess ||
      cudaMallocHost(&c->misc_host, sizeof(unsigned long long) * 8) != cudaSuccess) {
    delete c;
    return -3;
  }
  *out = c;
  return 0;
}

int ifadv_destroy(ifadv_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaFree(c->red_dev); cudaFreeHost(c->red_host); cudaFree(c->misc_dev); cudaFreeHost(c->misc_host);
  for (auto& p : c->w) if (p) cudaFree(p);
  if (c->pin_f) cudaFreeHost(c->pin_f);
  if (c->pin_u) cudaFreeHost(c->pin_u);
  if (c->pin_ru) cudaFreeHost(c->pin_ru);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->st_list) cudaFree(c->st_list);
  if (c->st_cnt) cudaFree(c->st_cnt);
  slab_p2p_free(c);
  if (c->slab_stream) { cudaStreamDestroy(c->slab_stream); cudaEventDestroy(c->slab_ev[0]); cudaEventDestroy(c->slab_ev[1]); }
  host_pipe_free(c);
  ifadv_poisson_free(c);
  if (c->prof_ev) { for (int k = 0; k < 2 * IFADV_PROF_MAX; ++k) cudaEventDestroy(c->prof_ev[k]); delete[] c->prof_ev; delete[] c->prof_tag; }
  delete c;
  return 0;
}

int ifadv_profile(ifadv_ctx* c, int enable) {
  if (!c) return -2;
  if (enable && !c->prof_ev) {
    c->prof_ev = new cudaEvent_t[2 * IFADV_PROF_MAX];
    c->prof_tag = new unsigned char[IFADV_PROF_MAX];
    for (int k = 0; k < 2 * IFADV_PROF_MAX; ++k) CU_CHECK(c, cudaEventCreate(&c->prof_ev[k]));
  }
  c->prof_on = enable ? 1 : 0;
  if (enable) c->prof_n = 0;
  return 0;
}
int ifadv_profile_read(ifadv_ctx* c, double* total_ms, int64_t* launches) {
  if (!c || !total_ms || !launches) return -2;
  double tot[2] = {0.0, 0.0};
  int64_t cnt[2] = {0, 0};
  for (int k = 0; k < c->prof_n; ++k) {
    float ms = 0.f;
    CU_CHECK(c, cudaEventSynchronize(c->prof_ev[2 * k + 1]));
    CU_CHECK(c, cudaEventElapsedTime(&ms, c->prof_ev[2 * k], c->prof_ev[2 * k + 1]));
    tot[(c->prof_tag[k] & 1) ? 1 : 0] += ms;
    cnt[(c->prof_tag[k] & 1) ? 1 : 0]++;
    c->prof_dir_ms[(c->prof_tag[k] >> 1) & 7] += ms;
    c->prof_dir_n[(c->prof_tag[k] >> 1) & 7]++;
  }
  total_ms[0] = tot[0]; total_ms[1] = tot[1];
  launches[0] = cnt[0]; launches[1] = cnt[1];
  c->prof_n = 0;
  return 0;
}
int ifadv_profile_read_dirs(ifadv_ctx* c, double* total_ms, int64_t* launches) {
  if (!c || !total_ms || !launches) return -2;
  for (int k = 0; k < 6; ++k) { total_ms[k] = c->prof_dir_ms[k]; launches[k] = c->prof_dir_n[k]; c->prof_dir_ms[k] = 0.0; c->prof_dir_n[k] = 0; }
  return 0;
}

const char* ifadv_last_error(const ifadv_ctx* c) { return c ? c->err.c_str() : "null context"; }
int64_t ifadv_launch_count(const ifadv_ctx* c) { return c ? c->launches : 0; }

int ifadv_advect_vof(ifadv_ctx* c, void* stream, void* f, void* ff, void* alpha, void* nhat, const void* u, const void* u0, double dt,
                     int8_t* cbar, void* rhouf, double lambda_rho, int normal_scheme, unsigned perdir_mask, const int dirO[3], int flags,
                     ifadv_report* report) {
  (void)nhat;
  int rc = check_common(c, normal_scheme, dirO);
  if (rc) return rc;
  if (c->slab.nranks > 1) return fail(c, -2, "z-slab contexts run the CMOM path (ifadv_advect_vof_rhouu / ifadv_u2rhou_advect_vof_rhouu)");
  if (!f || !ff || !u || !u0 || !cbar || (c->D == 3 && !alpha)) return fail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->dtype == IFADV_F32)
    return advect_vof_t<float>(c, st, (float*)f, (float*)ff, (float*)alpha, (const float*)u, (const float*)u0, dt, cbar, (float*)rhouf,
                               lambda_rho, normal_scheme, perdir_mask, dirO, flags, report);
  return advect_vof_t<double>(c, st, (double*)f, (double*)ff, (double*)alpha, (const double*)u, (const double*)u0, dt, cbar, (double*)rhouf,
                              lambda_rho, normal_scheme, perdir_mask, dirO, flags, report);
}

int ifadv_advect_vof_rhouu(ifadv_ctx* c, void* stream, void* f, void* ff, void* alpha, void* nhat, const void* u, const void* u0, double dt,
                           int8_t* cbar, void* rhou, void* r, void* Phi, void* rhouf, void* uStar, const void* uOld, void* dilaU,
                           const void* drho, double lambda_rho, int limiter, int normal_scheme, const double uBC[3], unsigned perdir_mask,
                           int exitBC, const int dirO[3], ifadv_report* report) {
  (void)alpha; (void)nhat; (void)uStar; (void)dilaU;
  cudaEvent_t wait_f = nullptr;
  if (c) { wait_f = c->wait_f; c->wait_f = nullptr; }  // one-shot, consumed even when the call is rejected below
  int rc = check_common(c, normal_scheme, dirO);
  if (rc) return rc;
  if (limiter < 0 || limiter > 10) return fail(c, -2, "invalid limiter");
  if (!f || !ff || !u || !u0 || !cbar || !rhou || !r || !uOld || !drho || !uBC || (c->D == 3 && (!Phi || !rhouf)))
    return fail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->dtype == IFADV_F32)
    return advect_vof_rhouu_t<float>(c, st, (float*)f, (float*)ff, (float*)Phi, (const float*)u, (const float*)u0, dt, cbar, (float*)rhou,
                                     (float*)r, (float*)rhouf, (const float*)uOld, (const float*)drho, lambda_rho, limiter, normal_scheme,
                                     uBC, perdir_mask, dirO, report, wait_f, exitBC);
  return advect_vof_rhouu_t<double>(c, st, (double*)f, (double*)ff, (double*)Phi, (const double*)u, (const double*)u0, dt, cbar,
                                    (double*)rhou, (double*)r, (double*)rhouf, (const double*)uOld, (const double*)drho, lambda_rho, limiter,
                                    normal_scheme, uBC, perdir_mask, dirO, report, wait_f, exitBC);
}

int ifadv_u2rhou_advect_vof_rhouu(ifadv_ctx* c, void* stream, const void* f_src, void* f, void* ff, void* Phi, const void* u, const void* u0,
                                  double dt, int8_t* cbar, void* rhou, void* r, void* rhouf, const void* uOld, const void* drho,
                                  double lambda_rho, int limiter, int normal_scheme, const double uBC[3], unsigned perdir_mask,
                                  int exitBC, const int dirO[3], ifadv_report* report) {
  cudaEvent_t wait_f = nullptr;
  if (c) { wait_f = c->wait_f; c->wait_f = nullptr; }  // one-shot, consumed even when the call is rejected below
  int rc = check_common(c, normal_scheme, dirO);
  if (rc) return rc;
  if (limiter < 0 || limiter > 10) return fail(c, -2, "invalid limiter");
  if (!f_src || !f || !ff || !u || !u0 || !cbar || !rhou || !r || !uOld || !drho || !uBC || (c->D == 3 && (!Phi || !rhouf)))
    return fail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->dtype == IFADV_F32)
    return advect_vof_rhouu_t<float>(c, st, (float*)f, (float*)ff, (float*)Phi, (const float*)u, (const float*)u0, dt, cbar, (float*)rhou,
                                     (float*)r, (float*)rhouf, (const float*)uOld, (const float*)drho, lambda_rho, limiter, normal_scheme,
                                     uBC, perdir_mask, dirO, report, wait_f, exitBC, (const float*)f_src, 1);
  return advect_vof_rhouu_t<double>(c, st, (double*)f, (double*)ff, (double*)Phi, (const double*)u, (const double*)u0, dt, cbar,
                                    (double*)rhou, (double*)r, (double*)rhouf, (const double*)uOld, (const double*)drho, lambda_rho, limiter,
                                    normal_scheme, uBC, perdir_mask, dirO, report, wait_f, exitBC, (const double*)f_src, 1);
}

int ifadv_u2rhou(ifadv_ctx* c, void* stream, void* rhou, const void* u, const void* f, double lr) {
  if (!c || !rhou || !u || !f) return -2;
  if (c->dtype == IFADV_F32) return u2rhou_t<float>(c, (cudaStream_t)stream, (float*)rhou, (const float*)u, (const float*)f, lr, true);
  return u2rhou_t<double>(c, (cudaStream_t)stream, (double*)rhou, (const double*)u, (const double*)f, lr, true);
}
int ifadv_rhou2u(ifadv_ctx* c, void* stream, void* u, const void* rhou, const void* f, double lr) {
  if (!c || !rhou || !u || !f) return -2;
  if (c->dtype == IFADV_F32) return u2rhou_t<float>(c, (cudaStream_t)stream, (float*)u, (const float*)rhou, (const float*)f, lr, false);
  return u2rhou_t<double>(c, (cudaStream_t)stream, (double*)u, (const double*)rhou, (const double*)f, lr, false);
}
int ifadv_bc_vec(ifadv_ctx* c, void* stream, void* a, const double A[3], int saveexit, unsigned perdir_mask) {
  if (!c || !a || !A) return -2;
  if (c->dtype == IFADV_F32) return bcvec_t<float>(c, (cudaStream_t)stream, (float*)a, A, saveexit, perdir_mask);
  return bcvec_t<double>(c, (cudaStream_t)stream, (double*)a, A, saveexit, perdir_mask);
}
int ifadv_bcf(ifadv_ctx* c, void* stream, void* f, unsigned perdir_mask) {
  if (!c || !f) return -2;
  if (c->dtype == IFADV_F32) return launch_bcf<float>(c, (cudaStream_t)stream, (float*)f, perdir_mask);
  return launch_bcf<double>(c, (cudaStream_t)stream, (double*)f, perdir_mask);
}
int ifadv_axpby(ifadv_ctx* c, void* stream, void* out, double a, const void* x, double b, const void* y) {
  if (!c || !out || !x || !y) return -2;
  if (c->dtype == IFADV_F32) return axpby_t<float>(c, (cudaStream_t)stream, (float*)out, a, (const float*)x, b, (const float*)y);
  return axpby_t<double>(c, (cudaStream_t)stream, (double*)out, a, (const double*)x, b, (const double*)y);
}

int ifadv_mpcfl(ifadv_ctx* c, void* stream, const void* u, double nu, double mu, double lambda_mu, double lambda_rho, double eta,
                double gnorm, double dt_max, double safety, double* dt_out) {
  if (!c || !u || !dt_out) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  CU_CHECK(c, cudaMemsetAsync(c->misc_dev, 0, sizeof(unsigned long long) * 8, st));
  const int bx = 256;
  const long long rows_ = (long long)(c->g.n[1] - 2) * (c->D == 3 ? c->kz1 - c->kz0 : 1);
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((rows_ + 7) / 8, 148LL * 8));
  if (c->dtype == IFADV_F32) {
    if (c->D == 2) cfl_kernel<float, 2><<<grid, bx, 0, st>>>((const float*)u, c->g, c->misc_dev, c->kz0, c->kz1);
    else cfl_kernel<float, 3><<<grid, bx, 0, st>>>((const float*)u, c->g, c->misc_dev, c->kz0, c->kz1);
  } else {
    if (c->D == 2) cfl_kernel<double, 2><<<grid, bx, 0, st>>>((const double*)u, c->g, c->misc_dev, c->kz0, c->kz1);
    else cfl_kernel<double, 3><<<grid, bx, 0, st>>>((const double*)u, c->g, c->misc_dev, c->kz0, c->kz1);
  }
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  if (c->slab.nranks > 1) {  // order-preserving keys of non-negative maxima: the global maximum is the maximum of the keys
    if (!nccl_api()->ok) return fail(c, -4, "NCCL is not available");
    NCCL_CHECK(c, nccl_api()->AllReduce(c->misc_dev, c->misc_dev, 2, ncclUint64, ncclMax, (ncclComm_t)c->slab.comm, st));
  }
  CU_CHECK(c, cudaMemcpyAsync(c->misc_host, c->misc_dev, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, st));
  CU_CHECK(c, cudaStreamSynchronize(st));
  // ghosts of σ are 0 after fill!(a.σ,0) (flow.jl:264), so the maxima are at least 0 -- the keys start at 0.0's floor
  double mfo = c->misc_host[0] ? key_to_double(c->misc_host[0]) : 0.0, mtf = c->misc_host[1] ? key_to_double(c->misc_host[1]) : 0.0;
  mfo = std::fmax(mfo, 0.0); mtf = std::fmax(mtf, 0.0);
  auto compute = [&](auto tag) {
    using T = decltype(tag);
    const T dtAdv = T(1) / ((T)mfo + T(5) * (T)nu);
    const T dtVOF = T(1) / (T(2) * (T)mtf);
    const T dtGrav = (gnorm > 0) ? T(1) / (T(2) * (T)gnorm) : (T)dt_max;
    const T dtVisc = (mu > 0) ? T(3) / (T(14) * (T)mu * std::max(T(1), (T)lambda_mu / (T)lambda_rho)) : (T)dt_max;
    const T dtSurf = (eta > 0) ? std::sqrt((T(1) + (T)lambda_rho) / (T(8 * M_PI) * (T)eta)) : (T)dt_max;
    return (double)((T)safety * std::min(std::min(std::min(dtVOF, dtAdv), std::min(dtGrav, dtVisc)), dtSurf));
  };
  *dt_out = (c->dtype == IFADV_F32) ? compute(float()) : compute(double());
  return 0;
}

int ifadv_sum_inside(ifadv_ctx* c, void* stream, const void* f, double* out) {
  if (!c || !f || !out) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  CU_CHECK(c, cudaMemsetAsync(c->misc_dev, 0, sizeof(unsigned long long) * 8, st));
  const int bx = 256;
  const long long rows_ = (long long)(c->g.n[1] - 2) * (c->D == 3 ? c->kz1 - c->kz0 : 1);
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((rows_ + 7) / 8, 148LL * 8));
  double* acc = reinterpret_cast<double*>(c->misc_dev);
  if (c->dtype == IFADV_F32) {
    if (c->D == 2) sum_inside_kernel<float, 2><<<grid, bx, 0, st>>>((const float*)f, c->g, acc, c->kz0, c->kz1);
    else sum_inside_kernel<float, 3><<<grid, bx, 0, st>>>((const float*)f, c->g, acc, c->kz0, c->kz1);
  } else {
    if (c->D == 2) sum_inside_kernel<double, 2><<<grid, bx, 0, st>>>((const double*)f, c->g, acc, c->kz0, c->kz1);
    else sum_inside_kernel<double, 3><<<grid, bx, 0, st>>>((const double*)f, c->g, acc, c->kz0, c->kz1);
  }
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  if (c->slab.nranks > 1) {  // a slab sums its owned planes; the result is the global mass on every rank
    if (!nccl_api()->ok) return fail(c, -4, "NCCL is not available");
    NCCL_CHECK(c, nccl_api()->AllReduce(c->misc_dev, c->misc_dev, 1, ncclDouble, ncclSum, (ncclComm_t)c->slab.comm, st));
  }
  CU_CHECK(c, cudaMemcpyAsync(c->misc_host, c->misc_dev, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, st));
  CU_CHECK(c, cudaStreamSynchronize(st));
  memcpy(out, c->misc_host, sizeof(double));
  return 0;
}

int ifadv_apply_vof_samples(ifadv_ctx* c, void* stream, void* f, void* alpha, void* nhat, const void* sc, const void* sp, const void* sm) {
  if (!c || !f || !alpha || !nhat || !sc || !sp || !sm) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  const int bx = 128;
  dim3 grid = row_grid(c->g, c->D, bx);
  if (c->dtype == IFADV_F32) {
    const float tol = 10 * std::numeric_limits<float>::epsilon();
    if (c->D == 2) applyvof_kernel<float, 2><<<grid, bx, 0, st>>>((float*)f, (float*)alpha, (float*)nhat, (const float*)sc, (const float*)sp, (const float*)sm, c->g, tol);
    else applyvof_kernel<float, 3><<<grid, bx, 0, st>>>((float*)f, (float*)alpha, (float*)nhat, (const float*)sc, (const float*)sp, (const float*)sm, c->g, tol);
  } else {
    const double tol = 10 * std::numeric_limits<double>::epsilon();
    if (c->D == 2) applyvof_kernel<double, 2><<<grid, bx, 0, st>>>((double*)f, (double*)alpha, (double*)nhat, (const double*)sc, (const double*)sp, (const double*)sm, c->g, tol);
    else applyvof_kernel<double, 3><<<grid, bx, 0, st>>>((double*)f, (double*)alpha, (double*)nhat, (const double*)sc, (const double*)sp, (const double*)sm, c->g, tol);
  }
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

// One CMOM advection step on host buffers (what bench.py's e2e leg times).  Work arrays (device):
// w[0]=f w[1]=f⁰ w[2]=fᶠ w[3]=Φ w[4]=u w[5]=u⁰ w[6]=ρu w[7]=r w[8]=ρuf w[9]=dρ w[10]=c̄
}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// z-slab pipeline of the host-buffer entry point: H2D of slab i+1, the advection step of slab i and D2H of slab i-1
// run concurrently on three streams.  Every slab is the owned planes extended by HOST_W overlap planes per interior
// side and is advanced by a child context of that shape through the unchanged device entry points; the artificial slab
// ends contaminate at most 7 (lower) / 4 (upper) planes per advectfq!, which the overlap absorbs, so the
// owned planes are bit-identical to the single-pass result (tests/test_gpu_parity.py::test_host_entry_pipelined).
// ------------------------------------------------------------------------------------------------------------
namespace {
constexpr int HOST_W = 8;
struct HostChunk {
  int a, b;    // owned storage planes [a, b) (0-based; the first / last chunk also own the ghost plane of their physical end)
  int lo, hi;  // storage planes [lo, hi) held by the child (including its two ghost planes)
  ifadv_ctx* ctx;
  cudaEvent_t ev_in, ev_cmp, ev_out;
};
struct HostPipe {
  std::vector<HostChunk> ch;
  int nset = 0;
  void* w[3][11];
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
};

void host_pipe_free(ifadv_ctx* c) {
  HostPipe* hp = (HostPipe*)c->pipe;
  if (!hp) return;
  for (auto& h : hp->ch) {
    if (h.ctx) ifadv_destroy(h.ctx);
    cudaEventDestroy(h.ev_in); cudaEventDestroy(h.ev_cmp); cudaEventDestroy(h.ev_out);
  }
  for (int k = 0; k < hp->nset; ++k)
    for (auto& p : hp->w[k]) if (p) cudaFree(p);
  if (hp->s_in) cudaStreamDestroy(hp->s_in);
  if (hp->s_cmp) cudaStreamDestroy(hp->s_cmp);
  if (hp->s_out) cudaStreamDestroy(hp->s_out);
  delete hp;
  c->pipe = nullptr;
}

// planes per slab: IFADV_HOST_CHUNK (0 = single pass); default: an eighth of the interior planes for grids of >= 256 planes (512^3 f32:
// 57.1 ms per step with 8 slabs, 59.5 with 6, 62.3 with 4, 58.0 with 16 -- PCIe runs both directions at ~78 GB/s combined, so the
// short pipeline fill / drain of small slabs outweighs their extra overlap planes)
int host_chunk_planes(const ifadv_ctx* c, unsigned perdir_mask) {
  if (c->D != 3 || (perdir_mask & 4u)) return 0;  // a periodic z would need wrapped overlaps: single pass
  const int NI = c->g.n[2] - 2;
  const char* e = getenv("IFADV_HOST_CHUNK");
  int cp = e ? atoi(e) : (NI >= 256 ? (NI + 7) / 8 : 0);
  if (cp <= 0 || cp >= NI) return 0;
  return cp;
}

int host_pipe_build(ifadv_ctx* c, int cp) {
  HostPipe* hp = new HostPipe();
  c->pipe = hp;
  for (auto& set : hp->w) for (auto& p : set) p = nullptr;
  const int n2 = c->g.n[2];
  size_t maxpl = 0;
  for (int a = 1; a < n2 - 1; a += cp) {
    HostChunk h;
    const int b = std::min(a + cp, n2 - 1);
    h.lo = std::max(a - HOST_W, 1) - 1;       // child ghost plane below
    h.hi = std::min(b + HOST_W, n2 - 1) + 1;  // child ghost plane above (exclusive end)
    h.a = (a == 1) ? 0 : a;
    h.b = (b == n2 - 1) ? n2 : b;
    const int64_t ng[3] = {c->g.n[0], c->g.n[1], h.hi - h.lo};
    h.ctx = nullptr;
    CU_CHECK(c, cudaEventCreateWithFlags(&h.ev_in, cudaEventDisableTiming));
    CU_CHECK(c, cudaEventCreateWithFlags(&h.ev_cmp, cudaEventDisableTiming));
    CU_CHECK(c, cudaEventCreateWithFlags(&h.ev_out, cudaEventDisableTiming));
    hp->ch.push_back(h);
    if (ifadv_create(&hp->ch.back().ctx, 3, ng, c->dtype, c->device) != 0) { c->err = "child context creation failed"; return -3; }
    hp->ch.back().ctx->use_march = c->use_march; hp->ch.back().ctx->use_along2 = c->use_along2; hp->ch.back().ctx->use_xrow = c->use_xrow; hp->ch.back().ctx->use_vofcell = c->use_vofcell;
    maxpl = std::max(maxpl, (size_t)(h.hi - h.lo));
  }
  hp->nset = (int)std::min<size_t>(3, hp->ch.size());
  const size_t es = (c->dtype == IFADV_F32) ? 4 : 8;
  const size_t S = (size_t)c->g.s2 * maxpl, sb = es * S, vb = 3 * sb;
  const size_t sz[11] = {sb, sb, sb, sb, vb, vb, vb, vb, vb, vb, S};
  CU_CHECK(c, cudaStreamCreateWithFlags(&hp->s_in, cudaStreamNonBlocking));
  CU_CHECK(c, cudaStreamCreateWithFlags(&hp->s_cmp, cudaStreamNonBlocking));
  CU_CHECK(c, cudaStreamCreateWithFlags(&hp->s_out, cudaStreamNonBlocking));
  for (int k = 0; k < hp->nset; ++k) {
    for (int i = 0; i < 11; ++i) {
      CU_CHECK(c, cudaMalloc(&hp->w[k][i], sz[i]));
      CU_CHECK(c, cudaMemsetAsync(hp->w[k][i], 0, sz[i], hp->s_cmp));  // ρu ghosts (never written by the sweeps) read back as zeros
    }
    // dρ keeps its constructor value 1 (cVOF.jl:74)
    const long long n = (long long)S * 3;
    if (c->dtype == IFADV_F32) fill_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, hp->s_cmp>>>((float*)hp->w[k][9], 1.f, n);
    else fill_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, hp->s_cmp>>>((double*)hp->w[k][9], 1.0, n);
    c->launches++;
  }
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

int host_step_pipelined(ifadv_ctx* c, char* fh, const char* uh, char* rh, double dt, double lambda_rho, int limiter, int normal_scheme,
                        const double uBC[3], unsigned perdir_mask, const int dirO[3], ifadv_report* report) {
  HostPipe* hp = (HostPipe*)c->pipe;
  const size_t es = (c->dtype == IFADV_F32) ? 4 : 8;
  const size_t pl = es * (size_t)c->g.s2;  // bytes per plane
  const size_t Sp = es * (size_t)c->g.S;   // bytes per parent component
  const int nch = (int)hp->ch.size();
  int64_t launches0 = 0;
  for (auto& h : hp->ch) launches0 += h.ctx->launches;
  c->host_h2d = c->host_d2h = 0; c->host_slabs = nch;
  for (int i = 0; i < nch; ++i) {
    HostChunk& h = hp->ch[i];
    void** w = hp->w[i % hp->nset];
    void *f = w[0], *f0 = w[1], *ff = w[2], *Phi = w[3], *u = w[4], *u0 = w[5], *ru = w[6], *r = w[7], *ruf = w[8], *drho = w[9];
    int8_t* cbar = (int8_t*)w[10];
    const size_t npl = (size_t)(h.hi - h.lo), Sc = pl * npl;  // child planes, bytes per child component
    // H2D: the buffer set is free once the slab that used it before has been copied out
    if (i >= hp->nset) CU_CHECK(c, cudaStreamWaitEvent(hp->s_in, hp->ch[i - hp->nset].ev_out, 0));
    CU_CHECK(c, cudaMemcpyAsync(f, fh + pl * h.lo, Sc, cudaMemcpyHostToDevice, hp->s_in));
    // u is read-only in the step, so the planes this slab shares with the previous one (its overlap + the neighbour's overlap + two
    // child ghost planes) are already on the device: copy them from the previous slab's buffer instead of over PCIe again
    size_t nshare = 0;
    if (i > 0 && hp->nset >= 2 && !getenv("IFADV_HOST_NOSHARE")) {
      const HostChunk& p = hp->ch[i - 1];
      const size_t Scp = pl * (size_t)(p.hi - p.lo);
      const void* up = hp->w[(i - 1) % hp->nset][4];
      nshare = (size_t)std::min(std::max(p.hi - h.lo, 0), h.hi - h.lo);
      for (int d = 0; d < 3 && nshare; ++d)
        CU_CHECK(c, cudaMemcpyAsync((char*)u + Sc * d, (const char*)up + Scp * d + pl * (size_t)(h.lo - p.lo), pl * nshare,
                                    cudaMemcpyDeviceToDevice, hp->s_in));
    }
    for (int d = 0; d < 3 && npl > nshare; ++d)
      CU_CHECK(c, cudaMemcpyAsync((char*)u + Sc * d + pl * nshare, uh + Sp * d + pl * (h.lo + nshare), pl * (npl - nshare),
                                  cudaMemcpyHostToDevice, hp->s_in));
    CU_CHECK(c, cudaEventRecord(h.ev_in, hp->s_in));
    c->host_h2d += (int64_t)(Sc + 3 * pl * (npl - nshare));
    // the step of this slab (same sequence as the single-pass entry)
    CU_CHECK(c, cudaStreamWaitEvent(hp->s_cmp, h.ev_in, 0));
    CU_CHECK(c, cudaMemcpyAsync(u0, u, 3 * Sc, cudaMemcpyDeviceToDevice, hp->s_cmp));
    int rc;
    if ((rc = ifadv_u2rhou_advect_vof_rhouu(h.ctx, hp->s_cmp, f, f0, ff, Phi, u, u, dt, cbar, ru, r, ruf, u, drho, lambda_rho, limiter,
                                            normal_scheme, uBC, perdir_mask, 0, dirO, nullptr)) < 0) { c->err = h.ctx->err; return rc; }
    if ((rc = ifadv_axpby(h.ctx, hp->s_cmp, f0, 0.5, f0, 0.5, f))) { c->err = h.ctx->err; return rc; }
    CU_CHECK(c, cudaMemcpyAsync(f0, f, Sc, cudaMemcpyDeviceToDevice, hp->s_cmp));
    if ((rc = ifadv_u2rhou_advect_vof_rhouu(h.ctx, hp->s_cmp, f, f, ff, Phi, u, u, dt, cbar, ru, r, ruf, u0, drho, lambda_rho, limiter,
                                            normal_scheme, uBC, perdir_mask, 0, dirO, nullptr)) < 0) { c->err = h.ctx->err; return rc; }
    if (report) CU_CHECK(c, cudaMemcpyAsync(h.ctx->red_host, h.ctx->red_dev, sizeof(unsigned long long) * 24, cudaMemcpyDeviceToHost, hp->s_cmp));
    CU_CHECK(c, cudaEventRecord(h.ev_cmp, hp->s_cmp));
    // D2H of the owned planes
    CU_CHECK(c, cudaStreamWaitEvent(hp->s_out, h.ev_cmp, 0));
    const size_t off_c = pl * (size_t)(h.a - h.lo), nown = pl * (size_t)(h.b - h.a);
    CU_CHECK(c, cudaMemcpyAsync(fh + pl * h.a, (char*)f + off_c, nown, cudaMemcpyDeviceToHost, hp->s_out));
    for (int d = 0; d < 3; ++d)
      CU_CHECK(c, cudaMemcpyAsync(rh + Sp * d + pl * h.a, (char*)ru + Sc * d + off_c, nown, cudaMemcpyDeviceToHost, hp->s_out));
    CU_CHECK(c, cudaEventRecord(h.ev_out, hp->s_out));
    c->host_d2h += (int64_t)(4 * nown);
  }
  CU_CHECK(c, cudaStreamSynchronize(hp->s_out));
  CU_CHECK(c, cudaStreamSynchronize(hp->s_cmp));
  int status = 0;
  for (auto& h : hp->ch) c->launches += h.ctx->launches;
  c->launches -= launches0;
  if (report) {
    // worst slab wins; extrema located in the overlap planes of a slab belong to the neighbour (or to the contaminated band) and are skipped
    const double filltol = 100.0 * ((c->dtype == IFADV_F32) ? (double)std::numeric_limits<float>::epsilon() : std::numeric_limits<double>::epsilon());
    report->status = 0; report->dir = -1; report->maxf = 0; report->minf = 0;
    for (int k = 0; k < 3; ++k) report->argmax[k] = report->argmin[k] = 0;
    bool have = false;
    for (auto& h : hp->ch) {
      ifadv_report r;
      const int st = decode_report(h.ctx, dirO, filltol, &r);
      if (st < 0) { *report = r; report->argmax[2] += h.lo; report->argmin[2] += h.lo; return st; }
      auto owned = [&](int64_t z) { const int64_t p = z - 1 + h.lo; return p >= h.a && p < h.b; };
      int s2 = 0;
      if ((st & 1) && owned(r.argmax[2])) s2 |= 1;
      if ((st & 2) && owned(r.argmin[2])) s2 |= 2;
      if (!have || s2 != 0) {
        const int keep = report->status;
        *report = r; report->argmax[2] += h.lo; report->argmin[2] += h.lo;
        report->status = keep | s2;
        have = true;
      }
      status |= s2;
    }
    report->status = status;
  }
  return status;
}
}  // namespace

extern "C" {

int ifadv_mom_advect_step_host(ifadv_ctx* c, void* f_host, const void* u_host, void* rhou_host, double dt, double lambda_rho, int limiter,
                               int normal_scheme, const double uBC[3], unsigned perdir_mask, const int dirO[3], ifadv_report* report) {
  if (!c || !f_host || !u_host || !rhou_host || !uBC || !dirO) return -2;
  const size_t es = esize(c->dtype), sb = es * c->g.S, vb = sb * c->D;
  CU_CHECK(c, cudaSetDevice(c->device));
  // large 3-D grids: z-slab pipeline (copies overlap the kernels); buffers must be page-locked for the copies to be asynchronous
  if (const int cp = host_chunk_planes(c, perdir_mask)) {
    auto pinned = [](const void* p) {
      cudaPointerAttributes a;
      if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
      return a.type == cudaMemoryTypeHost;
    };
    if (pinned(f_host) && pinned(u_host) && pinned(rhou_host)) {
      if (!c->pipe) { const int rb = host_pipe_build(c, cp); if (rb != 0) { host_pipe_free(c); return rb; } }
      return host_step_pipelined(c, (char*)f_host, (const char*)u_host, (char*)rhou_host, dt, lambda_rho, limiter, normal_scheme, uBC,
                                 perdir_mask, dirO, report);
    }
  }
  if (!c->own_stream) CU_CHECK(c, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  cudaStream_t st = c->own_stream;
  if (!c->w[0]) {
    const size_t sz[11] = {sb, sb, sb, sb, vb, vb, vb, vb, vb, vb, (size_t)c->g.S};
    for (int k = 0; k < 11; ++k) {
      CU_CHECK(c, cudaMalloc(&c->w[k], sz[k]));
      CU_CHECK(c, cudaMemsetAsync(c->w[k], 0, sz[k], st));  // ρu ghosts (never written by the sweeps) go back to the host as zeros
    }
    // dρ keeps its constructor value 1 (cVOF.jl:74)
    {
      const long long n = (long long)c->g.S * c->D;
      const int bs = 256;
      if (c->dtype == IFADV_F32) fill_kernel<float><<<(unsigned)((n + bs - 1) / bs), bs, 0, st>>>((float*)c->w[9], 1.f, n);
      else fill_kernel<double><<<(unsigned)((n + bs - 1) / bs), bs, 0, st>>>((double*)c->w[9], 1.0, n);
      c->launches++;
      CU_CHECK(c, cudaGetLastError());
    }
    CU_CHECK(c, cudaMallocHost(&c->pin_f, sb));
    CU_CHECK(c, cudaMallocHost(&c->pin_u, vb));
    CU_CHECK(c, cudaMallocHost(&c->pin_ru, vb));
  }
  void *f = c->w[0], *f0 = c->w[1], *ff = c->w[2], *Phi = c->w[3], *u = c->w[4], *u0 = c->w[5], *ru = c->w[6], *r = c->w[7], *ruf = c->w[8],
       *drho = c->w[9];
  int8_t* cbar = (int8_t*)c->w[10];
  // caller buffers that are already page-locked (cudaHostRegister / pinned allocators) are used directly
  auto pinned = [](const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  const bool pf = pinned(f_host), pu = pinned(u_host), pr = pinned(rhou_host);
  if (!pf) memcpy(c->pin_f, f_host, sb);
  if (!pu) memcpy(c->pin_u, u_host, vb);
  CU_CHECK(c, cudaMemcpyAsync(f, pf ? f_host : c->pin_f, sb, cudaMemcpyHostToDevice, st));
  CU_CHECK(c, cudaMemcpyAsync(u, pu ? u_host : c->pin_u, vb, cudaMemcpyHostToDevice, st));
  // copyto!(u⁰,u)                                                          flow.jl:61
  CU_CHECK(c, cudaMemcpyAsync(u0, u, vb, cudaMemcpyDeviceToDevice, st));
  int rc;
  // predictor: copyto!(f⁰,f); u2ρu!(ρu,u⁰,f⁰); BC!; advectfq!(f⁰; u⁰,u,uOld=u)     flow.jl:61,69-70 (fused entry)
  // (u is passed for both velocity arguments: u⁰≡u here, and one array selects the kernels without the second velocity stream)
  if ((rc = ifadv_u2rhou_advect_vof_rhouu(c, st, f, f0, ff, Phi, u, u, dt, cbar, ru, r, ruf, u, drho, lambda_rho, limiter, normal_scheme,
                                          uBC, perdir_mask, 0, dirO, nullptr)) < 0) return rc;
  if ((rc = ifadv_axpby(c, st, f0, 0.5, f0, 0.5, f))) return rc;  // flow.jl:74
  // corrector: copyto!(f⁰,f); u2ρu!(ρu,u⁰,f); BC!; advectfq!(f; u,u,uOld=u⁰)         flow.jl:89-92
  CU_CHECK(c, cudaMemcpyAsync(f0, f, sb, cudaMemcpyDeviceToDevice, st));
  rc = ifadv_u2rhou_advect_vof_rhouu(c, st, f, f, ff, Phi, u, u, dt, cbar, ru, r, ruf, u0, drho, lambda_rho, limiter, normal_scheme, uBC,
                                     perdir_mask, 0, dirO, report);
  if (rc < 0) return rc;
  CU_CHECK(c, cudaMemcpyAsync(pf ? f_host : c->pin_f, f, sb, cudaMemcpyDeviceToHost, st));
  CU_CHECK(c, cudaMemcpyAsync(pr ? rhou_host : c->pin_ru, ru, vb, cudaMemcpyDeviceToHost, st));
  CU_CHECK(c, cudaStreamSynchronize(st));
  c->host_h2d = (int64_t)(sb + vb); c->host_d2h = (int64_t)(sb + vb); c->host_slabs = 1;
  if (!pf) memcpy(f_host, c->pin_f, sb);
  if (!pr) memcpy(rhou_host, c->pin_ru, vb);
  return rc;
}

int ifadv_visc_surften_rhou(ifadv_ctx* c, void* stream, void* r, const void* u, void* Phi, const void* f, void* alpha, const void* nhat,
                            const void* fbuffer, double lambda_mu, double mu, double lambda_rho, double eta, unsigned perdir_mask) {
  (void)Phi; (void)alpha;
  if (!c) return -2;
  if (c->slab.nranks > 1) return fail(c, -2, "the forcing entry points are single-GPU (the height-function column walks are unbounded along z)");
  if (!r || !u || !f || (mu > 0.0 && !nhat) || (eta > 0.0 && !fbuffer)) return fail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->dtype == IFADV_F32)
    return visc_surften_t<float>(c, st, (float*)r, (const float*)u, (const float*)f, (const float*)nhat, (const float*)fbuffer, lambda_mu, mu,
                                 lambda_rho, eta, perdir_mask);
  return visc_surften_t<double>(c, st, (double*)r, (const double*)u, (const double*)f, (const double*)nhat, (const double*)fbuffer, lambda_mu, mu,
                                lambda_rho, eta, perdir_mask);
}

int ifadv_update_u(ifadv_ctx* c, void* stream, void* u, void* rhou, const void* rhou0, void* forcing, double dt, const void* f,
                   double lambda_rho, const double g[3], double w) {
  if (!c) return -2;
  if (!u || !rhou || !rhou0 || !forcing || !f) return fail(c, -2, "null array");
  if (!(w > 0.0)) return fail(c, -2, "invalid weight w");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->dtype == IFADV_F32)
    return update_u_t<float>(c, st, (float*)u, (float*)rhou, (const float*)rhou0, (float*)forcing, dt, (const float*)f, lambda_rho, g, w);
  return update_u_t<double>(c, st, (double*)u, (double*)rhou, (const double*)rhou0, (double*)forcing, dt, (const double*)f, lambda_rho, g, w);
}

int ifadv_update_l(ifadv_ctx* c, void* stream, void* mu0, const void* f, double lambda_rho, unsigned perdir_mask, int fill_one) {
  if (!c) return -2;
  if (!mu0 || !f) return fail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->dtype == IFADV_F32) return update_l_t<float>(c, st, (float*)mu0, (const float*)f, lambda_rho, perdir_mask, fill_one);
  return update_l_t<double>(c, st, (double*)mu0, (const double*)f, lambda_rho, perdir_mask, fill_one);
}

int ifadv_defer_f_writes_until(ifadv_ctx* c, void* event) {
  if (!c) return -2;
  c->wait_f = (cudaEvent_t)event;
  return 0;
}

// ---- z-slab decomposition ------------------------------------------------------------------------------------------------
int ifadv_nccl_unique_id(char id[128]) {
  if (!id) return -2;
  NcclApi* A = nccl_api();
  if (!A->ok) return -4;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  if (A->GetUniqueId(&u) != ncclSuccess) return -4;
  memcpy(id, &u, 128);
  return 0;
}
int ifadv_nccl_comm_init(void** comm, int nranks, const char id[128], int rank, int device) {
  if (!comm || !id || nranks < 1 || rank < 0 || rank >= nranks) return -2;
  NcclApi* A = nccl_api();
  if (!A->ok) return -4;
  if (cudaSetDevice(device) != cudaSuccess) return -3;
  ncclUniqueId u;
  memcpy(&u, id, 128);
  ncclComm_t c;
  if (A->CommInitRank(&c, nranks, u, rank) != ncclSuccess) return -4;
  *comm = (void*)c;
  return 0;
}
int ifadv_nccl_comm_destroy(void* comm) {
  if (!comm) return 0;
  NcclApi* A = nccl_api();
  if (!A->ok) return -4;
  return A->CommDestroy((ncclComm_t)comm) == ncclSuccess ? 0 : -4;
}

int ifadv_create_slab(ifadv_ctx** out, const int64_t Ng_local[3], int dtype, int device, void* nccl_comm, int rank, int nranks,
                      int ghost_planes, int periodic_z) {
  if (!out || !Ng_local || nranks < 1 || rank < 0 || rank >= nranks || ghost_planes < 3) return -2;
  const int G = ghost_planes;
  const bool lo = nranks > 1 && (rank > 0 || periodic_z), hi = nranks > 1 && (rank < nranks - 1 || periodic_z);
  const int glo = lo ? G : 0, ghi = hi ? G : 0;
  if (Ng_local[2] - 2 - glo - ghi < G) return -2;  // a slab sends G owned planes to each neighbour
  if (nranks > 1 && !nccl_comm) return -2;
  int rc = ifadv_create(out, 3, Ng_local, dtype, device);
  if (rc) return rc;
  ifadv_ctx* c = *out;
  c->slab.comm = nccl_comm; c->slab.rank = rank; c->slab.nranks = nranks; c->slab.G = G; c->slab.glo = glo; c->slab.ghi = ghi;
  c->slab.lower = lo ? (rank + nranks - 1) % nranks : -1;
  c->slab.upper = hi ? (rank + 1) % nranks : -1;
  c->slab.bytes_sent = 0;
  c->kz0 = 2 + glo;
  c->kz1 = (int)Ng_local[2] - ghi;
  {
    // IFADV_SLAB_OVERLAP=1: boundary planes first, exchange on a second (high-priority) stream underneath the interior planes.
    // Measured on 4 / 8 B200 (profiles/r02_multigpu.md): with NCCL send/recv the split does not pay -- the three launches per sweep
    // and the exchange kernel competing for SMs cost more than the exchange they hide -- so the default is the in-line exchange.
    const char* e = getenv("IFADV_SLAB_OVERLAP");
    c->slab.overlap = (e && atoi(e) != 0) ? 1 : 0;
  }
  if (nranks > 1) {
    int plo = 0, phi = 0;  // the exchange must not queue behind the CTAs of the interior sweep: highest priority
    CU_CHECK(c, cudaDeviceGetStreamPriorityRange(&plo, &phi));
    CU_CHECK(c, cudaStreamCreateWithPriority(&c->slab_stream, cudaStreamNonBlocking, phi));
    CU_CHECK(c, cudaEventCreateWithFlags(&c->slab_ev[0], cudaEventDisableTiming));
    CU_CHECK(c, cudaEventCreateWithFlags(&c->slab_ev[1], cudaEventDisableTiming));
    if ((rc = slab_p2p_setup(c))) return rc;
  }
  return 0;
}
int ifadv_slab_p2p(const ifadv_ctx* c) { return c ? c->p2p.on : 0; }
int ifadv_slab_info(const ifadv_ctx* c, int* kz0, int* kz1, int* lower, int* upper, int64_t* bytes_sent) {
  if (!c) return -2;
  if (kz0) *kz0 = c->kz0;
  if (kz1) *kz1 = c->kz1;
  if (lower) *lower = c->slab.lower;
  if (upper) *upper = c->slab.upper;
  if (bytes_sent) *bytes_sent = c->slab.bytes_sent;
  return 0;
}
int ifadv_exchange_planes(ifadv_ctx* c, void* stream, void* field, int ncomp, int elem_bytes) {
  if (!c || !field || ncomp < 1 || elem_bytes < 1) return -2;
  return slab_exchange(c, (cudaStream_t)stream, field, (size_t)elem_bytes, ncomp);
}

int ifadv_check_nan(ifadv_ctx* c, void* stream) {
  if (!c) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  CU_CHECK(c, cudaMemcpyAsync(c->red_host, c->red_dev, sizeof(unsigned long long) * 32, cudaMemcpyDeviceToHost, st));
  CU_CHECK(c, cudaStreamSynchronize(st));
  if (c->p2p.on) {  // a peer-to-peer exchange that gave up waiting for a neighbour (ranks out of step) is a communication error
    unsigned t = 0;
    CU_CHECK(c, cudaMemcpyAsync(&t, c->p2p.flags + 8, sizeof t, cudaMemcpyDeviceToHost, st));
    CU_CHECK(c, cudaStreamSynchronize(st));
    if (t) return fail(c, -4, "slab exchange timed out waiting for a neighbour");
  }
  bool nan = c->red_host[24] != 0ull;
  for (int s = 0; s < 3; ++s) nan = nan || c->red_host[8 * s + 4] != 0ull;  // the most recent call's own sweeps
  if (!nan) return 0;
  // consumed: clear the sticky flag and the counts it was derived from
  CU_CHECK(c, cudaMemsetAsync(c->red_dev + 24, 0, sizeof(unsigned long long), st));
  for (int s = 0; s < 3; ++s) CU_CHECK(c, cudaMemsetAsync(c->red_dev + 8 * s + 4, 0, sizeof(unsigned long long), st));
  return fail(c, -1, "NaN in f");
}

int ifadv_host_step_bytes(const ifadv_ctx* c, int64_t* h2d, int64_t* d2h, int* slabs) {
  if (!c) return -2;
  if (h2d) *h2d = c->host_h2d;
  if (d2h) *d2h = c->host_d2h;
  if (slabs) *slabs = c->host_slabs;
  return 0;
}

}  // extern "C"

// ---- post-processing (ifadv_post.cuh) ------------------------------------------------------------------------------------------------
template <class T> static int redist_l_t(ifadv_ctx* c, cudaStream_t st, T* L, const T* phi, const T* pini, unsigned per) {
  Geo g = c->g;
  g.per = per;
  const int bx = 128;
  const dim3 gi = row_grid(g, c->D, bx);
  if (c->D == 2) redist_l_kernel<T, 2><<<gi, bx, 0, st>>>(L, phi, pini, g);
  else redist_l_kernel<T, 3><<<gi, bx, 0, st>>>(L, phi, pini, g);
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}
template <class T>
static int redist_stage_t(ifadv_ctx* c, cudaStream_t st, T* phi, const T* phi0, const T* pini, T* L, double dtau, double alpha, unsigned per) {
  int rc = redist_l_t<T>(c, st, L, phi, pini, per);
  if (rc) return rc;
  const int bx = 128;
  const dim3 gi = row_grid(c->g, c->D, bx);
  if (c->D == 2) redist_stage_kernel<T, 2><<<gi, bx, 0, st>>>(phi, phi0, L, c->g, (T)dtau, (T)alpha);
  else redist_stage_kernel<T, 3><<<gi, bx, 0, st>>>(phi, phi0, L, c->g, (T)dtau, (T)alpha);
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}
template <class T>
static int redistance_t(ifadv_ctx* c, cudaStream_t st, T* phi, T* phi0, const T* pini, T* L, double d, double dtau, unsigned per) {
  const int itmx = (int)std::nearbyint(d / dtau);  // round(T, d/dτ), redistaning.jl:46
  const double al[3] = {0.0, 3.0 / 4, 1.0 / 3};    // third-order SSP Runge-Kutta (Shu-Osher), :49-54
  for (int it = 0; it < itmx; ++it) {
    CU_CHECK(c, cudaMemcpyAsync(phi0, phi, sizeof(T) * (size_t)c->g.S, cudaMemcpyDeviceToDevice, st));
    for (int s = 0; s < 3; ++s) {
      int rc = redist_stage_t<T>(c, st, phi, phi0, pini, L, dtau, (double)(T)al[s], per);
      if (rc) return rc;
      if ((rc = launch_bcf<T>(c, st, phi, per))) return rc;
    }
  }
  return 0;
}

extern "C" {
int ifadv_levelset_init(ifadv_ctx* c, void* stream, void* phi, void* phi_ini, const void* f) {
  if (!c) return -2;
  if (!phi || !phi_ini || !f) return fail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = c->g.S;
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (c->dtype == IFADV_F32) levelset_kernel<float><<<nb, 256, 0, st>>>((float*)phi, (float*)phi_ini, (const float*)f, n);
  else levelset_kernel<double><<<nb, 256, 0, st>>>((double*)phi, (double*)phi_ini, (const double*)f, n);
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  return 0;
}
int ifadv_redist_compute_l(ifadv_ctx* c, void* stream, void* L, const void* phi, const void* phi_ini, unsigned perdir_mask) {
  if (!c) return -2;
  if (c->slab.nranks > 1) return fail(c, -2, "the post-processing entry points are single-GPU");
  if (!L || !phi || !phi_ini) return fail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->dtype == IFADV_F32) return redist_l_t<float>(c, st, (float*)L, (const float*)phi, (const float*)phi_ini, perdir_mask);
  return redist_l_t<double>(c, st, (double*)L, (const double*)phi, (const double*)phi_ini, perdir_mask);
}
int ifadv_redist_stage(ifadv_ctx* c, void* stream, void* phi, const void* phi0, const void* phi_ini, void* L, double dtau, double alpha,
                       unsigned perdir_mask) {
  if (!c) return -2;
  if (c->slab.nranks > 1) return fail(c, -2, "the post-processing entry points are single-GPU");
  if (!L || !phi || !phi0 || !phi_ini) return fail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->dtype == IFADV_F32)
    return redist_stage_t<float>(c, st, (float*)phi, (const float*)phi0, (const float*)phi_ini, (float*)L, dtau, alpha, perdir_mask);
  return redist_stage_t<double>(c, st, (double*)phi, (const double*)phi0, (const double*)phi_ini, (double*)L, dtau, alpha, perdir_mask);
}
int ifadv_redistance(ifadv_ctx* c, void* stream, void* phi, void* phi0, const void* phi_ini, void* L, double d, double dtau,
                     unsigned perdir_mask) {
  if (!c) return -2;
  if (c->slab.nranks > 1) return fail(c, -2, "the post-processing entry points are single-GPU");
  if (!L || !phi || !phi0 || !phi_ini) return fail(c, -2, "null array");
  if (!(dtau > 0.0) || !(d >= 0.0)) return fail(c, -2, "invalid pseudo-time step");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->dtype == IFADV_F32) return redistance_t<float>(c, st, (float*)phi, (float*)phi0, (const float*)phi_ini, (float*)L, d, dtau, perdir_mask);
  return redistance_t<double>(c, st, (double*)phi, (double*)phi0, (const double*)phi_ini, (double*)L, d, dtau, perdir_mask);
}
int ifadv_metrics(ifadv_ctx* c, void* stream, const void* u, const void* f, double lambda_rho, const double U[3], const double g[3],
                  const double statWL[3], double out[5]) {
  if (!c || !out) return -2;
  if (c->slab.nranks > 1) return fail(c, -2, "the post-processing entry points are single-GPU");
  if (!u || !f) return fail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  CU_CHECK(c, cudaMemsetAsync(c->misc_dev, 0, sizeof(unsigned long long) * 8, st));
  const long long rows_ = (long long)(c->g.n[1] - 2) * (c->D == 3 ? c->g.n[2] - 2 : 1);
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((rows_ + 7) / 8, 148LL * 8));
  double* acc = reinterpret_cast<double*>(c->misc_dev);
  double UU[3] = {0, 0, 0}, G[3] = {0, 0, 0}, W[3] = {0, 0, 0};
  for (int i = 0; i < c->D; ++i) { if (U) UU[i] = U[i]; if (g) G[i] = g[i]; if (statWL) W[i] = statWL[i]; }
  if (c->dtype == IFADV_F32) {
    if (c->D == 2) metrics_kernel<float, 2><<<grid, 256, 0, st>>>((const float*)u, (const float*)f, c->g, (float)lambda_rho, (float)UU[0], (float)UU[1], (float)UU[2], (float)G[0], (float)G[1], (float)G[2], (float)W[0], (float)W[1], (float)W[2], acc);
    else metrics_kernel<float, 3><<<grid, 256, 0, st>>>((const float*)u, (const float*)f, c->g, (float)lambda_rho, (float)UU[0], (float)UU[1], (float)UU[2], (float)G[0], (float)G[1], (float)G[2], (float)W[0], (float)W[1], (float)W[2], acc);
  } else {
    if (c->D == 2) metrics_kernel<double, 2><<<grid, 256, 0, st>>>((const double*)u, (const double*)f, c->g, lambda_rho, UU[0], UU[1], UU[2], G[0], G[1], G[2], W[0], W[1], W[2], acc);
    else metrics_kernel<double, 3><<<grid, 256, 0, st>>>((const double*)u, (const double*)f, c->g, lambda_rho, UU[0], UU[1], UU[2], G[0], G[1], G[2], W[0], W[1], W[2], acc);
  }
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  CU_CHECK(c, cudaMemcpyAsync(c->misc_host, c->misc_dev, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, st));
  CU_CHECK(c, cudaStreamSynchronize(st));
  memcpy(out, c->misc_host, sizeof(double) * 5);
  return 0;
}
int ifadv_enstrophy(ifadv_ctx* c, void* stream, const void* omega, double* out) {
  if (!c || !out) return -2;
  if (c->slab.nranks > 1) return fail(c, -2, "the post-processing entry points are single-GPU");
  if (!omega) return fail(c, -2, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  CU_CHECK(c, cudaMemsetAsync(c->misc_dev, 0, sizeof(unsigned long long) * 8, st));
  const long long rows_ = (long long)(c->g.n[1] - 2) * (c->D == 3 ? c->g.n[2] - 2 : 1);
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((rows_ + 7) / 8, 148LL * 8));
  double* acc = reinterpret_cast<double*>(c->misc_dev);
  if (c->dtype == IFADV_F32) {
    if (c->D == 2) enstrophy_kernel<float, 2><<<grid, 256, 0, st>>>((const float*)omega, c->g, acc);
    else enstrophy_kernel<float, 3><<<grid, 256, 0, st>>>((const float*)omega, c->g, acc);
  } else {
    if (c->D == 2) enstrophy_kernel<double, 2><<<grid, 256, 0, st>>>((const double*)omega, c->g, acc);
    else enstrophy_kernel<double, 3><<<grid, 256, 0, st>>>((const double*)omega, c->g, acc);
  }
  c->launches++;
  CU_CHECK(c, cudaGetLastError());
  CU_CHECK(c, cudaMemcpyAsync(c->misc_host, c->misc_dev, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, st));
  CU_CHECK(c, cudaStreamSynchronize(st));
  memcpy(out, c->misc_host, sizeof(double));
  return 0;
}
}  // extern "C" (post-processing)
