// ifadv_post.cuh -- post-processing: level-set redistancing (src/redistaning.jl:8-143) and the energy /
// momentum / enstrophy metrics (src/metrics.jl:15-50).
//   redist_l_kernel      computeL!: per cell the D one-sided ENO contributions 𝛁ϕᵢ² (the Neumann / periodic boundary blocks of the
//                        reference folded into the index of the two outer stencil points), then L = ϕini·(1-√Σ);
//   redist_stage_kernel  ϕ ← αϕ⁰ + (1-α)(ϕ + dτ·L) on inside(ϕ)  (+ bcf_kernel for the ghost cells, ifadv_b200.cu);
//   levelset_kernel      ϕ = 2f-1, ϕini = ϕ;
//   metrics_kernel       Σ ρkeI, Σ ρgh, Σ ρuI(i) over inside(f) in one pass (rows walked by warps, Float64 partials, warp shuffle + one
//                        atomic per CTA and quantity -- the sum_inside_kernel pattern);  enstrophy_kernel: Σ EnsI.
// Compiled -fmad=false with IEEE division / square root in both precisions: the field results equal the oracle's bit for bit.
#pragma once
#include "ifadv_math.cuh"
#include "ifadv_sweep.cuh"

namespace ifadv {

template <class T> IFADV_DI T minmod2(T a, T b) { return (t_abs(a) <= t_abs(b)) ? a : b; }  // redistaning.jl:106
// 𝛁ϕᵢ²(a,b,c,d,e,s), redistaning.jl:119-143
template <class T> IFADV_DI T gradphi2(T a, T b, T c, T d, T e, T s) {
  const T dp = d - c, dm = c - b;
  const T ddp = e + c - T(2) * d, dd0 = d + b - T(2) * c, ddm = c + a - T(2) * b;
  const T dR = dp - minmod2(ddp, dd0) / T(2);
  const T dL = dm + minmod2(dd0, ddm) / T(2);
  const T wR = dR * s, wL = dL * s;
  if (wR < T(0) && (wR + wL) < T(0)) return dR * dR;
  if (wL > T(0) && (wR + wL) > T(0)) return dL * dL;
  return T(0);
}

// computeL!(L,ϕ,ϕini;perdir) on inside(ϕ), redistaning.jl:67-104
template <class T, int D> __global__ void __launch_bounds__(128) redist_l_kernel(T* __restrict__ L, const T* __restrict__ phi, const T* __restrict__ pini,
                                                                                 const Geo g) {
  const int x = 2 + blockIdx.x * blockDim.x + threadIdx.x, y = 2 + blockIdx.y, z = (D == 3) ? 2 + blockIdx.z : 1;
  if (x > g.n[0] - 1) return;
  const long long l = lin3(g, x, y, z);
  const int v[3] = {x, y, z};
  const T c = __ldg(phi + l), pi = __ldg(pini + l);
  const T s = t_sign(pi);
  T acc = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const long long st = (i == 0) ? 1 : ((i == 1) ? g.s1 : g.s2);
    const int n = g.n[i];
    const bool per = (g.per >> i) & 1u;
    const T b = __ldg(phi + l - st), d = __ldg(phi + l + st);
    T a, e;
    if (v[i] == 2) a = per ? __ldg(phi + l + (long long)(n - 2 - v[i]) * st) : b;   // lowerL!: ϕ[CIj(i,I,N-2)] / mirrored
    else a = __ldg(phi + l - 2 * st);
    if (v[i] == n - 1) e = per ? __ldg(phi + l + (long long)(3 - v[i]) * st) : d;    // upperL!: ϕ[CIj(i,I,3)] / mirrored
    else e = __ldg(phi + l + 2 * st);
    acc = acc + gradphi2(a, b, c, d, e, s);
  }
  L[l] = pi * (T(1) - t_sqrt(acc));
}

// ϕ[I] = α·ϕ⁰[I] + (1-α)·(ϕ[I] + dτ·L[I]) on inside(ϕ), redistaning.jl:33
template <class T, int D> __global__ void redist_stage_kernel(T* __restrict__ phi, const T* __restrict__ phi0, const T* __restrict__ L, const Geo g, T dtau,
                                                              T alpha) {
  const int x = 2 + blockIdx.x * blockDim.x + threadIdx.x, y = 2 + blockIdx.y, z = (D == 3) ? 2 + blockIdx.z : 1;
  if (x > g.n[0] - 1) return;
  const long long l = lin3(g, x, y, z);
  phi[l] = alpha * __ldg(phi0 + l) + (T(1) - alpha) * (phi[l] + dtau * __ldg(L + l));
}

// LevelSet(sim): ϕ = 2f-1; ϕini .= ϕ over all entries, redistaning.jl:21-23
template <class T> __global__ void levelset_kernel(T* __restrict__ phi, T* __restrict__ pini, const T* __restrict__ f, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T v = T(2) * __ldg(f + i) - T(1);
  phi[i] = v;
  pini[i] = v;
}

// Σ over inside(f): out[0] = ρkeI, out[1] = ρgh, out[2+i] = ρuI(i)   (metrics.jl:15-17,25,49-51)
template <class T, int D> __global__ void __launch_bounds__(256) metrics_kernel(const T* __restrict__ u, const T* __restrict__ f, const Geo g, T lr, T U0,
                                                                                T U1, T U2, T G0, T G1, T G2, T W0, T W1, T W2, double* out) {
  const int ny = g.n[1] - 2, nz = (D == 3) ? g.n[2] - 2 : 1;
  const long long rows = (long long)ny * nz;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wid = threadIdx.x >> 5;
  const T U[3] = {U0, U1, U2}, G[3] = {G0, G1, G2}, W[3] = {W0, W1, W2};
  const T omlr = T(1) - lr;
  double s[5] = {0, 0, 0, 0, 0};
  for (long long r = (long long)blockIdx.x * wpb + wid; r < rows; r += (long long)gridDim.x * wpb) {
    const int y = 2 + (int)(r % ny), z = (D == 3) ? 2 + (int)(r / ny) : 1;
    const long long l0 = lin3(g, 0, y, z);
    for (int x = 2 + lane; x <= g.n[0] - 1; x += 32) {
      const long long l = l0 + x;
      const int v[3] = {x, y, z};
      const T rho = lin_interp(__ldg(f + l), lr, omlr);
      T ke = T(0), gh = T(0);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const long long st = (i == 0) ? 1 : ((i == 1) ? g.s1 : g.s2);
        const T ul = __ldg(u + (long long)i * g.S + l), uh = __ldg(u + (long long)i * g.S + l + st);
        const T a = ul - U[i], b = uh - U[i];
        ke += (a * a + b * b) * rho;
        gh += G[i] * ((T(v[i]) - T(1.5)) - W[i]);
        s[2 + i] += 0.5 * (double)(ul + uh - T(2) * U[i]) * (double)rho;
      }
      s[0] += 0.25 * (double)ke;
      s[1] += (double)(-rho * gh);
    }
  }
  __shared__ double ws[8][5];
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    double t = s[q];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if (lane == 0) ws[wid][q] = t;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double t = 0.0;
    for (int w = 0; w < wpb; ++w) t += ws[w][threadIdx.x];
    atomicAdd(out + threadIdx.x, t);
  }
}

// Σ EnsI over the inside cells (metrics.jl:34-41); ω: vector field in 3-D, scalar field in 2-D
template <class T, int D> __global__ void __launch_bounds__(256) enstrophy_kernel(const T* __restrict__ om, const Geo g, double* out) {
  const int ny = g.n[1] - 2, nz = (D == 3) ? g.n[2] - 2 : 1;
  const long long rows = (long long)ny * nz;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wid = threadIdx.x >> 5;
  double s = 0.0;
  for (long long r = (long long)blockIdx.x * wpb + wid; r < rows; r += (long long)gridDim.x * wpb) {
    const int y = 2 + (int)(r % ny), z = (D == 3) ? 2 + (int)(r / ny) : 1;
    const long long l0 = lin3(g, 0, y, z);
    for (int x = 2 + lane; x <= g.n[0] - 1; x += 32) {
      const long long l = l0 + x;
      T t = T(0);
      if (D == 3) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int ix = (i + 1) % 3, iy = (i + 2) % 3;  // shiftDir(i,3,1), shiftDir(i,3,2)
          const long long sx = (ix == 0) ? 1 : ((ix == 1) ? g.s1 : g.s2), sy = (iy == 0) ? 1 : ((iy == 1) ? g.s1 : g.s2);
          const T* o = om + (long long)i * g.S + l;
          const T a = __ldg(o), b = __ldg(o + sx), c = __ldg(o + sy), d = __ldg(o + sx + sy);
          t += a * a + b * b + c * c + d * d;
        }
      } else {
        const T a = __ldg(om + l), b = __ldg(om + l + 1), c = __ldg(om + l + g.s1), d = __ldg(om + l + 1 + g.s1);
        t = a * a + b * b + c * c + d * d;
      }
      s += 0.5 * 0.25 * (double)t;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  __shared__ double ws[8];
  if (lane == 0) ws[wid] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < wpb; ++w) t += ws[w];
    atomicAdd(out, t);
  }
}

}  // namespace ifadv
