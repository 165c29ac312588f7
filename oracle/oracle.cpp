// oracle/oracle.cpp -- TEST INFRASTRUCTURE ONLY.
// C ABI over HOST pointers for the CPU restatement of the reference algorithm (oracle_*.hpp).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library; the product path (libifadv_b200.so) never does.
// dtype: 0 = Float32, 1 = Float64.  Directions, sweep order (dirO) and cell indices are 1-based as in
// the Julia reference; perdir is a bit mask (bit j-1 set <=> j ∈ perdir).
#include "oracle_flow.hpp"
#include "oracle_forcing.hpp"
#include "oracle_post.hpp"
#include "oracle_poisson.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

#define DISPATCH(dtype, ...)       \
  do {                             \
    if ((dtype) == 0) {            \
      using T = float;             \
      __VA_ARGS__;                 \
    } else if ((dtype) == 1) {     \
      using T = double;            \
      __VA_ARGS__;                 \
    } else                         \
      return -2;                   \
  } while (0)

static inline I3 mkI(int D, const int64_t* I) { return I3{{I[0], I[1], D == 3 ? I[2] : 1}}; }

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// the timing build ignores an inherited OMP_NUM_THREADS=1 (torchrun exports it): the CPU legs state their thread count explicitly
int orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

int orc_get_intercept(int dtype, int D, const double* n, double g, double* out) {
  DISPATCH(dtype, {
    T nn[3] = {(T)n[0], (T)n[1], D == 3 ? (T)n[2] : T(0)};
    *out = (double)getInterceptD<T>(D, nn, (T)g);
  });
  return 0;
}
int orc_get_volume_fraction(int dtype, int D, const double* n, double b, double* out) {
  DISPATCH(dtype, {
    T nn[3] = {(T)n[0], (T)n[1], D == 3 ? (T)n[2] : T(0)};
    *out = (double)getVolumeFractionD<T>(D, nn, (T)b);
  });
  return 0;
}
int orc_limiter(int dtype, int lam, double u, double c, double d, double* out) {
  DISPATCH(dtype, { *out = (double)limiter<T>(lam, (T)u, (T)c, (T)d); });
  return 0;
}
int orc_normal(int dtype, int D, const int64_t* Ng, int scheme, void* f, void* nhat, const int64_t* I) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, {
    SF<T> ff{(T*)f, &g};
    VF<T> nh{(T*)nhat, &g};
    normalScheme<T>(scheme, D, ff, nh, mkI(D, I));
  });
  return 0;
}
int orc_vof_flux_face(int dtype, int D, const int64_t* Ng, void* ff, void* f, void* alpha, void* nhat, double dl, int d,
                      const int64_t* IFace, void* rhouf, double lr) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, {
    getVOFFlux_face<T>(g, SF<T>{(T*)ff, &g}, SF<T>{(T*)f, &g}, SF<T>{(T*)alpha, &g}, VF<T>{(T*)nhat, &g}, (T)dl, d - 1, mkI(D, IFace),
                       VF<T>{(T*)rhouf, &g}, (T)lr);
  });
  return 0;
}
int orc_bcf(int dtype, int D, const int64_t* Ng, void* f, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, BCf<T>(g, SF<T>{(T*)f, &g}, perdir));
  return 0;
}
int orc_bcv1d(int dtype, int D, const int64_t* Ng, void* f, int d, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, BCv1D<T>(g, SF<T>{(T*)f, &g}, d - 1, perdir));
  return 0;
}
int orc_bcv(int dtype, int D, const int64_t* Ng, void* f, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, BCv<T>(g, VF<T>{(T*)f, &g}, perdir));
  return 0;
}
int orc_bcvof(int dtype, int D, const int64_t* Ng, void* f, void* alpha, void* nhat, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, BCVOF<T>(g, SF<T>{(T*)f, &g}, SF<T>{(T*)alpha, &g}, VF<T>{(T*)nhat, &g}, perdir));
  return 0;
}
int orc_bc_vec(int dtype, int D, const int64_t* Ng, void* a, const double* A, int saveexit, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, {
    T AA[3] = {(T)A[0], (T)A[1], D == 3 ? (T)A[2] : T(0)};
    BC_vec<T>(g, VF<T>{(T*)a, &g}, AA, saveexit != 0, perdir);
  });
  return 0;
}
int orc_clean_wisp(int dtype, int D, const int64_t* Ng, void* f) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, cleanWisp<T>(g, SF<T>{(T*)f, &g}, 10 * std::numeric_limits<T>::epsilon()));
  return 0;
}
int orc_u2rhou(int dtype, int D, const int64_t* Ng, void* rhou, void* u, void* f, double lr) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, u2rhou<T>(g, VF<T>{(T*)rhou, &g}, VF<T>{(T*)u, &g}, SF<T>{(T*)f, &g}, (T)lr));
  return 0;
}
int orc_rhou2u(int dtype, int D, const int64_t* Ng, void* u, void* rhou, void* f, double lr) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, rhou2u<T>(g, VF<T>{(T*)u, &g}, VF<T>{(T*)rhou, &g}, SF<T>{(T*)f, &g}, (T)lr));
  return 0;
}
int orc_f2face(int dtype, int D, const int64_t* Ng, void* fFace, void* fCen, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, f2face<T>(g, VF<T>{(T*)fFace, &g}, SF<T>{(T*)fCen, &g}, perdir));
  return 0;
}
int orc_apply_vof_samples(int dtype, int D, const int64_t* Ng, void* f, void* alpha, void* nhat, const void* sc, const void* sp,
                          const void* sm) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, applyVOF_samples<T>(g, SF<T>{(T*)f, &g}, SF<T>{(T*)alpha, &g}, VF<T>{(T*)nhat, &g}, (const T*)sc, (const T*)sp,
                                      (const T*)sm));
  return 0;
}
int orc_reconstruct_interface(int dtype, int D, const int64_t* Ng, void* f, void* alpha, void* nhat, int scheme, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, reconstructInterface<T>(g, SF<T>{(T*)f, &g}, SF<T>{(T*)alpha, &g}, VF<T>{(T*)nhat, &g}, scheme, perdir));
  return 0;
}
int orc_get_vof_flux(int dtype, int D, const int64_t* Ng, void* ff, void* f, void* alpha, void* nhat, void* u, void* u0, double dt, int d,
                     void* rhouf, double lr) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, getVOFFlux<T>(g, SF<T>{(T*)ff, &g}, SF<T>{(T*)f, &g}, SF<T>{(T*)alpha, &g}, VF<T>{(T*)nhat, &g}, VF<T>{(T*)u, &g},
                                VF<T>{(T*)u0, &g}, (T)dt, d - 1, VF<T>{(T*)rhouf, &g}, (T)lr));
  return 0;
}

// advectVOF!  (src/advection.jl:34).  report may be NULL.
int orc_advect_vof(int dtype, int D, const int64_t* Ng, void* f, void* ff, void* alpha, void* nhat, void* u, void* u0, double dt,
                   int8_t* cbar, void* rhouf, double lr, int scheme, unsigned perdir, const int* dirO, FillReport* rep) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, return advectVOF<T>(g, (T*)f, (T*)ff, (T*)alpha, (T*)nhat, (T*)u, (T*)u0, (T)dt, cbar, (T*)rhouf, (T)lr, scheme, perdir,
                                      dirO, rep));
  return 0;
}
// advectVOFρuu!  (src/flow.jl:165).  uStar may alias nhat and dilaU may alias alpha (flow.jl:157-160).
int orc_advect_vof_rhouu(int dtype, int D, const int64_t* Ng, void* f, void* ff, void* alpha, void* nhat, void* u, void* u0, double dt,
                         int8_t* cbar, void* rhou, void* r, void* Phi, void* rhouf, void* uStar, void* uOld, void* dilaU, void* drho,
                         double lr, int limiter_id, int scheme, const double* uBC, unsigned perdir, int exitBC, const int* dirO,
                         FillReport* rep) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, {
    T A[3] = {(T)uBC[0], (T)uBC[1], D == 3 ? (T)uBC[2] : T(0)};
    return advectVOFrhouu<T>(g, (T*)f, (T*)ff, (T*)alpha, (T*)nhat, (T*)u, (T*)u0, (T)dt, cbar, (T*)rhou, (T*)r, (T*)Phi, (T*)rhouf,
                             (T*)uStar, (T*)uOld, (T*)dilaU, (T*)drho, (T)lr, limiter_id, scheme, A, perdir, exitBC != 0, dirO, rep);
  });
  return 0;
}
// one directional sweep (number iop of the call) of advectVOFρuu! -- see advectVOFrhouu's only_op
int orc_advect_vof_rhouu_sweep(int dtype, int D, const int64_t* Ng, void* f, void* ff, void* alpha, void* nhat, void* u, void* u0, double dt,
                               int8_t* cbar, void* rhou, void* r, void* Phi, void* rhouf, void* uStar, void* uOld, void* dilaU, void* drho,
                               double lr, int limiter_id, int scheme, const double* uBC, unsigned perdir, int exitBC, const int* dirO,
                               FillReport* rep, int iop) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, {
    T A[3] = {(T)uBC[0], (T)uBC[1], D == 3 ? (T)uBC[2] : T(0)};
    return advectVOFrhouu<T>(g, (T*)f, (T*)ff, (T*)alpha, (T*)nhat, (T*)u, (T*)u0, (T)dt, cbar, (T*)rhou, (T*)r, (T*)Phi, (T*)rhouf,
                             (T*)uStar, (T*)uOld, (T*)dilaU, (T*)drho, (T)lr, limiter_id, scheme, A, perdir, exitBC != 0, dirO, rep, iop);
  });
  return 0;
}
// MPCFL (src/flow.jl:262).  mu<=0 / eta<=0 / gnorm<=0 disable the corresponding limit.
int orc_mpcfl(int dtype, int D, const int64_t* Ng, void* u, void* sigma, double nu, double mu, double lam_mu, double lam_rho, double eta,
              double gnorm, double dt_max, double safety, double* out) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, {
    *out = (double)MPCFL<T>(g, VF<T>{(T*)u, &g}, SF<T>{(T*)sigma, &g}, (T)nu, (T)mu, (T)lam_mu, (T)lam_rho, (T)eta, (T)gnorm, (T)dt_max,
                            (T)safety);
  });
  return 0;
}
// sum(f[inside(f)]) in Float64
int orc_sum_inside(int dtype, int D, const int64_t* Ng, const void* f, double* out) {
  Grid g = make_grid(D, Ng);
  double s = 0;
  Range r = r_inside(g);
  DISPATCH(dtype, {
    const T* p = (const T*)f;
    for (int64_t k = r.lo[2]; k <= r.hi[2]; ++k)
      for (int64_t j = r.lo[1]; j <= r.hi[1]; ++j)
        for (int64_t i = r.lo[0]; i <= r.hi[0]; ++i) s += (double)p[lin(g, I3{{i, j, k}})];
  });
  *out = s;
  return 0;
}

// ---- explicit forcing: src/flow.jl:113-153,244-259, src/VOFutil.jl:186-191, src/surfaceTension.jl ----------------
// getμ(i,j,I,fFace,λμ,μ,λρ); i, j, I 1-based
int orc_getmu(int dtype, int D, const int64_t* Ng, void* fFace, int i, int j, const int64_t* I, double lmu, double mu, double lr, double* out) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, { *out = (double)getmu<T>(i - 1, j - 1, mkI(D, I), VF<T>{(T*)fFace, &g}, (T)lmu, (T)mu, (T)lr); });
  return 0;
}
// getPopinetHeight(I,f,i) / getCurvature(I,f,i); i: signed 1-based direction
int orc_popinet_height(int dtype, int D, const int64_t* Ng, void* f, const int64_t* I, int i, double* out) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, { *out = (double)getPopinetHeight<T>(g, mkI(D, I), SF<T>{(T*)f, &g}, i); });
  return 0;
}
int orc_curvature(int dtype, int D, const int64_t* Ng, void* f, const int64_t* I, int i, double* out) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, { *out = (double)getCurvature<T>(g, mkI(D, I), SF<T>{(T*)f, &g}, i); });
  return 0;
}
// viscSurfTenρu!(r,u,Φ,f,α,n̂,fbuffer,λμ,μ,λρ,η;perdir); has_mu / has_eta = 0 stand for `nothing`
int orc_visc_surften_rhou(int dtype, int D, const int64_t* Ng, void* r, void* u, void* Phi, void* f, void* alpha, void* nhat, void* fbuffer,
                          double lmu, double mu, int has_mu, double lr, double eta, int has_eta, unsigned perdir) {
  (void)alpha;
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, {
    viscSurfTenRhou<T>(g, VF<T>{(T*)r, &g}, VF<T>{(T*)u, &g}, SF<T>{(T*)Phi, &g}, SF<T>{(T*)f, &g}, VF<T>{(T*)nhat, &g}, SF<T>{(T*)fbuffer, &g},
                       (T)lmu, (T)mu, has_mu != 0, (T)lr, (T)eta, has_eta != 0, perdir);
  });
  return 0;
}
// updateU!(u,ρu,ρu⁰,forcing,dt,f,λρ,tNow,g,uBC,w) with a constant gravity vector (grav == NULL: g = nothing)
int orc_update_u(int dtype, int D, const int64_t* Ng, void* u, void* rhou, void* rhou0, void* forcing, double dt, void* f, double lr,
                 const double* grav, double w) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, {
    T gg[3] = {0, 0, 0};
    if (grav) for (int i = 0; i < D; ++i) gg[i] = (T)grav[i];
    updateU<T>(g, VF<T>{(T*)u, &g}, VF<T>{(T*)rhou, &g}, VF<T>{(T*)rhou0, &g}, VF<T>{(T*)forcing, &g}, (T)dt, SF<T>{(T*)f, &g}, (T)lr,
               grav ? gg : nullptr, (T)w);
  });
  return 0;
}
// updateL!(μ₀,f,λρ;perdir)
int orc_update_l(int dtype, int D, const int64_t* Ng, void* mu0, void* f, double lr, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, { updateL<T>(g, VF<T>{(T*)mu0, &g}, SF<T>{(T*)f, &g}, (T)lr, perdir); });
  return 0;
}

// ---- post-processing: src/redistaning.jl, src/metrics.jl ---------------------------------------------------------
int orc_gradphi2(int dtype, double a, double b, double c, double d, double e, double s, double* out) {
  DISPATCH(dtype, { *out = (double)gradphi2<T>((T)a, (T)b, (T)c, (T)d, (T)e, (T)s); });
  return 0;
}
int orc_compute_l(int dtype, int D, const int64_t* Ng, void* L, void* phi, void* phi_ini, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, { computeL<T>(g, SF<T>{(T*)L, &g}, SF<T>{(T*)phi, &g}, SF<T>{(T*)phi_ini, &g}, perdir); });
  return 0;
}
int orc_redist_stage(int dtype, int D, const int64_t* Ng, void* phi, void* phi0, void* phi_ini, void* L, double dtau, double alpha, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, { redistStage<T>(g, SF<T>{(T*)phi, &g}, SF<T>{(T*)phi0, &g}, SF<T>{(T*)phi_ini, &g}, SF<T>{(T*)L, &g}, (T)dtau, (T)alpha, perdir); });
  return 0;
}
int orc_redistance(int dtype, int D, const int64_t* Ng, void* phi, void* phi0, void* phi_ini, void* L, double d, double dtau, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, { redistance<T>(g, SF<T>{(T*)phi, &g}, SF<T>{(T*)phi0, &g}, SF<T>{(T*)phi_ini, &g}, SF<T>{(T*)L, &g}, d, dtau, perdir); });
  return 0;
}
// per-cell metrics (I 1-based; i 1-based; NULL tuples are zeros) and their sums over inside(f): out = [Σρke, Σρgh, Σρu_1..D]
int orc_metrics_cell(int dtype, int D, const int64_t* Ng, void* u, void* f, double lr, const double* U, const double* grav, const double* statWL,
                     const int64_t* I, double* out) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, {
    T UU[3] = {0, 0, 0}, G[3] = {0, 0, 0}, W[3] = {0, 0, 0};
    for (int i = 0; i < D; ++i) { if (U) UU[i] = (T)U[i]; if (grav) G[i] = (T)grav[i]; if (statWL) W[i] = (T)statWL[i]; }
    const I3 II = mkI(D, I);
    out[0] = rhokeI<T>(g, II, VF<T>{(T*)u, &g}, SF<T>{(T*)f, &g}, (T)lr, UU);
    out[1] = (double)rhogh<T>(g, II, G, SF<T>{(T*)f, &g}, (T)lr, W);
    for (int i = 0; i < D; ++i) out[2 + i] = rhouI<T>(g, i, II, VF<T>{(T*)u, &g}, SF<T>{(T*)f, &g}, (T)lr, UU);
  });
  return 0;
}
int orc_metrics_sum(int dtype, int D, const int64_t* Ng, void* u, void* f, double lr, const double* U, const double* grav, const double* statWL,
                    double* out) {
  Grid g = make_grid(D, Ng);
  for (int k = 0; k < 5; ++k) out[k] = 0;
  Range r = r_inside(g);
  DISPATCH(dtype, {
    T UU[3] = {0, 0, 0}, G[3] = {0, 0, 0}, W[3] = {0, 0, 0};
    for (int i = 0; i < D; ++i) { if (U) UU[i] = (T)U[i]; if (grav) G[i] = (T)grav[i]; if (statWL) W[i] = (T)statWL[i]; }
    VF<T> uu{(T*)u, &g};
    SF<T> ff{(T*)f, &g};
    for (int64_t k = r.lo[2]; k <= r.hi[2]; ++k)
      for (int64_t j = r.lo[1]; j <= r.hi[1]; ++j)
        for (int64_t i = r.lo[0]; i <= r.hi[0]; ++i) {
          const I3 II{{i, j, k}};
          out[0] += rhokeI<T>(g, II, uu, ff, (T)lr, UU);
          out[1] += (double)rhogh<T>(g, II, G, ff, (T)lr, W);
          for (int c = 0; c < D; ++c) out[2 + c] += rhouI<T>(g, c, II, uu, ff, (T)lr, UU);
        }
  });
  return 0;
}
int orc_enstrophy(int dtype, int D, const int64_t* Ng, void* omega, const int64_t* I, double* cell, double* sum) {
  Grid g = make_grid(D, Ng);
  Range r = r_inside(g);
  DISPATCH(dtype, {
    if (I && cell) *cell = EnsI<T>(g, mkI(D, I), (const T*)omega);
    if (sum) {
      double s = 0;
      for (int64_t k = r.lo[2]; k <= r.hi[2]; ++k)
        for (int64_t j = r.lo[1]; j <= r.hi[1]; ++j)
          for (int64_t i = r.lo[0]; i <= r.hi[0]; ++i) s += EnsI<T>(g, I3{{i, j, k}}, (const T*)omega);
      *sum = s;
    }
  });
  return 0;
}

// ---- pressure projection: src/flow.jl:300-347 on WaterLily's Poisson struct ---------------------------------------
#define ORC_POIS(T, g) Pois<T>{VF<T>{(T*)L, &g}, SF<T>{(T*)Dg, &g}, SF<T>{(T*)iD, &g}, SF<T>{(T*)x, &g}, SF<T>{(T*)eps, &g}, SF<T>{(T*)r, &g}, SF<T>{(T*)z, &g}, perdir}
int orc_pois_update(int dtype, int D, const int64_t* Ng, void* Dg, void* iD, void* L) {
  Grid g = make_grid(D, Ng);
  void *x = nullptr, *eps = nullptr, *r = nullptr, *z = nullptr;
  const unsigned perdir = 0;
  DISPATCH(dtype, { pois_update<T>(g, ORC_POIS(T, g)); });
  return 0;
}
// mult!(p,x): perBC!(x); z = A x on inside(x) (ghost entries of z zeroed)
int orc_pois_mult(int dtype, int D, const int64_t* Ng, void* z, void* x, void* L, void* Dg, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  DISPATCH(dtype, {
    SF<T> zz{(T*)z, &g}, xx{(T*)x, &g}, dd{(T*)Dg, &g};
    VF<T> ll{(T*)L, &g};
    perBC<T>(g, xx, perdir);
    for (int64_t k = 0; k < g.S; ++k) zz.p[k] = 0;
    loop(r_inside(g), [&](I3 I) { zz(I) = pois_mult<T>(g, I, ll, dd, xx); });
  });
  return 0;
}
int orc_pois_residual(int dtype, int D, const int64_t* Ng, void* x, void* r, void* z, void* L, void* Dg, void* iD, unsigned perdir) {
  Grid g = make_grid(D, Ng);
  void* eps = nullptr;
  DISPATCH(dtype, { pois_residual<T>(g, ORC_POIS(T, g)); });
  return 0;
}
// tol < 0: 50eps(T); itmx <= 0: 6000 (the defaults of flow.jl:300).  Returns the iteration count (>= 0)
int orc_psolver(int dtype, int D, const int64_t* Ng, void* x, void* eps, void* r, void* z, void* L, void* Dg, void* iD, unsigned perdir,
                double tol, int itmx, double* r2) {
  Grid g = make_grid(D, Ng);
  int np = 0;
  DISPATCH(dtype, {
    const T t = tol < 0 ? T(50) * std::numeric_limits<T>::epsilon() : (T)tol;
    np = psolver<T>(g, ORC_POIS(T, g), t, itmx <= 0 ? 6000 : itmx, r2);
  });
  return np;
}
int orc_myproject(int dtype, int D, const int64_t* Ng, void* u, void* x, void* eps, void* r, void* z, void* L, void* Dg, void* iD,
                  unsigned perdir, double dt, double* r2) {
  Grid g = make_grid(D, Ng);
  int np = 0;
  DISPATCH(dtype, { np = myproject<T>(g, VF<T>{(T*)u, &g}, ORC_POIS(T, g), (T)dt, r2); });
  return np;
}

// ---- WaterLily.MultiLevelPoisson: opaque handle (levels >= 2 and level 1's D, iD, ϵ, r are owned by the handle) -------------------------
struct OrcML { int dtype; void* h; };
void* orc_ml_create(int dtype, int D, const int64_t* Ng, void* x, void* L, void* z, unsigned perdir, int maxlevels) {
  if (dtype == 0) return new OrcML{0, ml_create<float>(D, Ng, (float*)x, (float*)L, (float*)z, perdir, maxlevels)};
  if (dtype == 1) return new OrcML{1, ml_create<double>(D, Ng, (double*)x, (double*)L, (double*)z, perdir, maxlevels)};
  return nullptr;
}
#define ML_DISPATCH(m, ...)                          \
  do {                                               \
    if ((m)->dtype == 0) {                           \
      using T = float;                               \
      auto& ml = *(MLPois<float>*)(m)->h;            \
      __VA_ARGS__;                                   \
    } else {                                         \
      using T = double;                              \
      auto& ml = *(MLPois<double>*)(m)->h;           \
      __VA_ARGS__;                                   \
    }                                                \
  } while (0)
int orc_ml_destroy(void* m_) {
  OrcML* m = (OrcML*)m_;
  if (!m) return 0;
  if (m->dtype == 0) delete (MLPois<float>*)m->h; else delete (MLPois<double>*)m->h;
  delete m;
  return 0;
}
int orc_ml_levels(void* m_) {
  OrcML* m = (OrcML*)m_;
  int n = 0;
  ML_DISPATCH(m, { (void)sizeof(T); n = (int)ml.lv.size(); });
  return n;
}
// which: 0 L, 1 D, 2 iD, 3 x, 4 ϵ, 5 r, 6 z; level 0-based
int orc_ml_level_array(void* m_, int level, int which, void** ptr, int64_t* Ng) {
  OrcML* m = (OrcML*)m_;
  ML_DISPATCH(m, {
    (void)sizeof(T);
    if (level < 0 || level >= (int)ml.lv.size()) return -2;
    auto& lv = *ml.lv[level];
    for (int d = 0; d < 3; ++d) Ng[d] = lv.g.n[d];
    void* t[7] = {lv.p.L.p, lv.p.D.p, lv.p.iD.p, lv.p.x.p, lv.p.eps.p, lv.p.r.p, lv.p.z.p};
    *ptr = t[which];
  });
  return 0;
}
int orc_ml_update(void* m_) {
  OrcML* m = (OrcML*)m_;
  ML_DISPATCH(m, { (void)sizeof(T); ml_update(ml); });
  return 0;
}
int orc_ml_vcycle(void* m_) {
  OrcML* m = (OrcML*)m_;
  ML_DISPATCH(m, { (void)sizeof(T); ml_vcycle(ml); });
  return 0;
}
int orc_ml_smooth(void* m_, int level) {
  OrcML* m = (OrcML*)m_;
  ML_DISPATCH(m, { (void)sizeof(T); pois_pcg_smooth(ml.lv[level]->g, ml.lv[level]->p); });
  return 0;
}
int orc_ml_residual(void* m_) {
  OrcML* m = (OrcML*)m_;
  ML_DISPATCH(m, { (void)sizeof(T); pois_residual(ml.lv[0]->g, ml.lv[0]->p); });
  return 0;
}
// solver!(ml;tol,itmx): tol < 0 -> 1e-4, itmx <= 0 -> 32 (WaterLily's defaults)
int orc_ml_solver(void* m_, double tol, int itmx, double* r2) {
  OrcML* m = (OrcML*)m_;
  int np = 0;
  ML_DISPATCH(m, { np = ml_solver<T>(ml, tol < 0 ? T(1e-4) : (T)tol, itmx <= 0 ? 32 : itmx, r2); });
  return np;
}
int orc_ml_myproject(void* m_, void* u, double dt, double* r2) {
  OrcML* m = (OrcML*)m_;
  int np = 0;
  ML_DISPATCH(m, { np = ml_myproject<T>(ml, VF<T>{(T*)u, &ml.lv[0]->g}, (T)dt, r2); });
  return np;
}

}  // extern "C"
