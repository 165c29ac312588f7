/* include/ifadv.h -- C ABI of libifadv_b200.so: the B200-native (sm_100a) VOF + CMOM advection path of
 * InterfaceAdvection.jl.  This is the drop-in boundary: a Julia package extension (julia/IntfAdvB200Ext.jl,
 * see INTEGRATION.md) `ccall`s these entry points in place of the KernelAbstractions `@loop` kernels that
 * ext/IntfAdvCUDAExt.jl enables today.  Plain pointers and sizes only; no torch / CUDA.jl types.
 *
 * Conventions (identical to the reference's arrays, SURVEY.md §8 "Conventions for sizes"):
 *   - arrays are column-major with ONE ghost layer per side: scalar field extents Ng = N .+ 2,
 *     vector field (Ng..., D) with the component index slowest (SoA);
 *   - u[I,d] lives on the lower d-face of cell I;
 *   - directions, sweep orders (dirO) and reported cell indices are 1-based like Julia;
 *   - perdir_mask has bit (j-1) set iff direction j is periodic;
 *   - dtype: IFADV_F32 = Float32, IFADV_F64 = Float64; scalars cross the ABI as double and are rounded to T;
 *   - every pointer named in a *_dev / plain entry point is a DEVICE pointer borrowed for the call;
 *     all work is enqueued on the caller's stream (a cudaStream_t passed as void*); no device-wide sync;
 *   - return value: 0 ok; >0 advisory (bit0 overfill, bit1 underfill; only evaluated when report != NULL);
 *     <0 fatal: -1 NaN in f, -2 invalid argument / unsupported combination, -3 CUDA error, -4 comm error,
 *     -5 "divergence ... is exploding" (|∇·u⁰|+|∇·u| > 10 or NaN at an over/under-filled cell, src/advection.jl:160,180).
 *
 * Preconditions shared with the reference's call sites: f, u, u0 carry valid ghost values (BCf!/BC! applied,
 * as they always are inside MPFMomStep!, flow.jl:61-92).
 */
#ifndef IFADV_H
#define IFADV_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ifadv_ctx ifadv_ctx;

enum { IFADV_F32 = 0, IFADV_F64 = 1 };

/* normalScheme function identity -> enum (src/normalEstimation.jl) */
enum {
  IFADV_NS_WH = 0,     /* getInterfaceNormal_WH!     :105-154 (default) */
  IFADV_NS_WY = 1,     /* getInterfaceNormal_WY!     :35-68  */
  IFADV_NS_COLUMN = 2, /* getInterfaceNormal_Column! :75-97  */
  IFADV_NS_PCD = 3,    /* getInterfaceNormal_PCD!    :161-165 */
  IFADV_NS_SLIC = 4,   /* getInterfaceNormal_SLIC!   :172-176 */
  IFADV_NS_MYC = 5,    /* getInterfaceNormal_MYC!    :185-202 */
  IFADV_NS_YOUNGS = 6, /* getInterfaceNormal_Y!      :232-247 */
  IFADV_NS_CD = 7,     /* getInterfaceNormal_CD!     :266-270 */
  IFADV_NS_XYLIC = 8   /* getInterfaceNormal_XYLIC!  :279-283 */
};
/* limiter λ(u,c,d) function identity -> enum (src/flow.jl:5-15; 8-10 are WaterLily's) */
enum {
  IFADV_LIM_UPWIND = 0, IFADV_LIM_MINMOD = 1, IFADV_LIM_KOREN = 2, IFADV_LIM_VANALBADA1 = 3, IFADV_LIM_SWEBY = 4,
  IFADV_LIM_SUPERBEE = 5, IFADV_LIM_TVDCEN = 6, IFADV_LIM_TVDDOWN = 7, IFADV_LIM_QUICK = 8, IFADV_LIM_VANLEER = 9,
  IFADV_LIM_CDS = 10
};
/* flags */
enum {
  IFADV_NO_RHOUF = 1 /* advect_vof: do not materialise ρuf (the un-normalised mass flux `advect!` users may read) */
};

/* replaces reportFillError's findmax/findmin (src/advection.jl:145-149).  Filled only when passed non-NULL
 * (the call then synchronises the stream). */
typedef struct {
  double maxf, minf;             /* extreme values of f after the worst (or last) directional sweep, before cleanWisp! */
  int64_t argmax[3], argmin[3];  /* 1-based cell index (approximate tie-breaking) */
  int dir;                       /* 1-based sweep direction the values belong to */
  int status;                    /* same bits as the return value */
  double div_u0, div_u;          /* |∇·u⁰|, |∇·u| at the reported cell (reportFillError's diagnostics, src/advection.jl:151,170);
                                    0 unless status > 0 */
} ifadv_report;

/* ---- context ------------------------------------------------------------------------------------------- */
/* One context per (device, grid, dtype).  Owns reduction buffers and the pinned report; NOT the fields.
 * Ng: array extents INCLUDING ghosts (N .+ 2), Ng[2] ignored for D == 2. */
int ifadv_create(ifadv_ctx** ctx, int D, const int64_t Ng[3], int dtype, int device);
int ifadv_destroy(ifadv_ctx* ctx);
const char* ifadv_last_error(const ifadv_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t ifadv_launch_count(const ifadv_ctx* ctx);
const char* ifadv_version(void);
/* Measurement aid for bench.py: while enabled, every fused sweep launch is bracketed by CUDA events recorded on
 * the launching stream (up to 4096 launches).  ifadv_profile_read synchronises those events, returns the summed
 * kernel time in ms and the number of launches -- index 0: standard sweeps (13s+1 B/cell), index 1: fused first
 * sweeps of ifadv_u2rhou_advect_vof_rhouu (10s+1 B/cell); both arrays have 2 entries -- and resets the pool. */
int ifadv_profile(ifadv_ctx* ctx, int enable);
int ifadv_profile_read(ifadv_ctx* ctx, double* total_ms, int64_t* launches);
/* per sweep direction, accumulated by the ifadv_profile_read calls since the last call: 6 entries, index 2*j + fused
 * (j = 0-based sweep direction); resets the accumulators. */
int ifadv_profile_read_dirs(ifadv_ctx* ctx, double* total_ms, int64_t* launches);

/* ---- the hot path --------------------------------------------------------------------------------------- */
/* advectVOF!(f,fᶠ,α,n̂,u,u⁰,Δt,c̄,ρuf,λρ,normalScheme; perdir,dirO)            src/advection.jl:34-78
 * In/out f (all Ng entries, ghosts refreshed like BCf!).  fᶠ, α are used as ping-pong scratch; n̂ is untouched;
 * c̄ is rewritten; ρuf[·,d] receives δl·λρ+(1-λρ)fᶠ on inside_uWB faces (0 elsewhere) unless IFADV_NO_RHOUF. */
int ifadv_advect_vof(ifadv_ctx* ctx, void* stream, void* f, void* ff, void* alpha, void* nhat, const void* u, const void* u0,
                     double dt, int8_t* cbar, void* rhouf, double lambda_rho, int normal_scheme, unsigned perdir_mask,
                     const int dirO[3], int flags, ifadv_report* report);

/* advectVOFρuu!(f,fᶠ,α,n̂,u,u⁰,Δt,c̄,ρu,r,Φ,ρuf,uStar,uOld,dilaU,dρ,λρ,λ,normalScheme,uBC; perdir,exitBC,dirO)
 *                                                                              src/flow.jl:165-210
 * In/out f (all entries) and ρu (inside(f) entries); every other array is scratch exactly as in the
 * reference: fᶠ, Φ hold f ping-pong copies, r and ρuf hold ρu ping-pong copies, c̄ is
 * rewritten.  uStar/dilaU may alias n̂/α as in advectfq! (flow.jl:157-160); they are not touched.
 * dρ is only READ (plane Ng[j] of component j, which the reference never writes).  uBC: constant tuple.
 * exitBC = true (BC!'s saveexit at flow.jl:197,207): plane N of component 1 of u★ keeps the value the caller's r array
 * holds there at entry (the reference reads r[N,·,·,1] as left by its previous use: identical for the first sweep,
 * unspecified scratch afterwards -- DESIGN.md §5) and the exit face keeps its computed mass flux instead of uBC[1]. */
int ifadv_advect_vof_rhouu(ifadv_ctx* ctx, void* stream, void* f, void* ff, void* alpha, void* nhat, const void* u,
                           const void* u0, double dt, int8_t* cbar, void* rhou, void* r, void* Phi, void* rhouf, void* uStar,
                           const void* uOld, void* dilaU, const void* drho, double lambda_rho, int limiter, int normal_scheme,
                           const double uBC[3], unsigned perdir_mask, int exitBC, const int dirO[3], ifadv_report* report);

/* Fused form of the three calls MPFMomStep! makes back to back (src/flow.jl:61,69-70 and :89-92):
 *     f .= f_src;  u2ρu!(ρu,uOld,f,λρ);  BC!(ρu,uBC,exitBC,perdir);  advectVOFρuu!(f,…,ρu,…,uOld,…; exitBC)
 * Both call sites build ρu from the very velocity array they pass as uOld, so the first directional sweep can form
 * ρu = BC!(uOld·ρ(f̄)) on the fly: the u2ρu! pass, the BC! launch, the f⁰←f copy and three of the thirteen input streams
 * of sweep 1 disappear.  Results are bit-identical to the three separate calls.  f_src may equal f.  ρu is output only. */
int ifadv_u2rhou_advect_vof_rhouu(ifadv_ctx* ctx, void* stream, const void* f_src, void* f, void* ff, void* Phi, const void* u,
                                  const void* u0, double dt, int8_t* cbar, void* rhou, void* r, void* rhouf, const void* uOld,
                                  const void* drho, double lambda_rho, int limiter, int normal_scheme, const double uBC[3],
                                  unsigned perdir_mask, int exitBC, const int dirO[3], ifadv_report* report);

/* ---- secondary seams on the path ------------------------------------------------------------------------ */
/* u2ρu!(ρu,u,f,λρ) / ρu2u!(u,ρu,f,λρ)                                           src/VOFutil.jl:198-211 */
int ifadv_u2rhou(ifadv_ctx* ctx, void* stream, void* rhou, const void* u, const void* f, double lambda_rho);
int ifadv_rhou2u(ifadv_ctx* ctx, void* stream, void* u, const void* rhou, const void* f, double lambda_rho);
/* WaterLily.BC!(a,A,saveexit,perdir) for a constant tuple A                     used at src/flow.jl:69,91 */
int ifadv_bc_vec(ifadv_ctx* ctx, void* stream, void* a, const double A[3], int saveexit, unsigned perdir_mask);
/* BCf!(f;perdir)                                                                src/VOFutil.jl:64-75 */
int ifadv_bcf(ifadv_ctx* ctx, void* stream, void* f, unsigned perdir_mask);
/* out[i] = a*x[i] + b*y[i] over all Ng entries (the midpoint `@. f⁰ = (f⁰+f)/2`, src/flow.jl:74, is a=b=.5
 * evaluated as (x+y)*a when a == b) */
int ifadv_axpby(ifadv_ctx* ctx, void* stream, void* out, double a, const void* x, double b, const void* y);
/* MPCFL(a,c) (src/flow.jl:262-281).  mu/eta/gnorm <= 0 disable that limit (`nothing` in the reference).
 * Synchronises the stream; result in *dt_out. */
int ifadv_mpcfl(ifadv_ctx* ctx, void* stream, const void* u, double nu, double mu, double lambda_mu, double lambda_rho,
                double eta, double gnorm, double dt_max, double safety, double* dt_out);
/* sum(f[inside(f)]) accumulated in Float64 (total-mass check, test/maintests.jl:209,215).  Synchronises. */
int ifadv_sum_inside(ifadv_ctx* ctx, void* stream, const void* f, double* out);
/* applyVOF!(f,α,n̂,InterfaceSDF) (src/VOFutil.jl:9-37) with the SDF pre-sampled on the device at the cell
 * centres (sc) and at centre ± 0.01 e_i (sp, sm: vector fields); includes cleanWisp!; caller runs ifadv_bcf. */
int ifadv_apply_vof_samples(ifadv_ctx* ctx, void* stream, void* f, void* alpha, void* nhat, const void* sc, const void* sp,
                            const void* sm);

/* ---- explicit forcing between advection and projection (single-GPU contexts) ----------------------------
 * viscSurfTenρu!(r,u,Φ,f,α,n̂,fbuffer,λμ,μ,λρ,η;perdir)                          src/flow.jl:113-117
 *   = fill!(r,0); visc! (flow.jl:120-152, getμ VOFutil.jl:186-191); surfTen! (src/surfaceTension.jl:8-101: staggered f̄ = ϕ(d,·,f),
 *   Weymouth-Yue normal, Popinet column heights, height-function curvature).  Out: r on inside(f) (the reference also accumulates
 *   into upper ghost entries nothing reads).  mu <= 0 / eta <= 0 stand for `nothing`.  f, u carry valid ghosts.  Φ, α are unused;
 *   n̂ is only READ: plane N_d of component d, which f2face!+BCv! never write in a non-periodic direction (the reference reads it as
 *   left by BC! on u★≡n̂ in the previous advectfq!, i.e. uBC[d]); fbuffer is only READ as well: the staggered fields ϕ(d,·,f) with the
 *   ghost rules of BCf!(d,·) are evaluated from f where the stencils need them, and plane N_d of a non-periodic d, which BCf!(d,·)
 *   skips, resolves like in the reference's sequential passes (previous direction's field; the caller's values for d = 1).  The
 *   reference leaves ϕ(D,·,f) behind in fbuffer, which nothing reads; here the array is untouched. */
int ifadv_visc_surften_rhou(ifadv_ctx* ctx, void* stream, void* r, const void* u, void* Phi, const void* f, void* alpha, const void* nhat,
                            const void* fbuffer, double lambda_mu, double mu, double lambda_rho, double eta, unsigned perdir_mask);
/* updateU!(u,ρu,ρu⁰,forcing,dt,f,λρ,tNow,g,uBC,w)                                  src/flow.jl:244-252
 *   ρu ← (a ρu⁰ + ρu + forcing·dt)·w with a = 1/w - 1 (ALL entries); u ← ρu/ρ(f̄) on inside(f); forcing ← g; u ← u + dt·w·g.
 *   g: constant gravity vector (accelerate! with g(i,x,t) = g[i]) or NULL (`nothing`; time-dependent uBC is not supported). */
int ifadv_update_u(ifadv_ctx* ctx, void* stream, void* u, void* rhou, const void* rhou0, void* forcing, double dt, const void* f,
                   double lambda_rho, const double g[3], double w);
/* updateL!(μ₀,f,λρ;perdir): μ₀[I,d] /= ρ(f̄) on inside(f); BC!(μ₀,0,false,perdir)     src/flow.jl:254-259
 *   fill_one != 0 folds the fill!(a.μ₀,1) that precedes it in MPFMomStep! (flow.jl:73,96) into the same pass. */
int ifadv_update_l(ifadv_ctx* ctx, void* stream, void* mu0, const void* f, double lambda_rho, unsigned perdir_mask, int fill_one);

/* ---- post-processing (single-GPU contexts) ---------------------------------------------------------------
 * LevelSet(sim): ϕ = 2f-1, ϕini = ϕ over all entries                               src/redistaning.jl:8-29 (ϕ≡f⁰, ϕ⁰≡α, ϕini≡fᶠ, L≡σ) */
int ifadv_levelset_init(ifadv_ctx* ctx, void* stream, void* phi, void* phi_ini, const void* f);
/* computeL!(L,ϕ,ϕini;perdir): L = ϕini·(1-|∇ϕ|) on inside(ϕ), second-order ENO one-sided differences    src/redistaning.jl:67-143 */
int ifadv_redist_compute_l(ifadv_ctx* ctx, void* stream, void* L, const void* phi, const void* phi_ini, unsigned perdir_mask);
/* _redistaningStage!(ϕ,ϕ⁰,ϕini,L,dτ,α;perdir): computeL!; ϕ ← αϕ⁰+(1-α)(ϕ+dτL) on inside(ϕ)              src/redistaning.jl:31-34 */
int ifadv_redist_stage(ifadv_ctx* ctx, void* stream, void* phi, const void* phi0, const void* phi_ini, void* L, double dtau, double alpha,
                       unsigned perdir_mask);
/* redistaning!(ls; d, dτ, perdir): round(d/dτ) third-order SSP Runge-Kutta steps, BCf! after every stage       src/redistaning.jl:44-57 */
int ifadv_redistance(ifadv_ctx* ctx, void* stream, void* phi, void* phi0, const void* phi_ini, void* L, double d, double dtau,
                     unsigned perdir_mask);
/* Σ over inside(f) of ρkeI (out[0]), ρgh (out[1]) and ρuI(i) (out[2..]) accumulated in Float64                 src/metrics.jl:15-17,25,49-51
 * U (background flow), g (gravity tuple), statWL: constant tuples, NULL = zeros.  Synchronises the stream. */
int ifadv_metrics(ifadv_ctx* ctx, void* stream, const void* u, const void* f, double lambda_rho, const double U[3], const double g[3],
                  const double statWL[3], double out[5]);
/* Σ EnsI(I,ω) over the inside cells; ω: vector field in 3-D, scalar field in 2-D                              src/metrics.jl:34-41 */
int ifadv_enstrophy(ifadv_ctx* ctx, void* stream, const void* omega, double* out);

/* ---- pressure projection ----------------------------------------------------------------------------------
 * z-slab contexts: the dot products are all-reduced over the communicator, the ghost plane of ϵ (every iteration) and of x (at both
 * ends) comes from the z-neighbours, only the owned planes are updated; the caller refreshes the ghost planes of u afterwards
 * (ifadv_exchange_planes) and passes a perdir_mask without the z bit, as for the transport.
 * On WaterLily's Poisson arrays: L ≡ Flow.μ₀ (face coefficients, (Ng...,D)), x ≡ Flow.p, z ≡ Flow.σ, and the solver's own D, iD, ϵ, r
 * (scalar fields).  WaterLily's primitives (set_diag!, mult, perBC!, residual!, L₂) are restated from its published 1.x source.
 * update!(p::Poisson) = set_diag!(D,iD,L): D = -Σᵢ(L[I,i]+L[I+δᵢ,i]), iD = D² < 2eps ? 0 : 1/D on inside(x)     flow.jl:81,105 */
int ifadv_poisson_update(ifadv_ctx* ctx, void* stream, void* D, void* iD, const void* L);
/* psolver!(p::Poisson; tol=50eps(T), itmx=6e3): Jacobi-preconditioned conjugate gradients on A x = z                src/flow.jl:300-326
 *   tol < 0 and itmx <= 0 select the reference's defaults.  z is the source on entry and scratch afterwards, as in the reference.
 *   Every scalar of the recurrence and the loop condition stay on the device; the host polls the iteration state one batch of queued
 *   iterations behind.  Returns the iteration count nᵖ and the last r₂ = r·r; synchronises the stream.  -1: NaN residual. */
int ifadv_psolver(ifadv_ctx* ctx, void* stream, void* x, void* eps, void* r, void* z, const void* L, const void* D, const void* iD,
                  unsigned perdir_mask, double tol, int itmx, int* iters, double* r2);
/* myproject!(a,b,w) with inproject!(a,b::Poisson,dt), dt = T(w)·last(a.Δt) formed by the caller                     src/flow.jl:328-347
 *   z ← ∇·u; ϵ, r ← 0; x ← x·dt; psolver!(tol=50eps(T), itmx=2000); u[I,i] -= L[I,i]·∂ᵢx on inside(x); x ← x/dt.
 *   The caller applies BC!(u, ...) afterwards (flow.jl:82,106 -> ifadv_bc_vec). */
int ifadv_myproject(ifadv_ctx* ctx, void* stream, void* u, void* x, void* eps, void* r, void* z, const void* L, const void* D,
                    const void* iD, double dt, unsigned perdir_mask, int* iters, double* r2);

/* ---- WaterLily.MultiLevelPoisson: geometric multigrid, the solver of inproject!'s second method (src/flow.jl:343-347) ---------------
 * and WaterLily's default `psolver`.  WaterLily is not under the reference tree; MultiLevelPoisson(x,L,z;maxlevels,perdir), update!,
 * Vcycle!, smooth! (= pcg!(p;it=6)), residual!, solver! are restated from its published 1.x sources.  Level 1 works on the caller's
 * x ≡ Flow.p, L ≡ Flow.μ₀, z ≡ Flow.σ (borrowed for the handle's lifetime); D, iD, ϵ, r of level 1 and all arrays of the coarser levels
 * (extents 1 + N÷2 while every extent incl. ghosts is even and > 4, at most maxlevels restrictions; <= 0: 10) belong to the handle.
 * One solver cycle (Vcycle!; smooth!; L₂) is replayed as one CUDA graph (IFADV_ML_GRAPH=0: plain launches); the smoother's scalars and
 * early exits stay on the device, the host reads r₂ once per cycle as the reference's loop does.  Single-GPU contexts only (-2 on a
 * z-slab context).  create runs update!(ml). */
typedef struct ifadv_ml ifadv_ml;
int ifadv_ml_create(ifadv_ctx* ctx, ifadv_ml** ml, void* stream, void* x, void* L, void* z, unsigned perdir_mask, int maxlevels);
int ifadv_ml_destroy(ifadv_ml* ml);
int ifadv_ml_levels(const ifadv_ml* ml);
/* device pointer and extents (ghosts included) of a level's array: level 0-based; which: 0 L, 1 D, 2 iD, 3 x, 4 ϵ, 5 r, 6 z */
int ifadv_ml_level_array(ifadv_ml* ml, int level, int which, void** dev_ptr, int64_t Ng[3]);
/* update!(ml): set_diag! on level 1, then restrictL! + BC!(L,0,false,perdir) + set_diag! level by level */
int ifadv_ml_update(ifadv_ml* ml, void* stream);
/* residual!(ml) (level 1: r = z - A x, mean removed), Vcycle!(ml), smooth!(ml.levels[level]) -- the pieces of solver!, for tests */
int ifadv_ml_residual(ifadv_ml* ml, void* stream);
int ifadv_ml_vcycle(ifadv_ml* ml, void* stream);
int ifadv_ml_smooth(ifadv_ml* ml, void* stream, int level);
/* solver!(ml; tol=1e-4, itmx=32): tol < 0 and itmx <= 0 select WaterLily's defaults; returns the cycle count and the last r₂.
 * Synchronises the stream once per cycle.  -1: NaN residual. */
int ifadv_ml_solver(ifadv_ml* ml, void* stream, double tol, int itmx, int* cycles, double* r2);
/* myproject!(a,b::MultiLevelPoisson,w), dt = T(w)·last(a.Δt): z ← ∇·u; x ← x·dt; solver!(b;tol=1e-4,itmx=200); u -= L ∂x; x ← x/dt
 * (src/flow.jl:328-341,343-347).  The caller applies BC!(u,...) afterwards (flow.jl:82,106). */
int ifadv_ml_myproject(ifadv_ml* ml, void* stream, void* u, double dt, int* cycles, double* r2);

/* Stream overlap aid for MPFMomStep! (src/flow.jl:74,89): the midpoint f⁰=(f⁰+f)/2 and the copy f⁰<-f only READ the f that the
 * corrector's advectfq! is about to advance, and that call does not write f before its last directional sweep.  A caller that
 * runs those two field operations on a second stream records an event behind them and passes it here; the NEXT
 * ifadv_advect_vof_rhouu / ifadv_u2rhou_advect_vof_rhouu call on this context then inserts cudaStreamWaitEvent(stream, event)
 * before its first write to f (one-shot).  event: cudaEvent_t. */
int ifadv_defer_f_writes_until(ifadv_ctx* ctx, void* event);

/* ---- z-slab decomposition across the GPUs of one box (one process per GPU) ------------------------------------------------
 * New functionality (the reference is single-device); its oracle is the single-GPU run: owned cells come out
 * BIT-IDENTICAL.  The global grid N1 x N2 x (nz·nranks) is cut into z-slabs; a rank's arrays hold, along z,
 *     [array ghost | G ghost planes (if a lower neighbour exists) | nz owned planes | G ghost planes (upper neighbour) | array ghost]
 * i.e. Ng_local[2] = nz + 2 + (#neighbour sides)·G, with the reference's layout otherwise.  G >= 3 covers the reach of one
 * directional sweep (3 planes below / 2 above).  On a slab context the CMOM entry points update the owned planes only and, after
 * every directional sweep, exchange the G boundary planes of what the sweep produced (f, ρu and once c̄) with both neighbours --
 * ncclSend/ncclRecv in one group on the caller's stream -- so the next sweep reads them like any interior plane; f leaves the
 * call with valid ghost planes.  The caller keeps the ghost planes of u, u⁰/uOld valid (ifadv_exchange_planes after every change)
 * and passes a perdir_mask without the z bit when nranks > 1 (the periodic wrap goes through the exchange).  ifadv_sum_inside and
 * ifadv_mpcfl reduce over the owned planes and all-reduce, so every rank gets the global value.  Communication failures return -4.
 * comm: an ncclComm_t (from NCCL.jl / the host framework; or ifadv_nccl_comm_init below), borrowed for the context's lifetime. */
int ifadv_create_slab(ifadv_ctx** ctx, const int64_t Ng_local[3], int dtype, int device, void* nccl_comm, int rank, int nranks,
                      int ghost_planes, int periodic_z);
/* Transport of the exchanges.  Default: peer-to-peer -- every rank maps a staging buffer and a flag block of its two neighbours
 * through CUDA IPC when the context is created; an exchange is then copy-engine pushes over NVLink (cudaMemcpyAsync into the
 * neighbour's staging buffer) ordered by flags in the receiver's memory, without a host synchronisation and without an SM-resident
 * communication kernel.  The NCCL communicator carries the handles at creation and the scalar all-reduces.  When IPC is not
 * available on some rank (or IFADV_SLAB_P2P=0) all ranks use ncclSend/ncclRecv instead.  1 = peer-to-peer in use. */
int ifadv_slab_p2p(const ifadv_ctx* ctx);
/* owned plane range [kz0, kz1) (1-based), neighbour ranks (-1 = physical boundary), bytes sent by this context's exchanges */
int ifadv_slab_info(const ifadv_ctx* ctx, int* kz0, int* kz1, int* lower, int* upper, int64_t* bytes_sent);
/* exchange the ghost planes of `ncomp` fields (component stride = one scalar field) of `elem_bytes`-byte elements; no-op on a
 * single-GPU context */
int ifadv_exchange_planes(ifadv_ctx* ctx, void* stream, void* field, int ncomp, int elem_bytes);
/* convenience for hosts without an NCCL binding of their own: rank 0 creates the id, every rank initialises with it */
int ifadv_nccl_unique_id(char id[128]);
int ifadv_nccl_comm_init(void** comm, int nranks, const char id[128], int rank, int device);
int ifadv_nccl_comm_destroy(void* comm);

/* NaN detection without a per-call synchronisation: calls made with report == NULL never look at their reductions, so a NaN in f
 * (error("NaN!"), src/advection.jl:148) would go unnoticed.  The NaN count of every call is folded into a sticky device flag when the
 * next call starts; this entry point (and every call WITH a report) reads it: returns -1 if any sweep since the last check produced a
 * NaN, else 0, and clears the flag.  Synchronises the stream -- call it every k-th step. */
int ifadv_check_nan(ifadv_ctx* ctx, void* stream);

/* ---- host-buffer convenience (what bench.py's e2e leg times) --------------------------------------------- */
/* One CMOM advection step of MPFMomStep! on HOST arrays (flow.jl:61,69-70,74,89-92 with the forcing and
 * projection left out: velocities are prescribed): copies f,u to the device, runs
 *   u⁰←u; f⁰←f; u2ρu!,BC!,advectfq!(f⁰;u⁰,u,uOld=u); f⁰=(f⁰+f)/2; f⁰←f; u2ρu!,BC!,advectfq!(f;u,u,uOld=u⁰)
 * and copies f and ρu back.  The context keeps the device work arrays between calls. */
int ifadv_mom_advect_step_host(ifadv_ctx* ctx, void* f_host, const void* u_host, void* rhou_host, double dt, double lambda_rho,
                               int limiter, int normal_scheme, const double uBC[3], unsigned perdir_mask, const int dirO[3],
                               ifadv_report* report);
/* Bytes the last ifadv_mom_advect_step_host call copied host->device / device->host and the number of z-slabs it was
 * pipelined over (1 = single pass).  Large 3-D grids that are not periodic in z are split into z-slabs (an eighth of the planes
 * each, IFADV_HOST_CHUNK=<planes> overrides, 0 disables) extended by 8 overlap planes per interior end, so that the copies of
 * consecutive slabs overlap the kernels; the planes of u a slab shares with its predecessor are copied on the device instead of
 * over PCIe again; the result is bit-identical to the single pass.  Needs page-locked host buffers. */
int ifadv_host_step_bytes(const ifadv_ctx* ctx, int64_t* h2d_bytes, int64_t* d2h_bytes, int* slabs);

#ifdef __cplusplus
}
#endif
#endif /* IFADV_H */
