// ifadv_ctx.hpp -- context object and launch descriptors shared by the translation units of libifadv_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "ifadv_sweep.cuh"

// ------------------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------------------
// z-slab decomposition (ifadv_create_slab): the local arrays hold the owned planes plus G ghost planes per neighbour side
struct ifadv_slab {
  void* comm;           // ncclComm_t (borrowed)
  int rank, nranks;     // nranks <= 1: not a slab
  int G, glo, ghi;      // ghost planes per side with a neighbour; present below / above
  int lower, upper;     // neighbour ranks (-1: physical boundary)
  long long bytes_sent; // by the exchanges of this context
  int overlap;          // 1: boundary planes first, exchange on slab_stream underneath the interior planes
};
// Peer-to-peer exchange path of a z-slab context: neighbours PUSH their boundary planes with the copy engines (cudaMemcpyAsync over
// NVLink into a staging buffer of the receiver, mapped through CUDA IPC) and signal with flags in the receiver's memory.
struct ifadv_p2p {
  int on;                      // 1: in use (every rank of the communicator agreed at creation)
  char* stage[2];              // mine: [0] filled by the lower neighbour, [1] by the upper neighbour
  unsigned* flags;             // mine: [0] READY_LO [1] READY_UP (neighbour may be pushed to), [2] DONE_LO [3] DONE_UP (its push has landed); [8] timeout
  char* peer_stage[2];         // [0]: the lower neighbour's stage[1] (I am its upper neighbour), [1]: the upper neighbour's stage[0]
  unsigned* peer_flags[2];     // [0]: the lower neighbour's flag block, [1]: the upper neighbour's
  void* opened[6];             // IPC mappings to close
  size_t cap;                  // bytes per staging buffer
  unsigned seq;                // exchange counter (identical on all ranks: SPMD)
};
struct ifadv_ctx {
  int D, dtype, device;
  ifadv_p2p p2p;
  int kz0, kz1;  // planes [kz0, kz1) of dimension 3 the sweeps update: 2 .. n[2], or the owned planes of a z-slab
  ifadv_slab slab;
  cudaStream_t slab_stream;  // second stream of the overlapped exchange
  cudaEvent_t slab_ev[2];    // [0] boundary planes swept (main -> slab_stream), [1] exchange done (slab_stream -> main)
  ifadv::Geo g;
  int64_t Ng[3];
  unsigned long long* red_dev;   // 3 sweeps x 8 slots
  unsigned long long* red_host;  // pinned mirror
  unsigned long long* misc_dev;  // 8 slots for CFL / sum reductions
  unsigned long long* misc_host;
  int64_t launches;
  std::string err;
  // work arrays of the host-buffer convenience entry point (allocated lazily)
  void* w[16];
  void* pin_f;
  void* pin_u;
  void* pin_ru;
  cudaStream_t own_stream;
  int* st_list;        // interface-cell list of the surface-tension kernels (ifadv_forcing.cuh), allocated lazily
  unsigned* st_cnt;    // its device-side counter
  unsigned st_cap;
  cudaEvent_t wait_f;  // one-shot: the next CMOM advect call waits for it before its first write to f (ifadv_defer_f_writes_until)
  void* pipe;  // z-slab pipeline of the host-buffer entry point (HostPipe, ifadv_b200.cu), built lazily
  int64_t host_h2d, host_d2h;  // bytes the last ifadv_mom_advect_step_host call copied in / out
  int host_slabs;              // z-slabs it was pipelined over (1 = single pass)
  // optional per-launch CUDA-event timing of the fused sweep (ifadv_profile)
  int use_march;  // 1: register-marching (y,z) + plane-marching (x) kernels (default); 2: plane-marching only; 0: v1 tile kernel
  int use_along2;  // 1: lean register-marching kernel ifadv_along2.cuh for y/z sweeps (always; the first generation is retired)
  int use_vofcell; // 1 (default): cell-parallel kernel ifadv_vofcell.cuh for 3-D pure-VOF sweeps; 0 (IFADV_VOF_KERNEL=lean or any IFADV_KERNEL): along2 / xrow
  int use_xrow;    // 1 (default): warp-autonomous row kernel ifadv_xrow.cuh for CMOM x sweeps; 0: ifadv_xsweep.cuh
  int prof_on, prof_n;
  cudaEvent_t* prof_ev;  // 2 * IFADV_PROF_MAX events
  unsigned char* prof_tag;  // per launch: bit 0 = fused first sweep (10s+1 B/cell) / standard sweep (13s+1 B/cell); bits 1.. = 2*j + fused
  double prof_dir_ms[8];    // accumulated by ifadv_profile_read: index 2*j + fused
  int64_t prof_dir_n[8];
  // pressure solver (ifadv_poisson.cu), allocated lazily: control block in device memory, pinned mirror of two polls, their events
  void* pois_ctl;
  void* pois_host;
  cudaEvent_t pois_ev[2];
  int pois_slab_flag;  // what the control block's `slab` field holds
};
void ifadv_poisson_free(ifadv_ctx* c);  // ifadv_poisson.cu
#define IFADV_PROF_MAX 4096

#define CU_CHECK(ctx, call)                                                                 \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      if (ctx) (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);             \
      return -3;                                                                            \
    }                                                                                       \
  } while (0)


namespace ifadv {
template <class T> struct SweepCfg {
  const T *f_in, *u, *u0, *rhou_in, *uOld, *drho, *uexit;
  T *f_out, *rhou_out, *rhouf;
  int8_t* cbar;
  double dt, lr;
  double A[3];
  int scheme, lim, first, j;  // j 0-based
  int fused;                  // see SweepP::fused
  unsigned long long* red;
};


// the whole 2-D pure-VOF step in one cooperative launch (small grids are launch-latency bound); defined in the (T, 2, 0, 0) unit
template <class T> int launch_vof2d_step(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& qa, const SweepCfg<T>& qb, T* f_final, T* rhouf);
// one fused directional sweep; defined in ifadv_sweep_inst.cu, one translation unit per (T, D, MOM)
template <class T, int D, bool MOM> int launch_sweep_dim(ifadv_ctx* c, cudaStream_t st, const SweepCfg<T>& q);
}  // namespace ifadv
