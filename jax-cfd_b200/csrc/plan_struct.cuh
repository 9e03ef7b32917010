// The plan object behind the opaque cfd_plan handle (shared by plan.cu and multi_gpu.cu).
#pragma once
#include <vector>

#include "common.cuh"

struct cfd_plan {
  int ndim = 0;
  int64_t shape[CFD_MAX_DIM] = {1, 1, 1};
  double step[CFD_MAX_DIM] = {1, 1, 1};
  int batch = 1;
  int device = 0;
  size_t cells = 0;  // per batch member
  // FFT tables
  int lm_row = 0, lm_x = 0, lm_y = 0;  // log2 of: last axis / 2, axis 0, axis 1 (3-D only)
  float2* tw_row = nullptr;
  float2* tw_x = nullptr;
  float2* wbig = nullptr;      // 32768-point x lines: exp(-2 pi i m / 32768)
  float2* xscratch = nullptr;  // ... and the scratch of the split transform
  cfd::SideStreams side;       // ... and the streams that overlap its chunks
  float2* tw_y = nullptr;   // 3-D: complex lines along axis 1
  float2* T2 = nullptr;     // 3-D: second spectrum buffer
  float* nut = nullptr;     // 3-D: Smagorinsky eddy viscosity at cell centres
  float* sfield = nullptr;  // 3-D: six strain-rate fields (strain-field Smagorinsky path)
  float2* rtw = nullptr;
  double* lam[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};
  float* lamf[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};  // float32 copies for the fast path
  int fastd = 0;  // 1: only the mean mode is below the pseudo-inverse cutoff
  double cutoff = 0;
  float norm = 0;
  // workspace
  float* us[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};   // unprojected state (ping)
  float* us2[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};  // unprojected state (pong), lazy chains
  float* rhs = nullptr;
  float* qbuf = nullptr;   // pressure of the latest step
  float* qbuf2 = nullptr;  // pong
  float2* T = nullptr;
  int t_paired = 0;  // 2-D spectrum layout: 1 = pairs of ky lines interleaved (poisson_2d.cu)
  // implementation of the fast-diagonalisation transform (fast_diagonalization.py:101-108):
  // 0 = line FFTs (rfft; every axis a power of two), 1 = matmul along each axis (any shape)
  int impl = 0;
  double* mm_V[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};   // eigenvectors, N_j x N_j row-major, f64
  double* mm_Vt[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};  // ... transposed
  double* mm_diag = nullptr;                                 // pseudo-inverse diagonal in the eigenbasis
  double* mm_w1 = nullptr;                                   // f64 workspaces, batch * cells each
  double* mm_w2 = nullptr;
  // small grids: two chained lazy steps (us -> us2 -> us) captured once as a CUDA graph and
  // replayed, so that the 8 launches per pair cost one graph launch (plan.cu, repeated_lazy)
  cudaGraphExec_t pair_graph = nullptr;
  cfd::StepConsts pair_consts;  // the constants the graph was captured with
  size_t workspace_bytes = 0;
  // host-call staging
  float* dev_a[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};
  float* dev_b[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};
  float* dev_q = nullptr;
  cudaStream_t host_stream = nullptr;
  double* diag_dev = nullptr;
  // slab decomposition (multi_gpu.cu); world == 1 for ordinary plans
  int rank = 0, world = 1;
  int64_t nx_global = 0;    // global rows (shape[0] holds the LOCAL rows of a distributed plan)
  void* shared = nullptr;   // one IPC-exported allocation holding every peer-visible buffer
  size_t shared_bytes = 0;
  void* peer_shared[CFD_MAX_PEERS] = {nullptr};  // peers' `shared` mapped here (own entry = shared)
  unsigned long long epoch = 0;                   // barrier generation
  size_t flags_off = 0;                           // float offset of the flag block inside `shared`
  // staged transpose (multi_gpu.cu): the peers' blocks of this rank's ky lines are copied into
  // local buffers by the copy engines, chunk by chunk (one in / one out stream per peer), while
  // the x-line kernel works on the previous chunk
  float2* xstage = nullptr;
  cudaStream_t st_in[CFD_MAX_PEERS] = {nullptr}, st_out[CFD_MAX_PEERS] = {nullptr};
  cudaEvent_t ev_ready = nullptr, ev_out[CFD_MAX_PEERS] = {nullptr};
  cudaEvent_t ev_in[8][CFD_MAX_PEERS] = {{nullptr}}, ev_comp[8] = {nullptr};
  int dist_state = 0, dist_cur = 0;               // see multi_gpu.cu
  // push mode (multi_gpu.cu): copy kernels on a high-priority stream move the transposes
  int dist_push = 0;
  unsigned long long dist_step = 0;       // steps taken (flag value of the per-block / per-chunk flags)
  unsigned long long dist_nbr_epoch = 0;  // step the neighbour flags must have reached (0: use the barrier)
  cudaStream_t st_comm = nullptr;
  cudaEvent_t ev_blk[8] = {nullptr}, ev_chk[8] = {nullptr}, ev_comm_done = nullptr;
  // per-kernel timing
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<const char*> prof_names;
};


namespace cfd {
void prof_mark(cfd_plan* p, cudaStream_t st, const char* name);
int make_consts(const cfd_plan* p, const cfd_params* prm, StepConsts* c);
// (re)builds the tables that depend on the x extent for a slab plan: x-line twiddles, lambda_x,
// fast-path flag and the 1/(2 Nx Ny) normalisation, from the GLOBAL shape
int plan_tables_create(cfd_plan* p, int ndim, const int64_t* global_shape, const double* step);
// multi_gpu.cu: all-to-all flag barrier of the ranks of a slab plan, enqueued on `st`; size of the
// flag block every slab plan keeps at flags_off
int slab_barrier(cfd_plan* p, cudaStream_t st);
size_t slab_flag_floats();
// multi_gpu_3d.cu: the slab-decomposed 3-D step behind the cfd_dist_* entry points
int dist3_plan_create(cfd_plan** out, const int64_t* global_shape, const double* step, int rank, int world,
                      int device);
int dist3_load(cfd_plan* p, cudaStream_t st, const float* const* v_local);
int dist3_advance(cfd_plan* p, cudaStream_t st, int nsteps, const StepConsts& c);
int dist3_store(cfd_plan* p, cudaStream_t st, float* const* v_local_out, float* q_local_out);
}  // namespace cfd
