// Slab-decomposed multi-GPU time step: one process per GPU, the grid is
// split along axis 0, and ALL inter-GPU data movement is done by the compute kernels themselves
// through peer-mapped (CUDA IPC) memory over NVLink:
//   * stencil halo: explicit2d_kernel reads rows -3..-1 / Nloc..Nloc+3 of (u*, v*, q) directly
//     from the neighbouring ranks' buffers (SlabSrc::prev / next);
//   * distributed FFT: every rank row-transforms its slab into its local T[ky][x_loc]; the
//     x-direction kernel of rank r then assembles each of ITS ky lines from all ranks' T, applies
//     fwd * D * inv and writes the line back in place -- the all-to-all transpose and its inverse
//     are the loads and stores of xlines_kernel (LinePeers), large contiguous peer accesses;
//   * ordering: a one-CTA flag barrier kernel (system-scope stores/loads on peer-mapped flags)
//     between the phases.  No NCCL on the data path; the host side only exchanges 64-byte IPC
//     handles once (any transport: torch.distributed in jax_cfd_b200.distributed).
//
// Default since round 2 ("push" mode, CFD_DIST_MODE=pull selects the scheme above):
//   * the stencil (two launches) and the row FFT run block by block over the slab's rows; as soon
//     as a block of rows is transformed, a TMA bulk-copy kernel on a second, high-priority stream
//     (slab_push_tma_kernel: cp.async.bulk global -> shared -> peer global, SASS UBLKCP) stores its
//     part of every ky line into the line owners' receive buffers, overlapping the stencil / row
//     FFT of the next block; a per-(rank, block) flag follows every copy;
//   * the x-direction kernel then READS only local memory (its own spectrum + the receive buffers
//     through the peer table) and WRITES its results straight into the slab owners' spectra --
//     posted NVLink stores, so the backward transpose needs neither a copy nor a buffer; one flag
//     per rank says "my lines are back";
//   * flags replace two of the three global barriers: a rank waits only for the data it is about
//     to read; the barrier before the stencil becomes a neighbour-only flag.
// Results stay bit-identical to the single-GPU path (same kernels, same arithmetic).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "plan_struct.cuh"

namespace cfd {

int launch_explicit_2d_slab(cudaStream_t stream, SlabSrc su, SlabSrc sv, SlabSrc sq, float* us,
                            float* vs, float* rhs, int batch, int Nx, int Ny, int row0,
                            int nx_global, const StepConsts& c, int dvdt_mode, int tile_begin,
                            int tile_count);
int explicit_2d_tile_rows(int batch, int Nx, int Ny);
int launch_rfft_rows_block(cudaStream_t st, int lm_row, const float* rhs, float2* T, int batch, int Nx,
                           const float2* tw, const float2* rtw, int paired, int x_begin, int x_count);
int launch_rfft_rows(cudaStream_t, int lm_row, const float* rhs, float2* T, int batch, int Nx,
                     const float2* tw, const float2* rtw, int paired);
int launch_irfft_rows(cudaStream_t, int lm_row, const float2* T, float* q, int batch, int Nx,
                      const float2* tw, const float2* rtw, int paired);
int launch_xlines_peers(cudaStream_t st, int lm_x, const LinePeers& peers, int lnloc,
                        size_t line_begin, size_t nlines, int My, const float2* tw,
                        const double* lamx, const double* lamy, const float* lamxf,
                        const float* lamyf, int fastd, double cutoff, float norm, float2* scratch,
                        const float2* wbig, const SideStreams* side, int paired, const float* dtab,
                        const LinePeers* peers_out);
int launch_correct_2d(cudaStream_t, const float* us, const float* vs, const float* q,
                      const float* qnext, float* uo, float* vo, int batch, int Nx, int Ny,
                      float inv_hx, float inv_hy);

namespace {

struct FlagPeers {
  unsigned long long* p[CFD_MAX_PEERS];
};

// All-to-all flag barrier: rank r publishes `epoch` in slot r of every peer's flag array, then
// waits until every peer has published it in r's own array.  The kernels before it on the stream
// have completed (their stores are performed), so observing a peer's flag implies its data is
// visible.  A cycle budget bounds the spin so that a lost peer cannot hang the GPU.
// Spin budget of the device-side waits, in clock cycles (CFD_DIST_TIMEOUT_S seconds, default 120):
// calls are enqueue-only, so ordinary host skew between ranks (first-call set-up, a pause in the
// host program) must fit inside it.  After a time-out the error word is set and every later wait
// returns at once, so a lost peer cannot hang the GPU; cfd_dist_check reports it.
long long spin_budget() {
  static const long long v = [] {
    const char* e = getenv("CFD_DIST_TIMEOUT_S");
    const double s = e ? atof(e) : 120.0;
    return (long long)((s > 0 ? s : 120.0) * 1.9e9);
  }();
  return v;
}

__device__ __forceinline__ void spin_until(volatile unsigned long long* flag, unsigned long long value,
                                           volatile unsigned long long* err, long long budget) {
  if (*err) return;  // a previous wait already timed out: do not stall the chain again
  const long long t0 = clock64();
  while (*flag < value) {
    if (clock64() - t0 > budget) {
      *err = value;
      break;
    }
  }
}

__global__ void slab_barrier_kernel(FlagPeers fp, int rank, int world, unsigned long long epoch,
                                    unsigned long long* err, long long budget) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();
  volatile unsigned long long* remote = fp.p[p] + rank;
  *remote = epoch;
  __threadfence_system();
  spin_until(fp.p[rank] + p, epoch, err, budget);
  __threadfence_system();
}

// flag slots (8 bytes each) inside every rank's flag block
constexpr int kSlotErr = 16;    // 0..15: the all-to-all barrier above
constexpr int kSlotFwd = 32;    // + src * 8 + block   : row block `block` of rank `src` has arrived
constexpr int kSlotBack = 96;   // + src * 8 + chunk   : line chunk `chunk` of rank `src` has arrived
constexpr int kSlotNbr = 160;   // + 0 / 1             : previous / next rank has finished the step
constexpr int kFlagSlots = 192;
constexpr int kMaxBlocks = 8;

// thread p < ntargets: publish `value` in `slot` of target p's flag block.  The copy kernel before
// it on the stream has completed, so observing the flag implies its data is visible.
__global__ void slab_signal_kernel(FlagPeers fp, int ntargets, int slot, unsigned long long value) {
  const int p = threadIdx.x;
  if (p >= ntargets) return;
  __threadfence_system();
  volatile unsigned long long* f = fp.p[p] + slot;
  *f = value;
}

// thread i < nsrc * nper: wait until flags[base + (i / nper) * 8 + i % nper] >= value
__global__ void slab_wait_kernel(unsigned long long* flags, int base, int nsrc, int nper,
                                 unsigned long long value, long long budget) {
  const int i = threadIdx.x;
  if (i < nsrc * nper) spin_until(flags + base + (i / nper) * kMaxBlocks + i % nper, value, flags + kSlotErr, budget);
  __threadfence_system();
}

// layout of the shared (IPC-exported) allocation, identical on every rank; units = floats
struct SharedLayout {
  size_t field;  // floats per local field
  size_t off_vin[2], off_us[2][2], off_q[2], off_T, off_L[2], off_flags, total_bytes;
};
SharedLayout shared_layout(size_t nloc, size_t ny) {
  SharedLayout L;
  L.field = nloc * ny;
  size_t o = 0;
  for (int a = 0; a < 2; ++a) { L.off_vin[a] = o; o += L.field; }
  for (int s = 0; s < 2; ++s)
    for (int a = 0; a < 2; ++a) { L.off_us[s][a] = o; o += L.field; }
  for (int s = 0; s < 2; ++s) { L.off_q[s] = o; o += L.field; }
  L.off_T = o; o += L.field;  // My * Nloc float2 = Nloc * Ny floats
  // receive buffers of the push mode: one block of (My / world) lines x Nloc points per source
  // rank (the own slot stays unused) = one field
  L.off_L[0] = o; o += L.field;
  L.off_L[1] = L.off_L[0];
  L.off_flags = o; o += 2 * kFlagSlots;  // kFlagSlots 8-byte flags
  L.total_bytes = o * sizeof(float);
  return L;
}

float* fptr(void* base, size_t off) { return reinterpret_cast<float*>(base) + off; }

int barrier(cfd_plan* p, cudaStream_t st) {
  if (p->world == 1) return 0;
  FlagPeers fp;
  for (int r = 0; r < CFD_MAX_PEERS; ++r)
    fp.p[r] = reinterpret_cast<unsigned long long*>(fptr(p->peer_shared[r < p->world ? r : p->rank], p->flags_off));
  p->epoch += 1;
  unsigned long long* err = reinterpret_cast<unsigned long long*>(fptr(p->shared, p->flags_off)) + kSlotErr;
  slab_barrier_kernel<<<1, 32, 0, st>>>(fp, p->rank, p->world, p->epoch, err, spin_budget());
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  prof_mark(p, st, "barrier");
  return 0;
}

// CFD_DIST_STAGED=1 selects the copy-engine transpose below instead of the in-kernel one (x-line
// kernel loading / storing peer memory), which is the default because it measured faster.
int staged_mode() {
  static const int v = [] {
    const char* e = getenv("CFD_DIST_STAGED");
    return e ? atoi(e) : 0;
  }();
  return v;
}

// CFD_DIST_MODE=pull selects the round-1 scheme (x-line kernel loads / stores peer memory, three
// global barriers per step); default: push (see the header of this file)
bool push_mode() {
  static const bool v = [] {
    const char* e = getenv("CFD_DIST_MODE");
    return !(e && strcmp(e, "pull") == 0);
  }();
  return v && !staged_mode();
}

// The x pass of the distributed FFT with the transposes on the COPY ENGINES.
//
// Rank r owns lines [r L, (r+1) L), L = My / world.  Inside every rank's T[ky][x_loc] those lines
// are ONE contiguous block of L * Nloc points, so gathering a chunk of lines is one plain copy per
// peer into a local buffer of the same [line][x_loc] layout, and the x-line kernel runs unchanged
// with its peer table pointing at the local copies (its own block is read in place).  Chunks are
// pipelined: copy-in (one stream per peer) | x-line kernel (st) | copy-out (one stream per peer),
// so NVLink carries chunk c+1 in and chunk c-1 out while the SMs transform chunk c.
//
// Measured at 2 x 8192^2 (round 1), x pass per step: in-kernel transpose 0.47 ms; this path 0.59 ms
// (4 chunks).  Two reasons, both visible in the per-chunk event times: copy-engine copies from /
// to CUDA-IPC mapped peer memory ran at ~260 GB/s (the same copies between cudaMalloc'ed buffers
// of one process reach 700 GB/s, scripts/ubench/p2p_ce.cu), and the x-line kernel is 30 % slower
// through its peer table (0.37 ms) than on one contiguous line buffer (0.28 ms).  Kept as an
// opt-in experiment (results are bit-identical); the in-kernel path stays the default.
int xpass_staged(cfd_plan* p, cudaStream_t st, const SharedLayout& L, int lnloc) {
  const size_t nloc = (size_t)p->shape[0];
  const int My = (int)p->shape[1] / 2;
  const size_t lines = (size_t)My / p->world;
  const size_t gl0 = (size_t)p->rank * lines;
  int K = p->world >= 4 ? 4 : 4;
  if (const char* e = getenv("CFD_DIST_CHUNKS")) K = atoi(e);
  if (K > 8) K = 8;
  while (K > 1 && (lines % K || lines / K < 64)) K /= 2;
  const size_t chunk = lines / K;
  auto peerT = [&](int r) { return reinterpret_cast<float2*>(fptr(p->peer_shared[r], L.off_T)); };
  auto stage = [&](int r) { return p->xstage + (size_t)r * lines * nloc; };
  // peer table of the kernel: local copies, addressed with the GLOBAL line number like T itself
  LinePeers tab;
  for (int r = 0; r < CFD_MAX_PEERS; ++r) {
    const int q = r < p->world ? r : p->rank;
    tab.p[r] = (q == p->rank) ? peerT(q) : stage(q) - gl0 * nloc;
  }
  CFD_CUDA_OK(cudaEventRecord(p->ev_ready, st));
  for (int r = 0; r < p->world; ++r)
    if (r != p->rank) CFD_CUDA_OK(cudaStreamWaitEvent(p->st_in[r], p->ev_ready, 0));
  for (int c = 0; c < K; ++c) {
    const size_t lb = (size_t)c * chunk;
    const size_t bytes = chunk * nloc * sizeof(float2);
    for (int k = 1; k < p->world; ++k) {
      const int r = (p->rank + k) % p->world;  // staggered start so the peers are not hit in step
      CFD_CUDA_OK(cudaMemcpyAsync(stage(r) + lb * nloc, peerT(r) + (gl0 + lb) * nloc, bytes,
                                  cudaMemcpyDeviceToDevice, p->st_in[r]));
      CFD_CUDA_OK(cudaEventRecord(p->ev_in[c][r], p->st_in[r]));
      CFD_CUDA_OK(cudaStreamWaitEvent(st, p->ev_in[c][r], 0));
    }
    prof_mark(p, st, "xwait_in");
    if (int e = launch_xlines_peers(st, p->lm_x, tab, lnloc, gl0 + lb, chunk, My, p->tw_x, p->lam[0],
                                    p->lam[1], p->lamf[0], p->lamf[1], p->fastd, p->cutoff, p->norm,
                                    p->xscratch, p->wbig, nullptr, p->t_paired, nullptr, nullptr))
      return e;
    CFD_CUDA_OK(cudaEventRecord(p->ev_comp[c], st));
    prof_mark(p, st, "xchunk");
    for (int k = 1; k < p->world; ++k) {
      const int r = (p->rank + k) % p->world;
      CFD_CUDA_OK(cudaStreamWaitEvent(p->st_out[r], p->ev_comp[c], 0));
      CFD_CUDA_OK(cudaMemcpyAsync(peerT(r) + (gl0 + lb) * nloc, stage(r) + lb * nloc, bytes,
                                  cudaMemcpyDeviceToDevice, p->st_out[r]));
    }
  }
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    CFD_CUDA_OK(cudaEventRecord(p->ev_out[r], p->st_out[r]));
    CFD_CUDA_OK(cudaStreamWaitEvent(st, p->ev_out[r], 0));
  }
  return 0;
}

// ---- the same block copies with the TMA (bulk asynchronous copies, SASS UBLKCP) -----------------------
// One thread per CTA drives a ring of shared-memory stages: cp.async.bulk global -> shared (completion
// on an mbarrier), then cp.async.bulk shared -> global (local HBM or the peer's memory over NVLink).
// Megabytes stay in flight per CTA-sized footprint of 32 threads, so a few dozen CTAs saturate
// NVLink while the SMs keep computing -- LSU-driven copy kernels (the first push implementation)
// need the whole GPU for that (scripts/ubench/p2p.cu: 650 GB/s with >= 148 CTAs, 200 GB/s with 48).
// A "row" is a line (plain layout) or a pair of lines (pair-interleaved layout): both buffers use
// the same layout in push mode, so every piece is one contiguous run of bytes.
struct TmaPush {
  const char* src;          // + r * src_rank_stride + row * src_row_stride
  char* dst[CFD_MAX_PEERS]; // + row * dst_row_stride
  size_t src_rank_stride, src_row_stride, dst_row_stride;
  int nrows;
  unsigned row_bytes, chunk_bytes;  // chunk_bytes divides row_bytes, multiple of 16
};
constexpr int kTmaStages = 6;
constexpr unsigned kTmaChunk = 16384;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32) slab_push_tma_kernel(TmaPush a) {
  extern __shared__ __align__(128) unsigned char stage_mem[];
  __shared__ __align__(8) unsigned long long full[kTmaStages];
  if (threadIdx.x != 0) return;
  const int r = blockIdx.y;
  const char* src = a.src + (size_t)r * a.src_rank_stride;
  char* dst = a.dst[r];
  if (dst == nullptr) return;  // nothing to send to this rank (itself)
  const unsigned cpr = a.row_bytes / a.chunk_bytes;
  const long long nitems = (long long)a.nrows * cpr;
  const long long first = blockIdx.x, stride = gridDim.x;
  const long long mine = first < nitems ? (nitems - first + stride - 1) / stride : 0;
  for (int s = 0; s < kTmaStages; ++s)
    asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(smem_u32(&full[s])) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  auto src_of = [&](long long it) {
    const long long item = first + it * stride;
    return src + (size_t)(item / cpr) * a.src_row_stride + (size_t)(item % cpr) * a.chunk_bytes;
  };
  auto dst_of = [&](long long it) {
    const long long item = first + it * stride;
    return dst + (size_t)(item / cpr) * a.dst_row_stride + (size_t)(item % cpr) * a.chunk_bytes;
  };
  auto load = [&](long long it) {
    const int s = (int)(it % kTmaStages);
    const unsigned bar = smem_u32(&full[s]);
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(a.chunk_bytes)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(stage_mem + (size_t)s * kTmaChunk)),
                 "l"(src_of(it)), "r"(a.chunk_bytes), "r"(bar)
                 : "memory");
  };
  // prologue: fill the ring
  for (long long it = 0; it < mine && it < kTmaStages; ++it) load(it);
  for (long long it = 0; it < mine; ++it) {
    const int s = (int)(it % kTmaStages);
    const unsigned bar = smem_u32(&full[s]), parity = (unsigned)((it / kTmaStages) & 1);
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.b32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
    }
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_of(it)),
                 "r"(smem_u32(stage_mem + (size_t)s * kTmaChunk)), "r"(a.chunk_bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    // the stage of the PREVIOUS item is free once its store has read shared memory: refill it
    if (it >= 1 && it - 1 + kTmaStages < mine) {
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      load(it - 1 + kTmaStages);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // writes performed before the kernel ends
}

int launch_push_tma(cudaStream_t st, const TmaPush& a, int ctas, int world) {
  constexpr int smem = kTmaStages * (int)kTmaChunk;
  if (int e = opt_in_smem(slab_push_tma_kernel, smem)) return e;
  slab_push_tma_kernel<<<dim3(ctas, world), 32, smem, st>>>(a);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

unsigned pick_chunk(unsigned row_bytes) {
  unsigned c = row_bytes < kTmaChunk ? row_bytes : kTmaChunk;
  while (row_bytes % c) c -= 16;
  return c;
}

int signal_flags(cfd_plan* p, cudaStream_t st, const SharedLayout& L, int ntargets, const int* targets,
                 int slot, unsigned long long value) {
  FlagPeers fp;
  for (int i = 0; i < CFD_MAX_PEERS; ++i)
    fp.p[i] = reinterpret_cast<unsigned long long*>(
        fptr(p->peer_shared[i < ntargets ? targets[i] : p->rank], L.off_flags));
  slab_signal_kernel<<<1, 32, 0, st>>>(fp, ntargets, slot, value);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int wait_flags(cfd_plan* p, cudaStream_t st, const SharedLayout& L, int base, int nsrc, int nper,
               unsigned long long value, const char* name) {
  slab_wait_kernel<<<1, 64, 0, st>>>(reinterpret_cast<unsigned long long*>(fptr(p->shared, L.off_flags)), base,
                                     nsrc, nper, value, spin_budget());
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  prof_mark(p, st, name);
  return 0;
}

// how many row blocks / line chunks a step is pipelined in (CFD_DIST_BLOCKS / CFD_DIST_CHUNKS
// override): the largest count <= the wish whose pieces are whole stencil tiles / kernel line groups
int pick_pieces(int total, int unit, int wish) {
  int n = wish < 1 ? 1 : (wish > kMaxBlocks ? kMaxBlocks : wish);
  while (n > 1 && (total % n || (total / n) % unit)) --n;
  return n;
}

// One step in push mode (see the header).  cur / nxt: ping-pong slots of (u*, v*, q).
//
// Forward transpose: every finished row block is stored by the TMA copy kernel (comm stream) into
// the line owners' receive buffers R_s[rank] (layout of the slab spectrum, lines of s only), overlapping
// the stencil / row FFT of the next block.  The x-line kernel then READS only local memory (its own
// spectrum + the receive buffers, through the peer table) and WRITES its results straight into the
// slab owners' spectra -- posted NVLink stores that drain while other CTAs compute -- so the
// backward transpose needs neither a copy nor a buffer; one flag per rank says "my lines are back".
int step_push(cfd_plan* p, cudaStream_t st, const SharedLayout& L, const StepConsts& c) {
  const int W = p->world, rank = p->rank;
  const int nloc = (int)p->shape[0], Ny = (int)p->shape[1], My = Ny / 2;
  const int lines = My / W;
  const int prev = (rank + W - 1) % W, next = (rank + 1) % W;
  const int cur = p->dist_cur, nxt = cur ^ 1;
  auto src3 = [&](size_t off) {
    return SlabSrc{fptr(p->peer_shared[prev], off), fptr(p->shared, off), fptr(p->peer_shared[next], off)};
  };
  const unsigned long long step_id = ++p->dist_step;
  float2* Tloc = reinterpret_cast<float2*>(fptr(p->shared, L.off_T));
  static const int env_blocks = [] { const char* e = getenv("CFD_DIST_BLOCKS"); return e ? atoi(e) : 0; }();
  const int wish_blocks = env_blocks > 0 ? env_blocks : (W <= 2 ? 2 : 4);
  const int TX = explicit_2d_tile_rows(1, nloc, Ny);
  const int NB = pick_pieces(nloc, TX < 32 ? 32 : TX, wish_blocks);  // whole tiles and row-kernel CTAs
  const int bx = nloc / NB;
  const int paired = p->t_paired;
  const size_t esz = paired ? 16 : 8;                 // bytes per point of a row (line or line pair)
  const int nrows = paired ? lines / 2 : lines;       // rows per rank
  const size_t blk = (size_t)lines * nloc;            // float2 per (owner, source) receive buffer
  int all[CFD_MAX_PEERS];
  for (int r = 0; r < W; ++r) all[r] = r;
  static const int wish_ctas = [] { const char* e = getenv("CFD_DIST_COPY_CTAS"); return e ? atoi(e) : 0; }();
  // CTAs of the copy kernel per destination rank: enough bytes in flight to fill NVLink, few enough
  // to leave the SMs to the row FFT.  Measured: 2 GPUs (one destination) 8 / 32 CTAs = 1.343 / 1.264
  // ms per step; 4 GPUs 6 / 12 / 24 = 1.495 / 1.537 / 1.539 ms.
  const int tma_ctas = wish_ctas > 0 ? wish_ctas : (W <= 2 ? 32 : (W <= 4 ? 6 : 4));
  auto recv = [&](int owner, int source) {  // R_owner[source]
    return reinterpret_cast<float2*>(fptr(p->peer_shared[owner], L.off_L[0])) + (size_t)source * blk;
  };

  prof_mark(p, st, "begin");
  // ---- neighbours' (u*, v*, q) of the previous step are complete (first step after a load: everyone's input)
  if (p->dist_state == 1 || p->dist_nbr_epoch == 0) {
    if (int e = barrier(p, st)) return e;
  } else {
    if (int e = wait_flags(p, st, L, kSlotNbr, 1, 2, p->dist_nbr_epoch, "wait_nbr")) return e;
  }
  // ---- stencil + row FFT block by block; each finished block is pushed to the line owners.
  // The stencil runs in TWO launches (CFD_DIST_STENCIL_PARTS overrides): a single launch delays the
  // first push by the whole stencil, one launch per row block costs a ramp and a tail each.  Measured
  // on 4 x 8192^2 (32768 x 8192, four row blocks): 1 / 2 / 4 launches = 1.570 / 1.488 / 1.540 ms.
  static const int wish_parts = [] { const char* e = getenv("CFD_DIST_STENCIL_PARTS"); return e ? atoi(e) : 0; }();
  int nst = wish_parts > 0 ? (wish_parts > NB ? NB : wish_parts) : (NB >= 2 ? 2 : 1);
  while (NB % nst) --nst;
  const int bpst = NB / nst;  // row blocks per stencil launch
  for (int b = 0; b < NB; ++b) {
    const SlabSrc none = {nullptr, nullptr, nullptr};
    int e = 0;
    if (b % bpst == 0) {
      const int t0 = b * (bx / TX), tn = bpst * (bx / TX);
      if (p->dist_state == 1)
        e = launch_explicit_2d_slab(st, src3(L.off_vin[0]), src3(L.off_vin[1]), none,
                                    fptr(p->shared, L.off_us[nxt][0]), fptr(p->shared, L.off_us[nxt][1]), p->rhs, 1,
                                    nloc, Ny, rank * nloc, (int)p->nx_global, c, 0, t0, tn);
      else
        e = launch_explicit_2d_slab(st, src3(L.off_us[cur][0]), src3(L.off_us[cur][1]), src3(L.off_q[cur]),
                                    fptr(p->shared, L.off_us[nxt][0]), fptr(p->shared, L.off_us[nxt][1]), p->rhs, 1,
                                    nloc, Ny, rank * nloc, (int)p->nx_global, c, 0, t0, tn);
      if (e) return e;
    }
    prof_mark(p, st, "explicit_2d_slab");
    if (int e2 = launch_rfft_rows_block(st, p->lm_row, p->rhs, Tloc, 1, nloc, p->tw_row, p->rtw, paired, b * bx, bx))
      return e2;
    prof_mark(p, st, "rfft_rows");
    CFD_CUDA_OK(cudaEventRecord(p->ev_blk[b], st));
    CFD_CUDA_OK(cudaStreamWaitEvent(p->st_comm, p->ev_blk[b], 0));
    TmaPush a;
    a.src = reinterpret_cast<const char*>(Tloc) + (size_t)(b * bx) * esz;
    a.nrows = nrows;
    a.src_rank_stride = (size_t)nrows * nloc * esz;
    a.src_row_stride = (size_t)nloc * esz;
    a.dst_row_stride = (size_t)nloc * esz;
    a.row_bytes = (unsigned)(bx * esz);
    a.chunk_bytes = pick_chunk(a.row_bytes);
    for (int r = 0; r < CFD_MAX_PEERS; ++r)
      a.dst[r] = (r < W && r != rank) ? reinterpret_cast<char*>(recv(r, rank)) + (size_t)(b * bx) * esz : nullptr;
    if (int e4 = launch_push_tma(p->st_comm, a, tma_ctas, W)) return e4;
    if (int e3 = signal_flags(p, p->st_comm, L, W, all, kSlotFwd + rank * kMaxBlocks + b, step_id)) return e3;
  }
  // ---- every rank's blocks of MY lines have arrived
  if (int e = wait_flags(p, st, L, kSlotFwd, W, NB, step_id, "wait_fwd")) return e;
  // ---- x lines: read locally (own spectrum + receive buffers), write to the slab owners
  LinePeers rd, wr;
  for (int r = 0; r < CFD_MAX_PEERS; ++r) {
    const int q = r < W ? r : rank;
    // the kernel addresses lines by their GLOBAL number: a receive buffer holds lines rank * L ..
    rd.p[r] = (q == rank) ? Tloc : recv(rank, q) - (size_t)rank * blk;
    wr.p[r] = reinterpret_cast<float2*>(fptr(p->peer_shared[q], L.off_T));
  }
  int lnloc = 0;
  while ((1 << lnloc) < nloc) ++lnloc;
  if (int e = launch_xlines_peers(st, p->lm_x, rd, lnloc, (size_t)rank * lines, lines, My, p->tw_x, p->lam[0],
                                  p->lam[1], p->lamf[0], p->lamf[1], p->fastd, p->cutoff, p->norm, p->xscratch,
                                  p->wbig, &p->side, paired, nullptr, &wr))
    return e;
  prof_mark(p, st, "xlines");
  if (int e = signal_flags(p, st, L, W, all, kSlotBack + rank * kMaxBlocks, step_id)) return e;
  // ---- all ky lines of MY rows are back
  if (int e = wait_flags(p, st, L, kSlotBack, W, 1, step_id, "wait_back")) return e;
  if (int e = launch_irfft_rows(st, p->lm_row, Tloc, fptr(p->shared, L.off_q[nxt]), 1, nloc, p->tw_row, p->rtw,
                                paired))
    return e;
  prof_mark(p, st, "irfft_rows");
  // ---- tell the neighbours that this rank's (u*, v*, q) of this step are complete
  {
    const int nb[2] = {next, prev};  // I am `prev` of my next rank (its slot 0) and `next` of my previous (slot 1)
    if (int e = signal_flags(p, st, L, 1, &nb[0], kSlotNbr + 0, step_id)) return e;
    if (int e = signal_flags(p, st, L, 1, &nb[1], kSlotNbr + 1, step_id)) return e;
  }
  p->dist_nbr_epoch = step_id;
  p->dist_cur = nxt;
  p->dist_state = 2;
  return 0;
}

}  // namespace

int slab_barrier(cfd_plan* p, cudaStream_t st) { return barrier(p, st); }
size_t slab_flag_floats() { return 2 * (size_t)kFlagSlots; }

}  // namespace cfd

using namespace cfd;

extern "C" {

int cfd_dist_plan_create_nd(cfd_plan** out, int ndim, const int64_t* global_shape, const double* step, int rank,
                            int world, int device) {
  if (ndim == 3) return dist3_plan_create(out, global_shape, step, rank, world, device);
  if (ndim != 2) return set_error_msg("cfd_dist_plan_create_nd: ndim must be 2 or 3");
  return cfd_dist_plan_create(out, global_shape, step, rank, world, device);
}

int cfd_dist_plan_create(cfd_plan** out, const int64_t* global_shape, const double* step, int rank,
                         int world, int device) {
  if (!out || !global_shape || !step) return set_error_msg("null argument");
  *out = nullptr;
  if (world < 1 || world > CFD_MAX_PEERS || (world & (world - 1)))
    return set_error_msg("world size must be 1, 2, 4 or 8");
  if (rank < 0 || rank >= world) return set_error_msg("bad rank");
  const int64_t Nxg = global_shape[0], Ny = global_shape[1];
  if (Nxg % world) return set_error_msg("axis 0 must be divisible by the number of ranks");
  const int64_t nloc = Nxg / world;
  if (nloc < 16 || (nloc & (nloc - 1))) return set_error_msg("local slab must be a power of two >= 16 rows");
  if ((Ny / 2) % world || (Ny / 2 / world) % 16) return set_error_msg("Ny/2 lines must split evenly over the ranks");
  if (Nxg > (1 << 15)) return set_error_msg("global axis 0 longer than 32768 is not supported");
  // an ordinary plan for the LOCAL slab gives the row tables + local workspace ...
  int64_t local_shape[2] = {nloc, Ny};
  cfd_plan* p = nullptr;
  if (int e = cfd_plan_create(&p, 2, local_shape, step, 1, device)) return e;
  // ... then replace what depends on the GLOBAL x extent (x-line twiddles, eigenvalues, norm)
  p->rank = rank;
  p->world = world;
  if (int e = plan_tables_create(p, 2, global_shape, step)) {
    cfd_plan_destroy(p);
    return e;
  }
  p->rank = rank;
  p->world = world;
  p->nx_global = Nxg;
  const SharedLayout L = shared_layout((size_t)nloc, (size_t)Ny);
  if (cudaMalloc(&p->shared, L.total_bytes) != cudaSuccess) {
    cudaGetLastError();
    cfd_plan_destroy(p);
    return set_error_msg("shared slab allocation failed");
  }
  cudaMemset(p->shared, 0, L.total_bytes);
  p->shared_bytes = L.total_bytes;
  p->flags_off = L.off_flags;
  for (int r = 0; r < CFD_MAX_PEERS; ++r) p->peer_shared[r] = p->shared;
  p->dist_push = (world > 1 && push_mode()) ? 1 : 0;
  if (p->dist_push) {
    // the line buffers use the layout of the slab spectrum (plain or pair-interleaved, chosen by
    // choose_t_paired, plan.cu), so every transposed piece is one contiguous run of bytes
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
    bool ok = cudaStreamCreateWithPriority(&p->st_comm, cudaStreamNonBlocking, hi) == cudaSuccess;
    for (int i = 0; i < 8 && ok; ++i)
      ok = cudaEventCreateWithFlags(&p->ev_blk[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&p->ev_chk[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&p->ev_comm_done, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();
      cfd_plan_destroy(p);
      return set_error_msg("push-mode streams could not be created");
    }
  }
  if (world > 1 && staged_mode()) {
    // world blocks of (My / world) x Nloc float2 = one field's bytes (the own block stays unused)
    bool ok = cudaMalloc(&p->xstage, L.field * sizeof(float)) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&p->ev_ready, cudaEventDisableTiming) == cudaSuccess;
    for (int r = 0; r < world && ok; ++r) {
      if (r == rank) continue;
      ok = cudaStreamCreateWithFlags(&p->st_in[r], cudaStreamNonBlocking) == cudaSuccess &&
           cudaStreamCreateWithFlags(&p->st_out[r], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&p->ev_out[r], cudaEventDisableTiming) == cudaSuccess;
      for (int c = 0; c < 8 && ok; ++c)
        ok = cudaEventCreateWithFlags(&p->ev_in[c][r], cudaEventDisableTiming) == cudaSuccess;
    }
    for (int c = 0; c < 8 && ok; ++c)
      ok = cudaEventCreateWithFlags(&p->ev_comp[c], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();
      cfd_plan_destroy(p);
      return set_error_msg("staged-transpose buffers could not be created");
    }
    p->workspace_bytes += L.field * sizeof(float);
  }
  *out = p;
  return 0;
}

size_t cfd_dist_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int cfd_dist_export(cfd_plan* p, void* blob) {
  if (!p || !p->shared || !blob) return set_error_msg("not a distributed plan");
  cudaIpcMemHandle_t h;
  CFD_CUDA_OK(cudaSetDevice(p->device));
  CFD_CUDA_OK(cudaIpcGetMemHandle(&h, p->shared));
  memcpy(blob, &h, sizeof h);
  return 0;
}

int cfd_dist_connect(cfd_plan* p, const void* all_blobs) {
  if (!p || !p->shared || !all_blobs) return set_error_msg("not a distributed plan");
  CFD_CUDA_OK(cudaSetDevice(p->device));
  const char* b = reinterpret_cast<const char*>(all_blobs);
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) {
      p->peer_shared[r] = p->shared;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, b + (size_t)r * sizeof h, sizeof h);
    void* ptr = nullptr;
    CFD_CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->peer_shared[r] = ptr;
  }
  return 0;
}

// dist_state: 0 = nothing loaded, 1 = projected state in VIN, 2 = lazy (u*, v*, q) in slot dist_cur
int cfd_dist_load(cfd_plan* p, cfd_stream stream, const float* const* v_local) {
  if (!p || !p->shared) return set_error_msg("not a distributed plan");
  CFD_CUDA_OK(cudaSetDevice(p->device));
  if (p->ndim == 3) return dist3_load(p, (cudaStream_t)stream, v_local);
  const SharedLayout L = shared_layout((size_t)p->shape[0], (size_t)p->shape[1]);
  for (int a = 0; a < 2; ++a)
    CFD_CUDA_OK(cudaMemcpyAsync(fptr(p->shared, L.off_vin[a]), v_local[a], L.field * sizeof(float),
                                cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  p->dist_state = 1;
  p->dist_cur = 0;
  return 0;
}

int cfd_dist_advance(cfd_plan* p, cfd_stream stream, int nsteps, const cfd_params* params) {
  if (!p || !p->shared || !params) return set_error_msg("not a distributed plan");
  if (p->dist_state == 0) return set_error_msg("cfd_dist_advance: no state loaded");
  CFD_CUDA_OK(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  StepConsts c;
  if (int e = make_consts(p, params, &c)) return e;
  if (p->ndim == 3) return dist3_advance(p, st, nsteps, c);
  for (int t = 0; t < c.n_terms; ++t) {
    if (c.term_kind[t] == CFD_FORCE_FIELD && p->world > 1)
      return set_error_msg("field forcing is not supported on slab-decomposed grids");
    if (c.term_kind[t] == CFD_FORCE_SMAGORINSKY)
      return set_error_msg("the Smagorinsky closure is not supported on slab-decomposed 2-D grids");
  }
  const int nloc = (int)p->shape[0], Ny = (int)p->shape[1], My = Ny / 2;
  const SharedLayout L = shared_layout((size_t)nloc, (size_t)Ny);
  const int prev = (p->rank + p->world - 1) % p->world, next = (p->rank + 1) % p->world;
  auto src3 = [&](size_t off) {
    return SlabSrc{fptr(p->peer_shared[prev], off), fptr(p->shared, off), fptr(p->peer_shared[next], off)};
  };
  LinePeers peers;
  for (int r = 0; r < CFD_MAX_PEERS; ++r)
    peers.p[r] = reinterpret_cast<float2*>(fptr(p->peer_shared[r < p->world ? r : p->rank], L.off_T));
  int lnloc = 0;
  while ((1 << lnloc) < nloc) ++lnloc;
  const size_t lines_per_rank = (size_t)My / p->world;
  if (p->dist_push) {
    for (int n = 0; n < nsteps; ++n)
      if (int e = step_push(p, st, L, c)) return e;
    // the caller's stream owns the plan again only when the comm stream has drained
    CFD_CUDA_OK(cudaEventRecord(p->ev_comm_done, p->st_comm));
    CFD_CUDA_OK(cudaStreamWaitEvent(st, p->ev_comm_done, 0));
    return 0;
  }
  for (int n = 0; n < nsteps; ++n) {
    const int cur = p->dist_cur, nxt = cur ^ 1;
    prof_mark(p, st, "begin");
    if (int e = barrier(p, st)) return e;  // neighbours' inputs (VIN or u*, v*, q) are complete
    if (p->dist_state == 1) {
      const SlabSrc none = {nullptr, nullptr, nullptr};
      if (int e = launch_explicit_2d_slab(st, src3(L.off_vin[0]), src3(L.off_vin[1]), none,
                                          fptr(p->shared, L.off_us[nxt][0]), fptr(p->shared, L.off_us[nxt][1]),
                                          p->rhs, 1, nloc, Ny, p->rank * nloc, (int)p->nx_global, c, 0, 0, -1))
        return e;
    } else {
      if (int e = launch_explicit_2d_slab(st, src3(L.off_us[cur][0]), src3(L.off_us[cur][1]),
                                          src3(L.off_q[cur]), fptr(p->shared, L.off_us[nxt][0]),
                                          fptr(p->shared, L.off_us[nxt][1]), p->rhs, 1, nloc, Ny,
                                          p->rank * nloc, (int)p->nx_global, c, 0, 0, -1))
        return e;
    }
    prof_mark(p, st, "explicit_2d_slab");
    float2* Tloc = reinterpret_cast<float2*>(fptr(p->shared, L.off_T));
    if (int e = launch_rfft_rows(st, p->lm_row, p->rhs, Tloc, 1, nloc, p->tw_row, p->rtw, p->t_paired))
      return e;
    prof_mark(p, st, "rfft_rows");
    if (int e = barrier(p, st)) return e;  // every rank's slab spectrum is complete
    if (p->xstage) {
      if (int e = xpass_staged(p, st, L, lnloc)) return e;
    } else if (int e = launch_xlines_peers(st, p->lm_x, peers, lnloc, (size_t)p->rank * lines_per_rank,
                                           lines_per_rank, My, p->tw_x, p->lam[0], p->lam[1],
                                           p->lamf[0], p->lamf[1], p->fastd, p->cutoff, p->norm,
                                           p->xscratch, p->wbig, &p->side, p->t_paired, nullptr, nullptr)) {
      return e;
    }
    prof_mark(p, st, "xlines_peers");
    if (int e = barrier(p, st)) return e;  // every rank has written its lines back into my slab
    if (int e = launch_irfft_rows(st, p->lm_row, Tloc, fptr(p->shared, L.off_q[nxt]), 1, nloc,
                                  p->tw_row, p->rtw, p->t_paired))
      return e;
    prof_mark(p, st, "irfft_rows");
    p->dist_cur = nxt;
    p->dist_state = 2;
  }
  return 0;
}

int cfd_dist_store(cfd_plan* p, cfd_stream stream, float* const* v_local_out, float* q_local_out) {
  if (!p || !p->shared) return set_error_msg("not a distributed plan");
  if (p->dist_state == 0) return set_error_msg("cfd_dist_store: no state loaded");
  CFD_CUDA_OK(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (p->ndim == 3) return dist3_store(p, st, v_local_out, q_local_out);
  const int nloc = (int)p->shape[0], Ny = (int)p->shape[1];
  const SharedLayout L = shared_layout((size_t)nloc, (size_t)Ny);
  if (p->dist_state == 1) {
    for (int a = 0; a < 2; ++a)
      CFD_CUDA_OK(cudaMemcpyAsync(v_local_out[a], fptr(p->shared, L.off_vin[a]), L.field * sizeof(float),
                                  cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  if (int e = barrier(p, st)) return e;  // the next rank's q is complete
  const int cur = p->dist_cur, next = (p->rank + 1) % p->world;
  if (int e = launch_correct_2d(st, fptr(p->shared, L.off_us[cur][0]), fptr(p->shared, L.off_us[cur][1]),
                                fptr(p->shared, L.off_q[cur]), fptr(p->peer_shared[next], L.off_q[cur]),
                                v_local_out[0], v_local_out[1], 1, nloc, Ny, (float)(1.0 / p->step[0]),
                                (float)(1.0 / p->step[1])))
    return e;
  if (q_local_out)
    CFD_CUDA_OK(cudaMemcpyAsync(q_local_out, fptr(p->shared, L.off_q[cur]), L.field * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
  // peers may still be reading my q for their own store: fence before anyone advances again
  return barrier(p, st);
}

// Per-phase CUDA-event times of `nsteps` slab steps, aggregated by name: milliseconds per STEP (a
// phase that is launched block by block is summed; waits on peers appear as "barrier" / "wait_*").
// Every rank must call it.
int cfd_dist_profile(cfd_plan* p, cfd_stream stream, int nsteps, const cfd_params* params,
                     int max_kernels, float* ms, const char** names, int* n_kernels) {
  if (!p || !p->shared) return set_error_msg("not a distributed plan");
  cudaStream_t st = (cudaStream_t)stream;
  for (auto ev : p->prof_events) cudaEventDestroy(ev);
  p->prof_events.clear();
  p->prof_names.clear();
  p->profiling = true;
  int e = cfd_dist_advance(p, stream, nsteps, params);
  p->profiling = false;
  if (e) return e;
  CFD_CUDA_OK(cudaStreamSynchronize(st));
  std::vector<const char*> uniq;
  std::vector<double> tot;
  std::vector<int> cnt;
  for (size_t i = 1; i < p->prof_events.size(); ++i) {
    const char* nm = p->prof_names[i];
    float t = 0.f;
    CFD_CUDA_OK(cudaEventElapsedTime(&t, p->prof_events[i - 1], p->prof_events[i]));
    size_t k = 0;
    for (; k < uniq.size(); ++k)
      if (strcmp(uniq[k], nm) == 0) break;
    if (k == uniq.size()) {
      uniq.push_back(nm);
      tot.push_back(0.0);
      cnt.push_back(0);
    }
    tot[k] += t;
    cnt[k] += 1;
  }
  const int n = (int)uniq.size();
  if (n_kernels) *n_kernels = n < max_kernels ? n : max_kernels;
  for (int i = 0; i < n && i < max_kernels; ++i) {
    ms[i] = (float)(tot[i] / (nsteps > 0 ? nsteps : 1));
    names[i] = uniq[i];
  }
  return 0;
}

// 0 when no barrier ever timed out
int cfd_dist_check(cfd_plan* p) {
  if (!p || !p->shared) return set_error_msg("not a distributed plan");
  unsigned long long err = 0;
  CFD_CUDA_OK(cudaMemcpy(&err, reinterpret_cast<unsigned long long*>(fptr(p->shared, p->flags_off)) + kSlotErr,
                         sizeof err, cudaMemcpyDeviceToHost));
  if (err) return set_error_msg("a slab barrier timed out (a peer rank did not arrive)");
  return 0;
}

}  // extern "C"
