// Slab-decomposed 3-D time step (BASELINE config #5: 512^3 Taylor-Green with
// the Smagorinsky closure across the GPUs of one box).  Same scheme as the 2-D slab step
// (multi_gpu.cu): one process per GPU, the grid is split along axis 0, and every inter-GPU byte is
// moved by the compute kernels themselves through CUDA-IPC peer mappings over NVLink:
//   * stencil halo: the plane-marching kernels (explicit_3d.cu) read planes -2..-1 / Nloc..Nloc+1
//     of (u, v, w) -- and planes -1 / Nloc of nu_t -- from the neighbouring ranks' buffers
//     (SlabSrc::prev / next); divergence and pressure correction read the single neighbour plane
//     they need (u* below, q above) the same way;
//   * distributed FFT: the z and y transforms are local to a slab and leave the slab spectrum
//     T2[kz][ky][x_loc]; rank r then owns the (kz, ky) lines [r L, (r+1) L) and its x-line kernel
//     assembles each line from all ranks' T2 (element x lives in rank x >> log2(Nloc)), applies
//     fwd * D * inv and writes the line back in place: the two all-to-all transposes are that
//     kernel's loads and stores;
//   * ordering: the one-CTA flag barrier of multi_gpu.cu between the phases.
// The arithmetic is the single-GPU path's, kernel for kernel, so results are bit-identical to it.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "plan_struct.cuh"

namespace cfd {

int launch_smag_nut_3d_slab(cudaStream_t st, SlabSrc su, SlabSrc sv, SlabSrc sw, float* nut, int batch, int N0,
                            int N1, int N2, const StepConsts& c);
int launch_explicit_3d_slab(cudaStream_t st, SlabSrc su, SlabSrc sv, SlabSrc sw, float* us, float* vs, float* ws,
                            int batch, int N0, int N1, int N2, const StepConsts& c, int dvdt_mode, int row0,
                            const SlabSrc* snut);
bool smag_fused();
int launch_smag_acc_3d_slab(cudaStream_t st, SlabSrc su, SlabSrc sv, SlabSrc sw, SlabSrc snut, float* us,
                            float* vs, float* ws, int batch, int N0, int N1, int N2, const StepConsts& c,
                            int dvdt_mode);
int launch_divergence_3d_slab(cudaStream_t st, const float* u, const float* u_below, const float* v,
                              const float* w, float* rhs, int batch, int N0, int N1, int N2, float ihx,
                              float ihy, float ihz);
int launch_correct_3d_slab(cudaStream_t st, const float* us, const float* vs, const float* ws, const float* q,
                           const float* q_above, float* uo, float* vo, float* wo, int batch, int N0, int N1,
                           int N2, float ihx, float ihy, float ihz);
int launch_rfft_rows3(cudaStream_t, int lm, const float* rhs, float2* T, int batch, int NR, const float2* tw,
                      const float2* rtw);
int launch_irfft_rows3(cudaStream_t, int lm, const float2* T, float* q, int batch, int NR, const float2* tw,
                       const float2* rtw);
int launch_lines_scatter(cudaStream_t, int lm, const float2* A, float2* B, int planes, int NL, const float2* tw);
int launch_lines_gather(cudaStream_t, int lm, const float2* B, float2* A, int planes, int NL, const float2* tw);
int launch_xlines3_peers(cudaStream_t st, int lm, const LinePeers& peers, int lnloc, size_t line_begin,
                         size_t nlines, int N1, int NZP, const float2* tw, const double* const* lam,
                         const float* const* lamf, int fastd, double cutoff, float norm, const float* dtab);
bool explicit_3d_uses_march(int N0, int N1, int N2);

namespace {

// layout of the shared (IPC-exported) allocation, identical on every rank; units = floats
struct Layout3 {
  size_t field;  // floats per local field
  size_t off_v[2][3], off_us[3], off_nut, off_q, off_T2, off_flags, total_bytes;
};
Layout3 layout3(size_t nloc, size_t n1, size_t n2) {
  Layout3 L;
  L.field = nloc * n1 * n2;
  size_t o = 0;
  for (int s = 0; s < 2; ++s)
    for (int a = 0; a < 3; ++a) { L.off_v[s][a] = o; o += L.field; }
  for (int a = 0; a < 3; ++a) { L.off_us[a] = o; o += L.field; }
  L.off_nut = o; o += L.field;
  L.off_q = o; o += L.field;
  L.off_T2 = o; o += 2 * (n2 / 2 + 1) * n1 * nloc;  // T2[kz][ky][x_loc], float2
  L.off_flags = o; o += slab_flag_floats();
  L.total_bytes = o * sizeof(float);
  return L;
}
Layout3 layout_of(const cfd_plan* p) {
  return layout3((size_t)p->shape[0], (size_t)p->shape[1], (size_t)p->shape[2]);
}

float* fp(void* base, size_t off) { return reinterpret_cast<float*>(base) + off; }

bool has_smagorinsky(const StepConsts& c) {
  for (int t = 0; t < c.n_terms; ++t)
    if (c.term_kind[t] == CFD_FORCE_SMAGORINSKY) return true;
  return false;
}

}  // namespace

int dist3_plan_create(cfd_plan** out, const int64_t* global_shape, const double* step, int rank, int world,
                      int device) {
  if (!out || !global_shape || !step) return set_error_msg("null argument");
  *out = nullptr;
  if (world < 1 || world > CFD_MAX_PEERS || (world & (world - 1)))
    return set_error_msg("world size must be 1, 2, 4 or 8");
  if (rank < 0 || rank >= world) return set_error_msg("bad rank");
  const int64_t N0g = global_shape[0], N1 = global_shape[1], N2 = global_shape[2];
  if (N0g % world) return set_error_msg("axis 0 must be divisible by the number of ranks");
  const int64_t nloc = N0g / world;
  if (nloc < 16 || (nloc & (nloc - 1))) return set_error_msg("local slab must be a power of two >= 16 planes");
  if (!explicit_3d_uses_march((int)nloc, (int)N1, (int)N2))
    return set_error_msg("slab-decomposed 3-D grids need N1 % 8 == 0 and N2 % 64 == 0");
  if (N1 % (16 * world)) return set_error_msg("axis 1 must be divisible by 16 x the number of ranks");
  if (N0g > (1 << 14)) return set_error_msg("global axis 0 longer than 16384 is not supported in 3-D");
  // an ordinary plan for the LOCAL slab gives the z / y tables and the local workspace (rhs, T1) ...
  int64_t local_shape[3] = {nloc, N1, N2};
  cfd_plan* p = nullptr;
  if (int e = cfd_plan_create_impl(&p, 3, local_shape, step, 1, device, CFD_IMPL_RFFT)) return e;
  // ... then replace what depends on the GLOBAL x extent (x-line twiddles, eigenvalues, norm)
  p->rank = rank;
  p->world = world;
  if (int e = plan_tables_create(p, 3, global_shape, step)) {
    cfd_plan_destroy(p);
    return e;
  }
  p->nx_global = N0g;
  const Layout3 L = layout3((size_t)nloc, (size_t)N1, (size_t)N2);
  if (cudaMalloc(&p->shared, L.total_bytes) != cudaSuccess) {
    cudaGetLastError();
    cfd_plan_destroy(p);
    return set_error_msg("shared slab allocation failed");
  }
  cudaMemset(p->shared, 0, L.total_bytes);
  p->shared_bytes = L.total_bytes;
  p->flags_off = L.off_flags;
  for (int r = 0; r < CFD_MAX_PEERS; ++r) p->peer_shared[r] = p->shared;
  // the local plan's own copies of the buffers that now live in the shared block are not needed
  for (int a = 0; a < 3; ++a) {
    cudaFree(p->us[a]);
    p->us[a] = nullptr;
  }
  cudaFree(p->T2);
  p->T2 = nullptr;
  cudaFree(p->qbuf);
  p->qbuf = nullptr;
  p->workspace_bytes = L.total_bytes;
  *out = p;
  return 0;
}

int dist3_load(cfd_plan* p, cudaStream_t st, const float* const* v_local) {
  const Layout3 L = layout_of(p);
  // peers may still be reading the slot that is overwritten here (their last step's halo loads)
  if (int e = slab_barrier(p, st)) return e;
  for (int a = 0; a < 3; ++a)
    CFD_CUDA_OK(cudaMemcpyAsync(fp(p->shared, L.off_v[0][a]), v_local[a], L.field * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
  p->dist_state = 1;
  p->dist_cur = 0;
  return 0;
}

int dist3_advance(cfd_plan* p, cudaStream_t st, int nsteps, const StepConsts& c) {
  for (int t = 0; t < c.n_terms; ++t)
    if (c.term_kind[t] == CFD_FORCE_FIELD && p->world > 1)
      return set_error_msg("field forcing is not supported on slab-decomposed grids");
  const int nloc = (int)p->shape[0], N1 = (int)p->shape[1], N2 = (int)p->shape[2];
  const int NZP = N2 / 2 + 1;
  const Layout3 L = layout_of(p);
  const int W = p->world, rank = p->rank;
  const int prev = (rank + W - 1) % W, next = (rank + 1) % W;
  auto src3 = [&](size_t off) {
    return SlabSrc{fp(p->peer_shared[prev], off), fp(p->shared, off), fp(p->peer_shared[next], off)};
  };
  const size_t planeN = (size_t)N1 * N2;
  const float ih[3] = {c.inv_h[0], c.inv_h[1], c.inv_h[2]};
  LinePeers peers;
  for (int r = 0; r < CFD_MAX_PEERS; ++r)
    peers.p[r] = reinterpret_cast<float2*>(fp(p->peer_shared[r < W ? r : rank], L.off_T2));
  int lnloc = 0;
  while ((1 << lnloc) < nloc) ++lnloc;
  const size_t lines_per_rank = (size_t)NZP * N1 / W;
  const bool smag = has_smagorinsky(c);
  float* us[3] = {fp(p->shared, L.off_us[0]), fp(p->shared, L.off_us[1]), fp(p->shared, L.off_us[2])};
  float* nut = fp(p->shared, L.off_nut);
  float* q = fp(p->shared, L.off_q);
  float2* T2 = reinterpret_cast<float2*>(fp(p->shared, L.off_T2));

  for (int n = 0; n < nsteps; ++n) {
    const int cur = p->dist_cur, nxt = cur ^ 1;
    const SlabSrc su = src3(L.off_v[cur][0]), sv = src3(L.off_v[cur][1]), sw = src3(L.off_v[cur][2]);
    prof_mark(p, st, "begin");
    if (int e = slab_barrier(p, st)) return e;  // the neighbours' velocity (previous correction) is complete
    if (smag) {
      if (int e = launch_smag_nut_3d_slab(st, su, sv, sw, nut, 1, nloc, N1, N2, c)) return e;
      prof_mark(p, st, "smag_nut");
    }
    const SlabSrc snut = src3(L.off_nut);
    const bool fused = smag && smag_fused();
    if (fused) {
      if (int e = slab_barrier(p, st)) return e;  // the neighbours' nu_t planes are complete
    }
    if (int e = launch_explicit_3d_slab(st, su, sv, sw, us[0], us[1], us[2], 1, nloc, N1, N2, c, 0, rank * nloc,
                                        fused ? &snut : nullptr))
      return e;
    prof_mark(p, st, "explicit_3d");
    if (smag && !fused) {
      if (int e = slab_barrier(p, st)) return e;  // the neighbours' nu_t planes are complete
      if (int e = launch_smag_acc_3d_slab(st, su, sv, sw, snut, us[0], us[1], us[2], 1, nloc, N1, N2, c, 0)) return e;
      prof_mark(p, st, "smag_acc");
    }
    if (int e = slab_barrier(p, st)) return e;  // the previous rank's last u* plane is complete
    const float* u_below = fp(p->peer_shared[prev], L.off_us[0]) + (size_t)(nloc - 1) * planeN;
    if (int e = launch_divergence_3d_slab(st, us[0], u_below, us[1], us[2], p->rhs, 1, nloc, N1, N2, ih[0], ih[1],
                                          ih[2]))
      return e;
    prof_mark(p, st, "divergence_3d");
    if (int e = launch_rfft_rows3(st, p->lm_row, p->rhs, p->T, 1, nloc * N1, p->tw_row, p->rtw)) return e;
    prof_mark(p, st, "rfft_z");
    if (int e = launch_lines_scatter(st, p->lm_y, p->T, T2, NZP, nloc, p->tw_y)) return e;
    prof_mark(p, st, "fft_y");
    if (int e = slab_barrier(p, st)) return e;  // every rank's slab spectrum is complete
    if (int e = launch_xlines3_peers(st, p->lm_x, peers, lnloc, (size_t)rank * lines_per_rank, lines_per_rank, N1,
                                     NZP, p->tw_x, p->lam, p->lamf, p->fastd, p->cutoff, p->norm, nullptr))
      return e;
    prof_mark(p, st, "xlines3_peers");
    if (int e = slab_barrier(p, st)) return e;  // every rank has written its lines back into my slab
    if (int e = launch_lines_gather(st, p->lm_y, T2, p->T, NZP, nloc, p->tw_y)) return e;
    prof_mark(p, st, "ifft_y");
    if (int e = launch_irfft_rows3(st, p->lm_row, p->T, q, 1, nloc * N1, p->tw_row, p->rtw)) return e;
    prof_mark(p, st, "irfft_z");
    if (int e = slab_barrier(p, st)) return e;  // the next rank's first q plane is complete
    const float* q_above = fp(p->peer_shared[next], L.off_q);
    if (int e = launch_correct_3d_slab(st, us[0], us[1], us[2], q, q_above, fp(p->shared, L.off_v[nxt][0]),
                                       fp(p->shared, L.off_v[nxt][1]), fp(p->shared, L.off_v[nxt][2]), 1, nloc, N1,
                                       N2, ih[0], ih[1], ih[2]))
      return e;
    prof_mark(p, st, "correct_3d");
    p->dist_cur = nxt;
  }
  return 0;
}

int dist3_store(cfd_plan* p, cudaStream_t st, float* const* v_local_out, float* q_local_out) {
  const Layout3 L = layout_of(p);
  const int cur = p->dist_cur;
  for (int a = 0; a < 3; ++a)
    CFD_CUDA_OK(cudaMemcpyAsync(v_local_out[a], fp(p->shared, L.off_v[cur][a]), L.field * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
  if (q_local_out)
    CFD_CUDA_OK(cudaMemcpyAsync(q_local_out, fp(p->shared, L.off_q), L.field * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace cfd
