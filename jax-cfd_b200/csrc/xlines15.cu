// X pass for lines of 2^15 points (Nx = 32768: BASELINE config #4 and the 4- / 8-GPU members of
// the weak-scaling family) on thread-block clusters with a distributed-shared-memory exchange.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "fft_rows.cuh"
#include "fft_smem.cuh"
#include "xline_scale.cuh"

namespace cg = cooperative_groups;

namespace cfd {

namespace {

// ------------------------------------------------------------------------------------------
// 32768-point x lines on a thread-block CLUSTER (CFD_X15=split selects the scratch path of
// poisson_2d.cu instead).  A line does not fit one CTA (16384 points = 1024 threads x 16 registers
// is the most), so one radix-2 decimation-in-frequency step splits it over the CTAs of a cluster:
//   in:   CTA h forms y_h[m] = (x[m] +- x[m + N/2]) w^(h m) for all m < N/2 straight from the
//         spectrum -- both CTAs of a line load both halves; the second copy of every sector comes
//         out of L2, which measured faster than handing the halves over through DSMEM;
//   mid:  y_0 / y_1 are ordinary 16384-point lines whose transforms are the even / odd frequencies:
//         forward passes, pseudo-inverse scaling (kx = 2 d + h), inverse passes, as in xlines_kernel;
//   out:  x'[m] = y_0' + y_1' conj(w^m), x'[m + N/2] = y_0' - y_1' conj(w^m): every CTA publishes its
//         half-length result in its own shared memory, cluster barrier, and reads the partner's
//         through distributed shared memory (ld.shared::cluster).
// Operation for operation this is split_lines_kernel -> xlines_kernel<14, SPLIT> ->
// merge_lines_kernel (bit-identical results) without the scratch: the spectrum is read once from
// HBM and written once, 8 B/cell instead of 24.
//   plain layout:   cluster of 2 CTAs = one line.
//   PAIRED layout:  cluster of 4 CTAs = the two interleaved lines of a pair, CTA rank = 2 l + h.
//                   Loads are the stride-2 accesses of the layout (the partner line's CTAs use the
//                   other half of every sector at the same time); for the STORES -- posted NVLink
//                   writes to the slab owners on several GPUs, where half-filled sectors would
//                   double the link traffic -- each CTA gathers both lines from all four CTAs so
//                   that every thread writes whole float4 = both lines of the pair.
// Cluster barrier halves: work that does not depend on the partners (register-only passes, global
// stores) runs between the arrive and the wait.
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <bool FASTD, bool PAIRED>
__global__ void __launch_bounds__(1024, 1)
xlines15_cluster_kernel(LinePeers peers, LinePeers peers_w, int lnloc, size_t line_begin, int My,
                        const float2* __restrict__ tw, const double* __restrict__ lamx,
                        const double* __restrict__ lamy, const float* __restrict__ lamxf,
                        const float* __restrict__ lamyf, double cutoff, float norm,
                        const float* __restrict__ dtab, const float2* __restrict__ wbig) {
  using P = FftPlan<14, 4, 4>;
  constexpr int M = P::M, G = P::G, E = P::E;  // 16384-point half lines, 1024 threads x 16 points
  static_assert(G == 1024, "one half line per CTA");
  extern __shared__ float2 smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank();
  const int h = (int)(crank & 1u);                 // half of x held on entry = parity of the frequencies
  const int l = PAIRED ? (int)(crank >> 1) : 0;    // line within the pair
  constexpr int CS = PAIRED ? 4 : 2;
  const int t = threadIdx.x;
  const size_t unit = blockIdx.x / CS;             // line (plain) or pair (PAIRED) handled by this cluster
  const size_t line = line_begin + (PAIRED ? 2 * unit + l : unit);
  const int ky = (int)(line % My);
  const int nloc_mask = (1 << lnloc) - 1;
  constexpr int XS = PAIRED ? 2 : 1;
  const size_t loff = PAIRED ? (((line >> 1) << lnloc) << 1) + (line & 1) : line << lnloc;
  __shared__ float2* s_peer[CFD_MAX_PEERS];
  __shared__ float2* s_peer_w[CFD_MAX_PEERS];
  if (t < CFD_MAX_PEERS) {
    s_peer[t] = peers.p[t];
    s_peer_w[t] = peers_w.p[t];
  }
  __syncthreads();
  auto elem = [&](int x) -> float2* { return s_peer[x >> lnloc] + loff + XS * (x & nloc_mask); };
  auto elem_w = [&](int x) -> float2* { return s_peer_w[x >> lnloc] + loff + XS * (x & nloc_mask); };

  float2 v[E];
  // ---- radix-2 decimation in frequency: both CTAs of a line read BOTH halves of it straight from
  // the spectrum (every sector is requested by the 2 / 4 CTAs of the cluster at about the same time:
  // one HBM read, the other copies come out of L2) -- no exchange on the way in
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int m = t + G * e;
    const float2 a = *elem(m), b = *elem(M + m);
    if (h == 0) {
      v[e] = make_float2(a.x + b.x, a.y + b.y);  // y0[m] = x[m] + x[m + N/2]
    } else {
      const float2 d = make_float2(a.x - b.x, a.y - b.y);
      v[e] = cmul(d, __ldg(wbig + m));           // y1[m] = (x[m] - x[m + N/2]) w^m
    }
  }
  fft_pass_compute<P, P::lr_fwd(0), P::lns_fwd(0), -1>(v, t, tw + P::tw_off_fwd(0));
  FftRun<P, -1, SyncCta, false, 0, false>::template exchange<0>(v, t, smem, 0, 0);
  FftRun<P, -1, SyncCta, false, 0, false>::template passes<1>(v, t, smem, tw, 0, 0);
  scale_line<P, FASTD>(v, t, smem, ky, My, ky == 0, 2, h, lamx, lamy, lamxf, lamyf, cutoff, norm,
                       FASTD ? nullptr : dtab);
  FftRun<P, +1, SyncCta, false, (P::NP - 1) & 1, false>::run(v, t, smem, tw, 0, 0);
  // ---- merge: x'[m] = y0' + y1' conj(w^m),  x'[m + N/2] = y0' - y1' conj(w^m)
  if (h == 1) {
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = cmulc(v[e], __ldg(wbig + t + G * e));
  }
  __syncthreads();  // every thread is done with the last exchange of the inverse transform
#pragma unroll
  for (int e = 0; e < E; ++e) smem[P::pad(t + G * e)] = v[e];
  cluster_arrive_release();
  cluster_wait();
  // Both operands come back from shared memory into the (now dead) data registers; the arrive
  // that lets the partners go ("I have read your shared memory") is issued BEFORE the global
  // stores, so that its release fence never waits for an HBM / NVLink round trip.
  if constexpr (!PAIRED) {
    const float2* y0 = cluster.map_shared_rank(smem, 0u);
    const float2* y1 = cluster.map_shared_rank(smem, 1u);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = P::pad(t + G * e);
      const float2 a = y0[i], b = y1[i];
      v[e] = h == 0 ? make_float2(a.x + b.x, a.y + b.y) : make_float2(a.x - b.x, a.y - b.y);
    }
    cluster_arrive_release();
#pragma unroll
    for (int e = 0; e < E; ++e) *elem_w(h * M + t + G * e) = v[e];
  } else {
    // CTA (l, h) writes the float4s (both lines of the pair) of half h whose slot e has parity l
    const float2* yA0 = cluster.map_shared_rank(smem, 0u);
    const float2* yA1 = cluster.map_shared_rank(smem, 1u);
    const float2* yB0 = cluster.map_shared_rank(smem, 2u);
    const float2* yB1 = cluster.map_shared_rank(smem, 3u);
    const size_t poff = ((line >> 1) << lnloc) << 1;
#pragma unroll
    for (int e2 = 0; e2 < E / 2; ++e2) {
      const int i = P::pad(t + G * (2 * e2 + l));
      const float2 a0 = yA0[i], a1 = yA1[i], b0 = yB0[i], b1 = yB1[i];
      if (h == 0) {
        v[2 * e2] = make_float2(a0.x + a1.x, a0.y + a1.y);
        v[2 * e2 + 1] = make_float2(b0.x + b1.x, b0.y + b1.y);
      } else {
        v[2 * e2] = make_float2(a0.x - a1.x, a0.y - a1.y);
        v[2 * e2 + 1] = make_float2(b0.x - b1.x, b0.y - b1.y);
      }
    }
    cluster_arrive_release();
#pragma unroll
    for (int e2 = 0; e2 < E / 2; ++e2) {
      const int x = h * M + t + G * (2 * e2 + l);
      *reinterpret_cast<float4*>(s_peer_w[x >> lnloc] + poff + 2 * (x & nloc_mask)) =
          make_float4(v[2 * e2].x, v[2 * e2].y, v[2 * e2 + 1].x, v[2 * e2 + 1].y);
    }
  }
  cluster_wait();  // nobody leaves while its shared memory may still be read
}


}  // namespace

// CFD_X15=split selects the scratch path (split -> 16384-point lines -> merge) for 32768-point lines
// Which implementation runs 32768-point lines: CFD_X15=cluster|split forces one.  Default: the
// cluster kernel, except for the pair-interleaved layout on ONE GPU.  Measured x pass (B200):
//   plain,  1 GPU, 32768 x 8192:              cluster 1.66 ms, scratch path 2.46 ms
//   paired, 1 GPU, 32768 x 16384:             cluster 4.48 ms, scratch path 4.00 ms (the 4-CTA output
//                                             exchange moves 3/4 of its data through DSMEM)
//   paired, 8 GPUs, 32768 x 16384 / 32768^2:  cluster 0.64 / 1.24 ms, scratch path 0.74 / 1.45 ms
bool x15_cluster(int paired, int several_gpus) {
  static const int forced = [] {
    const char* e = getenv("CFD_X15");
    if (e && strcmp(e, "split") == 0) return 0;
    if (e && strcmp(e, "cluster") == 0) return 1;
    return -1;
  }();
  if (forced >= 0) return forced == 1;
  return !paired || several_gpus;
}

int launch_xlines15_cluster(cudaStream_t st, const LinePeers& peers, const LinePeers& peers_w, int lnloc,
                                   size_t line_begin, size_t nlines, int My, const float2* tw, const double* lamx,
                                   const double* lamy, const float* lamxf, const float* lamyf, int fastd,
                                   double cutoff, float norm, const float2* wbig, int paired, const float* dtab) {
  using P = FftPlan<14, 4, 4>;
  constexpr size_t smem = (size_t)row_stride(P::M, 16) * sizeof(float2);
  if (dtab) fastd = 0;
  if (paired && ((line_begin | nlines) & 1)) return set_error_msg("internal: paired lines need even line ranges");
  auto go = [&](auto k, unsigned cs) -> int {
    if (int e = set_smem(k, smem)) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * nlines));
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CFD_CUDA_OK(cudaLaunchKernelEx(&cfg, k, peers, peers_w, lnloc, line_begin, My, tw, lamx, lamy, lamxf, lamyf,
                                   cutoff, norm, dtab, wbig));
    count_launch();
    return 0;
  };
  if (paired) return fastd ? go(xlines15_cluster_kernel<true, true>, 4u) : go(xlines15_cluster_kernel<false, true>, 4u);
  return fastd ? go(xlines15_cluster_kernel<true, false>, 2u) : go(xlines15_cluster_kernel<false, false>, 2u);
}

}  // namespace cfd
