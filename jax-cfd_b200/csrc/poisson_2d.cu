// Sweeps "R", "X", "C" (2-D): the fast-diagonalisation pressure projection
//   q = irfftn( D * rfftn(rhs) ),  v' = u* - grad q
// replacing fast_diagonalization.py:199-225,257-262 (`_circulant_rfft_transform`, pseudoinverse),
// pressure.py:115-157 (`solve_fast_diag`) and pressure.py:181-198 (`projection`).
//
//   R  rfft_rows:   real FFT of every row (axis 1, contiguous) of rhs; the half spectrum is
//                   written TRANSPOSED, T[ky][x], packed to exactly Ny/2 complex columns
//                   (Re/Im of column 0 hold the two real coefficients ky = 0 and ky = Ny/2).
//   X  xlines:      for every ky: contiguous complex FFT along x, multiply by the pseudo-inverse
//                   eigenvalue table D(kx, ky) (built in f64 from the analytic circulant
//                   eigenvalues), inverse FFT along x, in place.  One read + one write of T.
//   Q  irfft_rows:  gather T[ky][x] for the CTA's rows, inverse real FFT, coalesced q store.
//   C  correct2d:   v' = u* - forward_difference(q) (pressure.py:194-196); only run when a
//                   projected state must be materialised (chained steps project lazily inside
//                   the next explicit kernel).
//
// Scaling: R stores 2*X, C's pre-processing produces 2*Z, the inverse passes are unnormalised, so
// the table D carries 1 / (2 * Nx * Ny) (a power of two: exact).
#include "common.cuh"
#include "fft_smem.cuh"
#include "fft_rows.cuh"
#include "xline_scale.cuh"

#include <stdlib.h>
#include <string.h>

#include <type_traits>

namespace cfd {

namespace {

// ------------------------------------------------------------------------------------------
// Layouts of the packed spectrum T (M = Ny/2 lines per batch member, M even):
//   plain:   T[b][ky][x]
//   PAIRED:  T[b][ky >> 1][x][ky & 1]   pairs of ky lines interleaved, so that the transposed
//            accesses of the row kernels (ROWS consecutive x for one ky) are chunks of ROWS * 16
//            bytes instead of ROWS * 8.  An x line is then a stride-2 sequence (the partner line's
//            CTA uses the other half of every sector).  The plan picks the layout
//            (choose_t_paired, plan.cu, with the measurements).
//
// R: rows of `rhs` (N = 2M reals each)  ->  T   (ky = 0..M-1, packed)
// LEMAX = 5: 32 points (two radix-16 butterflies per pass) per thread, half the threads per row.
template <int LM, int ROWS, int LEMAX, bool PAIRED>
__global__ void __launch_bounds__(ROWS * FftPlan<LM, LEMAX, 4>::G, (ROWS * FftPlan<LM, LEMAX, 4>::G <= 512) ? 2 : 0)
rfft_rows_kernel(const float* __restrict__ rhs, float2* __restrict__ T, int Nx,
                 const float2* __restrict__ tw, const float2* __restrict__ rtw, int x_begin) {
  using P = FftPlan<LM, LEMAX, 4>;
  constexpr int M = P::M, G = P::G, E = P::E;
  constexpr int RS = row_stride(M, ROWS);
  extern __shared__ float2 smem[];
  const int tid = threadIdx.x;
  const int row = tid / G, t = tid % G;
  const int x0 = x_begin + blockIdx.x * ROWS;  // rows [x_begin, x_begin + gridDim.x * ROWS)
  const size_t b = blockIdx.y;
  float2* s = smem + row * RS;

  float2 v[E];
  {
    const float2* src = reinterpret_cast<const float2*>(rhs + (b * Nx + x0 + row) * (size_t)(2 * M));
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = __ldg(src + t + G * e);
  }
  using LSYNC = std::conditional_t<(ROWS <= 15), SyncLine<G>, SyncCta>;
  FftRun<P, -1, LSYNC>::run(v, t, s, tw, row);
  LSYNC::sync(row);
#pragma unroll
  for (int e = 0; e < E; ++e) s[PAD(t + G * e)] = v[e];
  __syncthreads();
  // Split of the half-size complex transform into the real-input spectrum (pairs k, M - k), merged
  // with the transposed store: every pair is read from the line buffers once and goes straight to
  // T (no in-place split pass: two shared-memory traversals fewer).  For each ky the ROWS values of
  // this CTA are contiguous in T.
  float2* Tb = T + b * (size_t)M * Nx + (PAIRED ? 2 : 1) * (size_t)x0;
  auto tptr = [&](int ky, int r) -> float2* {
    return PAIRED ? Tb + ((size_t)(ky >> 1) * Nx + r) * 2 + (ky & 1) : Tb + (size_t)ky * Nx + r;
  };
  constexpr int NT = ROWS * G;
#pragma unroll 4
  for (int idx = tid; idx < ROWS * (M / 2); idx += NT) {
    int r, k;
    if constexpr (PAIRED) {
      r = (idx >> 1) % ROWS;
      k = 2 * (idx / (2 * ROWS)) + (idx & 1);
    } else {
      r = idx % ROWS;
      k = idx / ROWS;
    }
    const float2* sr = smem + r * RS;
    if (k == 0) {
      const float2 z = sr[0], zh = sr[PAD(M / 2)];
      *tptr(0, r) = make_float2(2.f * (z.x + z.y), 2.f * (z.x - z.y));
      *tptr(M / 2, r) = make_float2(2.f * zh.x, -2.f * zh.y);
    } else {
      const float2 zk = sr[PAD(k)], zm = sr[PAD(M - k)];
      const float2 A = make_float2(zk.x + zm.x, zk.y - zm.y);  // Zk + conj(Zm)
      const float2 B = make_float2(zk.x - zm.x, zk.y + zm.y);  // Zk - conj(Zm)
      const float2 WB = cmul(__ldg(rtw + k), B);               // (-i w^k) B
      *tptr(k, r) = make_float2(A.x + WB.x, A.y + WB.y);
      *tptr(M - k, r) = make_float2(A.x - WB.x, -(A.y - WB.y));
    }
  }
}

// ------------------------------------------------------------------------------------------
// X: lines T[line][0..Nx)  (line = b * My + ky), forward FFT * D * inverse FFT, in place.
//
// Multi-GPU: the line is assembled from the slab-local spectra of all ranks -- element x lives in
// rank (x >> lnloc)'s buffer at [line][x & (Nloc-1)] -- so the loads/stores below ARE the
// all-to-all transpose of the distributed FFT, done with peer accesses over NVLink inside the
// kernel (each rank transforms its own range of ky lines).  One GPU: a single peer, Nloc = M.
// lines per CTA of the x-line kernel for short lines (a CTA is at most 256 threads).  8 instead of
// 16 (256-point lines: 128-thread CTAs): 1024 x 256^2, x pass 0.151 -> 0.129 ms.
#ifndef CFD_XL_LINES_MAX
#define CFD_XL_LINES_MAX 8
#endif
#ifndef CFD_XL_PRETW
#define CFD_XL_PRETW true
#endif
#ifndef CFD_XL_MINB
#define CFD_XL_MINB 1
#endif
// BAL: balanced radix schedule with radix-32 passes (fft_rows.cuh, xlines_balanced)
template <int LM, int LINES, bool FASTD, int LEMAX, bool SPLIT, bool DB, bool PAIRED, bool BAL = false>
__global__ void __launch_bounds__(LINES * FftPlan<LM, LEMAX, 4>::G,
                                  (LEMAX == 5 && LINES * FftPlan<LM, LEMAX, 4>::G <= 256) ? 2 : CFD_XL_MINB)
xlines_kernel(LinePeers peers, LinePeers peers_w, int lnloc, size_t line_begin, int My,
              const float2* __restrict__ tw, const double* __restrict__ lamx,
              const double* __restrict__ lamy, const float* __restrict__ lamxf,
              const float* __restrict__ lamyf, double cutoff, float norm,
              const float* __restrict__ dtab) {
  using P = FftPlan<LM, LEMAX, BAL ? 5 : 4, BAL>;
  constexpr int M = P::M, G = P::G, E = P::E;
  constexpr int RS = row_stride(M, 16);
  extern __shared__ float2 smem[];
  const int tid = threadIdx.x;
  const int ln = tid / G, t = tid % G;
  // split == 1: the "lines" are the half-length lines (y0, y1 interleaved per original line) of
  // the 32768-point transform in a local scratch; original line = line_begin + (index >> 1).
  const size_t idx0 = (size_t)blockIdx.x * LINES;
  constexpr bool split = SPLIT;
  const size_t line0 = split ? line_begin + (idx0 >> 1) : line_begin + idx0;
  const size_t line = split ? line_begin + ((idx0 + ln) >> 1) : line0 + ln;
  const int ky = (int)(line % My);
  const int kmul = split ? 2 : 1, kadd = split ? (int)((idx0 + ln) & 1) : 0;
  float2* s = smem + ln * RS;
  const int nloc_mask = (1 << lnloc) - 1;
  // this line's offset inside every rank's buffer (PAIRED: pair-interleaved, element stride 2);
  // split: inside the scratch (plain contiguous half-lines)
  const size_t loff = split ? (idx0 + ln) << LM
                            : (PAIRED ? (((line >> 1) << lnloc) << 1) + (line & 1) : line << lnloc);
  constexpr int XS = (!SPLIT && PAIRED) ? 2 : 1;
  // One GPU (lnloc == LM): plain contiguous line.  Several GPUs: peer table in shared memory (a
  // dynamically indexed kernel parameter would live in local memory).
  // peers: where the line is READ; peers_w: where the result is WRITTEN (the same buffers in place,
  // or -- slab decomposition, push mode -- local receive buffers in, the slab owners' spectra out)
  __shared__ float2* s_peer[CFD_MAX_PEERS];
  __shared__ float2* s_peer_w[CFD_MAX_PEERS];
  const bool single = (lnloc == LM);
  if (!single) {
    if (tid < CFD_MAX_PEERS) {
      s_peer[tid] = peers.p[tid];
      s_peer_w[tid] = peers_w.p[tid];
    }
    __syncthreads();
  }
  float2* const Tl = peers.p[0] + loff;
  auto elem = [&](int x) -> float2* {
    return single ? Tl + XS * x : s_peer[x >> lnloc] + loff + XS * (x & nloc_mask);
  };
  auto elem_w = [&](int x) -> float2* {
    return single ? Tl + XS * x : s_peer_w[x >> lnloc] + loff + XS * (x & nloc_mask);
  };

  float2 v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) v[e] = *elem(t + G * e);
  // DB: second exchange buffer behind the first (see FftRun); the inverse continues the
  // forward transform's buffer alternation
  constexpr int ALT = LINES * RS;
  // twiddle loads ahead of the exchange only where the kernel is not capped at 64 registers
  constexpr bool PRE = CFD_XL_PRETW && LEMAX == 4 && (LINES * G <= 512);
  FftRun<P, -1, SyncCta, DB, 0, PRE>::run(v, t, s, tw, 0, ALT);

  // only the first line(s) of a CTA can be ky = 0 (split: both halves of that line)
  const bool cta_has_packed = (line0 % My) == 0;
  scale_line<P, FASTD>(v, t, s, ky, My, cta_has_packed, kmul, kadd, lamx, lamy, lamxf, lamyf, cutoff,
                       norm, FASTD ? nullptr : dtab);
  if (DB && cta_has_packed) __syncthreads();  // scale_line's reads of buffer 0 are done
  FftRun<P, +1, SyncCta, DB, (P::NP - 1) & 1, PRE>::run(v, t, s, tw, 0, ALT);
#pragma unroll
  for (int e = 0; e < E; ++e) *elem_w(t + G * e) = v[e];
}

// ------------------------------------------------------------------------------------------
// Lines of 2^15 points (Nx = 32768) do not fit one CTA: one radix-2 decimation-in-frequency step
//   y0[m] = x[m] + x[m + N/2],  y1[m] = (x[m] - x[m + N/2]) w_N^m              (split_lines_kernel)
// turns a line into two half-length lines whose transforms are the even / odd frequencies; they
// are scaled and inverse-transformed by xlines_kernel (split mode: kx = 2 d + parity) in a local
// scratch, and  x'[m] = y0' + y1' conj(w^m),  x'[m + N/2] = y0' - y1' conj(w^m)   (merge_lines_kernel).
// The spectrum itself (peer memory on several GPUs) is still read once and written once.
__global__ void split_lines_kernel(LinePeers peers, int lnloc, size_t line_begin, int half,
                                   float2* __restrict__ scratch, const float2* __restrict__ wbig) {
  __shared__ float2* s_peer[CFD_MAX_PEERS];
  if (threadIdx.x < CFD_MAX_PEERS) s_peer[threadIdx.x] = peers.p[threadIdx.x];
  __syncthreads();
  const size_t li = blockIdx.y;
  const size_t loff = (line_begin + li) << lnloc;
  const int nloc_mask = (1 << lnloc) - 1;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= half) return;
  const int x1 = m + half;
  const float2 a = s_peer[m >> lnloc][loff + (m & nloc_mask)];
  const float2 b = s_peer[x1 >> lnloc][loff + (x1 & nloc_mask)];
  const float2 w = __ldg(wbig + m);
  const float2 d = make_float2(a.x - b.x, a.y - b.y);
  float2* y = scratch + li * (size_t)(2 * half);
  y[m] = make_float2(a.x + b.x, a.y + b.y);
  y[half + m] = cmul(d, w);
}

__global__ void merge_lines_kernel(LinePeers peers, int lnloc, size_t line_begin, int half,
                                   const float2* __restrict__ scratch, const float2* __restrict__ wbig) {
  __shared__ float2* s_peer[CFD_MAX_PEERS];
  if (threadIdx.x < CFD_MAX_PEERS) s_peer[threadIdx.x] = peers.p[threadIdx.x];
  __syncthreads();
  const size_t li = blockIdx.y;
  const size_t loff = (line_begin + li) << lnloc;
  const int nloc_mask = (1 << lnloc) - 1;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= half) return;
  const int x1 = m + half;
  const float2* y = scratch + li * (size_t)(2 * half);
  const float2 y0 = y[m];
  const float2 y1 = cmulc(y[half + m], __ldg(wbig + m));  // * conj(w^m)
  s_peer[m >> lnloc][loff + (m & nloc_mask)] = make_float2(y0.x + y1.x, y0.y + y1.y);
  s_peer[x1 >> lnloc][loff + (x1 & nloc_mask)] = make_float2(y0.x - y1.x, y0.y - y1.y);
}

// The same two steps on the PAIRED layout: a thread handles element m of BOTH lines of a pair
// (one float4 = the two interleaved lines), so the spectrum is still read and written in full
// 16-byte pieces; blockIdx.y counts pairs, line_begin is even.
__global__ void split_pairs_kernel(LinePeers peers, int lnloc, size_t line_begin, int half,
                                   float2* __restrict__ scratch, const float2* __restrict__ wbig) {
  __shared__ float2* s_peer[CFD_MAX_PEERS];
  if (threadIdx.x < CFD_MAX_PEERS) s_peer[threadIdx.x] = peers.p[threadIdx.x];
  __syncthreads();
  const size_t pi = blockIdx.y;
  const size_t loff = (((line_begin >> 1) + pi) << lnloc) << 1;
  const int nloc_mask = (1 << lnloc) - 1;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= half) return;
  const int x1 = m + half;
  const float4 a = *reinterpret_cast<const float4*>(s_peer[m >> lnloc] + loff + 2 * (m & nloc_mask));
  const float4 b = *reinterpret_cast<const float4*>(s_peer[x1 >> lnloc] + loff + 2 * (x1 & nloc_mask));
  const float2 w = __ldg(wbig + m);
  float2* y = scratch + (2 * pi) * (size_t)(2 * half);  // lines 2 pi and 2 pi + 1 of this chunk
  y[m] = make_float2(a.x + b.x, a.y + b.y);
  y[half + m] = cmul(make_float2(a.x - b.x, a.y - b.y), w);
  y += 2 * half;
  y[m] = make_float2(a.z + b.z, a.w + b.w);
  y[half + m] = cmul(make_float2(a.z - b.z, a.w - b.w), w);
}

__global__ void merge_pairs_kernel(LinePeers peers, int lnloc, size_t line_begin, int half,
                                   const float2* __restrict__ scratch, const float2* __restrict__ wbig) {
  __shared__ float2* s_peer[CFD_MAX_PEERS];
  if (threadIdx.x < CFD_MAX_PEERS) s_peer[threadIdx.x] = peers.p[threadIdx.x];
  __syncthreads();
  const size_t pi = blockIdx.y;
  const size_t loff = (((line_begin >> 1) + pi) << lnloc) << 1;
  const int nloc_mask = (1 << lnloc) - 1;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= half) return;
  const int x1 = m + half;
  const float2 w = __ldg(wbig + m);
  const float2* y = scratch + (2 * pi) * (size_t)(2 * half);
  const float2 a0 = y[m], a1 = cmulc(y[half + m], w);  // * conj(w^m)
  y += 2 * half;
  const float2 b0 = y[m], b1 = cmulc(y[half + m], w);
  *reinterpret_cast<float4*>(s_peer[m >> lnloc] + loff + 2 * (m & nloc_mask)) =
      make_float4(a0.x + a1.x, a0.y + a1.y, b0.x + b1.x, b0.y + b1.y);
  *reinterpret_cast<float4*>(s_peer[x1 >> lnloc] + loff + 2 * (x1 & nloc_mask)) =
      make_float4(a0.x - a1.x, a0.y - a1.y, b0.x - b1.x, b0.y - b1.y);
}

// ------------------------------------------------------------------------------------------
// Q: gather ROWS rows of T, inverse real FFT, write q rows (coalesced).  No halo row, no
// redundant transform; the pressure-gradient correction is applied either by correct2d_kernel or
// lazily by the next step's explicit kernel (explicit_2d.cu, LAZY mode).
template <int LM, int ROWS, int LEMAX, bool PAIRED>
__global__ void __launch_bounds__(ROWS * FftPlan<LM, LEMAX, 4>::G, (ROWS * FftPlan<LM, LEMAX, 4>::G <= 512) ? 2 : 0)
irfft_rows_kernel(const float2* __restrict__ T, float* __restrict__ q, int Nx,
                  const float2* __restrict__ tw, const float2* __restrict__ rtw) {
  using P = FftPlan<LM, LEMAX, 4>;
  constexpr int M = P::M, G = P::G, E = P::E;
  constexpr int RS = row_stride(M, ROWS);
  constexpr int NT = ROWS * G;
  extern __shared__ float2 smem[];
  const int tid = threadIdx.x;
  const int row = tid / G, t = tid % G;
  const int x0 = blockIdx.x * ROWS;
  const size_t b = blockIdx.y;
  float2* s = smem + row * RS;
  {
    // Gather merged with the pre-processing of the inverse real transform: a thread loads the pairs
    // (k, M - k) of one row straight from T, forms Z[k] and Z[M - k] in registers and stores them
    // to the line buffers (no separate in-place pass: two shared-memory traversals fewer).  All of
    // a thread's gather loads are issued before the first one is consumed: the transposed reads
    // are the latency-critical start of this kernel.
    const float2* Tb = T + b * (size_t)M * Nx + (PAIRED ? 2 : 1) * (size_t)x0;
    auto tptr = [&](int ky, int r) -> const float2* {
      return PAIRED ? Tb + ((size_t)(ky >> 1) * Nx + r) * 2 + (ky & 1) : Tb + (size_t)ky * Nx + r;
    };
    auto pair_of = [&](int idx, int& r, int& k) {
      if constexpr (PAIRED) {
        r = (idx >> 1) % ROWS;
        k = 2 * (idx / (2 * ROWS)) + (idx & 1);
      } else {
        r = idx % ROWS;
        k = idx / ROWS;
      }
    };
    constexpr int NG = (ROWS * (M / 2) + NT - 1) / NT;  // pairs per thread
    float2 xa[NG], xb[NG];
#pragma unroll
    for (int n = 0; n < NG; ++n) {
      const int idx = tid + n * NT;
      int r, k;
      pair_of(idx, r, k);
      if (idx < ROWS * (M / 2)) {
        xa[n] = __ldg(tptr(k, r));
        xb[n] = __ldg(tptr(k == 0 ? M / 2 : M - k, r));
      }
    }
#pragma unroll
    for (int n = 0; n < NG; ++n) {
      const int idx = tid + n * NT;
      int r, k;
      pair_of(idx, r, k);
      if (idx < ROWS * (M / 2)) {
        float2* sr = smem + r * RS;
        if (k == 0) {
          sr[0] = make_float2(xa[n].x + xa[n].y, xa[n].x - xa[n].y);
          sr[PAD(M / 2)] = make_float2(2.f * xb[n].x, -2.f * xb[n].y);
        } else {
          const float2 xk = xa[n], xm = xb[n];
          const float2 A = make_float2(xk.x + xm.x, xk.y - xm.y);
          const float2 B = make_float2(xk.x - xm.x, xk.y + xm.y);
          const float2 WB = cmulc(B, __ldg(rtw + k));
          sr[PAD(k)] = make_float2(A.x + WB.x, A.y + WB.y);
          sr[PAD(M - k)] = make_float2(A.x - WB.x, -(A.y - WB.y));
        }
      }
    }
  }
  __syncthreads();
  using LSYNC = std::conditional_t<(ROWS <= 15), SyncLine<G>, SyncCta>;
  float2 v[E];
  fft_load_regs<P>(v, t, s);
  FftRun<P, +1, LSYNC>::run(v, t, s, tw, row);
  // v[e] = (q[2m], q[2m+1]) with m = t + G*e: a warp writes 32 consecutive float2 = 256 B
  float2* dst = reinterpret_cast<float2*>(q + (b * Nx + x0 + row) * (size_t)(2 * M));
#pragma unroll
  for (int e = 0; e < E; ++e) dst[t + G * e] = v[e];
}

// v' = u* - forward_difference(q)   (pressure.py:194-196), 4 columns per thread
// (qnext: the array holding row Nx of q -- the next rank's slab, or q itself on one GPU)
__global__ void correct2d_kernel(const float* __restrict__ us, const float* __restrict__ vs,
                                 const float* __restrict__ q, const float* __restrict__ qnext,
                                 float* __restrict__ uo, float* __restrict__ vo, int Nx, int Ny,
                                 float inv_hx, float inv_hy) {
  const size_t b = blockIdx.z;
  const int x = blockIdx.y;
  const int j = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (j >= Ny) return;
  const size_t row = (b * Nx + x) * (size_t)Ny;
  const float* qrow1 = (x == Nx - 1) ? qnext + (b * Nx) * (size_t)Ny : q + row + Ny;
  const float4 q0 = ldg4(q + row + j), q1 = ldg4(qrow1 + j);
  const float qr = __ldg(q + row + (j + 4 == Ny ? 0 : j + 4));
  const float4 u4 = ldg4(us + row + j), v4 = ldg4(vs + row + j);
  float4 ou, ov;
  ou.x = u4.x - (q1.x - q0.x) * inv_hx;
  ou.y = u4.y - (q1.y - q0.y) * inv_hx;
  ou.z = u4.z - (q1.z - q0.z) * inv_hx;
  ou.w = u4.w - (q1.w - q0.w) * inv_hx;
  ov.x = v4.x - (q0.y - q0.x) * inv_hy;
  ov.y = v4.y - (q0.z - q0.y) * inv_hy;
  ov.z = v4.z - (q0.w - q0.z) * inv_hy;
  ov.w = v4.w - (qr - q0.w) * inv_hy;
  stg4(uo + row + j, ou);
  stg4(vo + row + j, ov);
}

// rhs = divergence(v)   (finite_differences.py:136-143) -- used by cfd_project only
__global__ void divergence2d_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                    float* __restrict__ rhs, int Nx, int Ny, float inv_hx,
                                    float inv_hy) {
  const size_t b = blockIdx.z;
  const int x = blockIdx.y;
  const int j = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (j >= Ny) return;
  const int xm = x == 0 ? Nx - 1 : x - 1;
  const size_t off = (b * Nx + x) * (size_t)Ny + j, offm = (b * Nx + xm) * (size_t)Ny + j;
  const float4 u0 = ldg4(u + off), um = ldg4(u + offm), v0 = ldg4(v + off);
  const float vl = __ldg(v + (b * Nx + x) * (size_t)Ny + (j == 0 ? Ny - 1 : j - 1));
  float4 d;
  d.x = (u0.x - um.x) * inv_hx + (v0.x - vl) * inv_hy;
  d.y = (u0.y - um.y) * inv_hx + (v0.y - v0.x) * inv_hy;
  d.z = (u0.z - um.z) * inv_hx + (v0.z - v0.y) * inv_hy;
  d.w = (u0.w - um.w) * inv_hx + (v0.w - v0.z) * inv_hy;
  stg4(rhs + off, d);
}

template <int LM>
int launch_rfft_rows_t(cudaStream_t st, const float* rhs, float2* T, int batch, int Nx,
                       const float2* tw, const float2* rtw, int paired, int x_begin, int x_count) {
  constexpr int ROWS_MAX = rows_for(LM);
  using P = FftPlan<LM>;
  // ROWS must divide Nx
  auto go = [&](auto rows_c) -> int {
    constexpr int ROWS = decltype(rows_c)::value;
    constexpr size_t smem = (size_t)ROWS * row_stride(P::M, ROWS) * sizeof(float2);
    if (x_begin % ROWS || x_count % ROWS) return set_error_msg("internal: row block not a multiple of the rows per CTA");
    if (rows_lemax(LM) == 5 && LM >= 5) {
      using P5 = FftPlan<LM, 5, 4>;
      auto k = rfft_rows_kernel<LM, ROWS, 5, false>;
      if (int e = set_smem(k, smem)) return e;
      k<<<dim3(x_count / ROWS, batch), ROWS * P5::G, smem, st>>>(rhs, T, Nx, tw, rtw, x_begin);
    } else if (paired) {
      auto k = rfft_rows_kernel<LM, ROWS, 4, true>;
      if (int e = set_smem(k, smem)) return e;
      k<<<dim3(x_count / ROWS, batch), ROWS * P::G, smem, st>>>(rhs, T, Nx, tw, rtw, x_begin);
    } else {
      auto k = rfft_rows_kernel<LM, ROWS, 4, false>;
      if (int e = set_smem(k, smem)) return e;
      k<<<dim3(x_count / ROWS, batch), ROWS * P::G, smem, st>>>(rhs, T, Nx, tw, rtw, x_begin);
    }
    count_launch();
    CFD_CUDA_OK(cudaGetLastError());
    return 0;
  };
  if constexpr (ROWS_MAX >= 2) {
    if (rows_shift() == 1) return go(std::integral_constant<int, ROWS_MAX / 2>{});
  }
  if constexpr (ROWS_MAX >= 4) {
    if (rows_shift() == 2) return go(std::integral_constant<int, ROWS_MAX / 4>{});
  }
  if (Nx >= ROWS_MAX) return go(std::integral_constant<int, ROWS_MAX>{});
  if constexpr (ROWS_MAX > 16) {
    if (Nx >= 16) return go(std::integral_constant<int, 16>{});
  }
  return set_error_msg("grid axis 0 too small for the row FFT kernel (need >= 16)");
}

template <int LM, int LEMAX>
int launch_xlines_le(cudaStream_t st, const LinePeers& peers, const LinePeers& peers_w, int lnloc, size_t line_begin,
                     size_t nlines, int My, int split, const float2* tw, const double* lamx, const double* lamy,
                     const float* lamxf, const float* lamyf, int fastd, double cutoff, float norm, int paired,
                     const float* dtab) {
  if (dtab) fastd = 0;  // the table mode lives in the variant that tests every eigenvalue
  using P = FftPlan<LM, LEMAX, 4>;
  constexpr int LINES = (P::G >= 256) ? 1 : (256 / P::G > CFD_XL_LINES_MAX ? CFD_XL_LINES_MAX : 256 / P::G);
  // two exchange buffers (one barrier per pass) whenever both fit beside a second CTA's
  constexpr bool DB = LEMAX == 4 && (size_t)LINES * row_stride(P::M, 16) * sizeof(float2) <= 72 * 1024;
  constexpr size_t smem = (size_t)(DB ? 2 : 1) * LINES * row_stride(P::M, 16) * sizeof(float2);
  if (nlines % LINES || (!split && line_begin % LINES)) return set_error_msg("internal: line count not divisible");
  auto go = [&](auto k) -> int {
    if (int e = set_smem(k, smem)) return e;
    k<<<(unsigned)(nlines / LINES), LINES * P::G, smem, st>>>(peers, peers_w, lnloc, line_begin, My, tw, lamx, lamy,
                                                            lamxf, lamyf, cutoff, norm, dtab);
    return 0;
  };
  int e;
  if constexpr (LM == 14) {  // split mode exists for the half-lines of 32768-point lines only
    if (paired && !split) return set_error_msg("internal: the paired layout is not built for 16384-point lines");
    if (split)
      e = fastd ? go(xlines_kernel<LM, LINES, true, LEMAX, true, DB, false>) : go(xlines_kernel<LM, LINES, false, LEMAX, true, DB, false>);
    else
      e = fastd ? go(xlines_kernel<LM, LINES, true, LEMAX, false, DB, false>) : go(xlines_kernel<LM, LINES, false, LEMAX, false, DB, false>);
  } else {
    if (split) return set_error_msg("internal: split x lines need 16384-point transforms");
    bool bal = false;
    if constexpr (LM == 13 && LEMAX == 5) bal = xlines_balanced(LM);
    if (paired) {
      if constexpr (LM == 12 || LM == 13) {  // the only lengths the plan selects the paired layout for
        if constexpr (LM == 13 && LEMAX == 5) {
          if (bal) {
            e = fastd ? go(xlines_kernel<LM, LINES, true, LEMAX, false, DB, true, true>)
                      : go(xlines_kernel<LM, LINES, false, LEMAX, false, DB, true, true>);
            if (e) return e;
            count_launch();
            CFD_CUDA_OK(cudaGetLastError());
            return 0;
          }
        }
        e = fastd ? go(xlines_kernel<LM, LINES, true, LEMAX, false, DB, true>) : go(xlines_kernel<LM, LINES, false, LEMAX, false, DB, true>);
      } else {
        return set_error_msg("internal: the paired layout is built for 4096/8192-point x lines only");
      }
    } else {
      if constexpr (LM == 13 && LEMAX == 5) {
        if (bal) {
          e = fastd ? go(xlines_kernel<LM, LINES, true, LEMAX, false, DB, false, true>)
                    : go(xlines_kernel<LM, LINES, false, LEMAX, false, DB, false, true>);
          if (e) return e;
          count_launch();
          CFD_CUDA_OK(cudaGetLastError());
          return 0;
        }
      }
      e = fastd ? go(xlines_kernel<LM, LINES, true, LEMAX, false, DB, false>) : go(xlines_kernel<LM, LINES, false, LEMAX, false, DB, false>);
    }
  }
  if (e) return e;
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int LM>
int launch_xlines_t(cudaStream_t st, const LinePeers& peers, const LinePeers& peers_w, int lnloc, size_t line_begin,
                    size_t nlines, int My, int split, const float2* tw, const double* lamx, const double* lamy,
                    const float* lamxf, const float* lamyf, int fastd, double cutoff, float norm, int paired,
                    const float* dtab) {
  if (xlines_lemax(LM) == 5)
    return launch_xlines_le<LM, 5>(st, peers, peers_w, lnloc, line_begin, nlines, My, split, tw, lamx, lamy, lamxf,
                                   lamyf, fastd, cutoff, norm, paired, dtab);
  return launch_xlines_le<LM, 4>(st, peers, peers_w, lnloc, line_begin, nlines, My, split, tw, lamx, lamy, lamxf,
                                 lamyf, fastd, cutoff, norm, paired, dtab);
}

template <int LM>
int launch_irfft_rows_t(cudaStream_t st, const float2* T, float* q, int batch, int Nx,
                        const float2* tw, const float2* rtw, int paired) {
  constexpr int ROWS_MAX = rows_for(LM);
  using P = FftPlan<LM>;
  auto go = [&](auto rows_c) -> int {
    constexpr int ROWS = decltype(rows_c)::value;
    constexpr size_t smem = (size_t)ROWS * row_stride(P::M, ROWS) * sizeof(float2);
    if (rows_lemax(LM) == 5 && LM >= 5) {
      using P5 = FftPlan<LM, 5, 4>;
      auto k = irfft_rows_kernel<LM, ROWS, 5, false>;
      if (int e = set_smem(k, smem)) return e;
      k<<<dim3(Nx / ROWS, batch), ROWS * P5::G, smem, st>>>(T, q, Nx, tw, rtw);
    } else if (paired) {
      auto k = irfft_rows_kernel<LM, ROWS, 4, true>;
      if (int e = set_smem(k, smem)) return e;
      k<<<dim3(Nx / ROWS, batch), ROWS * P::G, smem, st>>>(T, q, Nx, tw, rtw);
    } else {
      auto k = irfft_rows_kernel<LM, ROWS, 4, false>;
      if (int e = set_smem(k, smem)) return e;
      k<<<dim3(Nx / ROWS, batch), ROWS * P::G, smem, st>>>(T, q, Nx, tw, rtw);
    }
    count_launch();
    CFD_CUDA_OK(cudaGetLastError());
    return 0;
  };
  if constexpr (ROWS_MAX >= 2) {
    if (rows_shift() == 1) return go(std::integral_constant<int, ROWS_MAX / 2>{});
  }
  if constexpr (ROWS_MAX >= 4) {
    if (rows_shift() == 2) return go(std::integral_constant<int, ROWS_MAX / 4>{});
  }
  if (Nx >= ROWS_MAX) return go(std::integral_constant<int, ROWS_MAX>{});
  if constexpr (ROWS_MAX > 16) {
    if (Nx >= 16) return go(std::integral_constant<int, 16>{});
  }
  return set_error_msg("grid axis 0 too small for the row FFT kernel (need >= 16)");
}

}  // namespace

// lm_row = log2(Ny / 2)
// x_begin / x_count: the block of rows to transform (x_count < 0: all Nx rows)
int launch_rfft_rows_block(cudaStream_t st, int lm_row, const float* rhs, float2* T, int batch, int Nx,
                           const float2* tw, const float2* rtw, int paired, int x_begin, int x_count) {
  if (x_count < 0) {
    x_begin = 0;
    x_count = Nx;
  }
  CFD_DISPATCH_LM(lm_row, 4, 14,
                  return launch_rfft_rows_t<LM_>(st, rhs, T, batch, Nx, tw, rtw, paired, x_begin, x_count));
  return 0;
}
int launch_rfft_rows(cudaStream_t st, int lm_row, const float* rhs, float2* T, int batch, int Nx,
                     const float2* tw, const float2* rtw, int paired) {
  return launch_rfft_rows_block(st, lm_row, rhs, T, batch, Nx, tw, rtw, paired, 0, -1);
}
// lm_x = log2(Nx)
int launch_xlines15_cluster(cudaStream_t st, const LinePeers& peers, const LinePeers& peers_w, int lnloc,
                            size_t line_begin, size_t nlines, int My, const float2* tw, const double* lamx,
                            const double* lamy, const float* lamxf, const float* lamyf, int fastd,
                            double cutoff, float norm, const float2* wbig, int paired, const float* dtab);
bool x15_cluster(int paired, int several_gpus);

// `scratch` / `wbig` are only needed for lm_x == 15 (32768-point lines): scratch holds
// nlines * 32768 float2, wbig[m] = exp(-2 pi i m / 32768), m < 16384.  With side streams the split
// path runs in chunks of lines on those streams, so that the NVLink reads of one chunk (split), the
// transforms of another (x lines, local scratch) and the NVLink writes of a third (merge) overlap.
int launch_xlines_peers(cudaStream_t st, int lm_x, const LinePeers& peers, int lnloc,
                        size_t line_begin, size_t nlines, int My, const float2* tw,
                        const double* lamx, const double* lamy, const float* lamxf,
                        const float* lamyf, int fastd, double cutoff, float norm, float2* scratch,
                        const float2* wbig, const SideStreams* side, int paired, const float* dtab,
                        const LinePeers* peers_out) {
  const LinePeers& peers_w = peers_out ? *peers_out : peers;  // results go back in place by default
  // (pull mode -- the line is READ from the peers -- keeps the scratch path: the cluster kernel
  // loads every element twice, which would double the NVLink reads)
  if (lm_x == 15 && !(lnloc != lm_x && peers_out == nullptr) && x15_cluster(paired, lnloc != lm_x)) {
    if (!wbig) return set_error_msg("internal: 32768-point lines need the w table");
    return launch_xlines15_cluster(st, peers, peers_w, lnloc, line_begin, nlines, My, tw, lamx, lamy, lamxf, lamyf,
                                   fastd, cutoff, norm, wbig, paired, dtab);
  }
  if (lm_x == 15) {
    if (!scratch || !wbig) return set_error_msg("internal: 32768-point lines need the split scratch");
    const int half = 1 << 14;
    int nchunks = 1;
    if (side && side->n > 0) {
      nchunks = 8;
      while (nchunks > 1 && (nlines % nchunks || nlines / nchunks < 32)) nchunks /= 2;
    }
    const size_t chunk = nlines / nchunks;
    if (nchunks > 1) CFD_CUDA_OK(cudaEventRecord(side->start, st));
    for (int c = 0; c < nchunks; ++c) {
      cudaStream_t s = nchunks > 1 ? side->s[c % side->n] : st;
      if (nchunks > 1 && c < side->n) CFD_CUDA_OK(cudaStreamWaitEvent(s, side->start, 0));
      float2* sc = scratch + (size_t)c * chunk * (size_t)(2 * half);
      const size_t lb = line_begin + (size_t)c * chunk;
      dim3 grid(half / 256, (unsigned)chunk);
      dim3 pgrid(half / 256, (unsigned)(chunk / 2));
      if (paired) {
        if ((lb | chunk) & 1) return set_error_msg("internal: paired lines need even line ranges");
        split_pairs_kernel<<<pgrid, 256, 0, s>>>(peers, lnloc, lb, half, sc, wbig);
      } else {
        split_lines_kernel<<<grid, 256, 0, s>>>(peers, lnloc, lb, half, sc, wbig);
      }
      count_launch();
      CFD_CUDA_OK(cudaGetLastError());
      LinePeers local;
      for (int i = 0; i < CFD_MAX_PEERS; ++i) local.p[i] = sc;
      if (int e = launch_xlines_t<14>(s, local, local, 14, lb, 2 * chunk, My, 1, tw, lamx, lamy, lamxf, lamyf,
                                      fastd, cutoff, norm, 0, dtab))
        return e;
      if (paired)
        merge_pairs_kernel<<<pgrid, 256, 0, s>>>(peers_w, lnloc, lb, half, sc, wbig);
      else
        merge_lines_kernel<<<grid, 256, 0, s>>>(peers_w, lnloc, lb, half, sc, wbig);
      count_launch();
      CFD_CUDA_OK(cudaGetLastError());
    }
    if (nchunks > 1) {
      for (int i = 0; i < side->n; ++i) {
        CFD_CUDA_OK(cudaEventRecord(side->done[i], side->s[i]));
        CFD_CUDA_OK(cudaStreamWaitEvent(st, side->done[i], 0));
      }
    }
    return 0;
  }
  CFD_DISPATCH_LM(lm_x, 4, 14,
                  return launch_xlines_t<LM_>(st, peers, peers_w, lnloc, line_begin, nlines, My, 0, tw, lamx, lamy,
                                              lamxf, lamyf, fastd, cutoff, norm, paired, dtab));
  return 0;
}
int launch_xlines(cudaStream_t st, int lm_x, float2* T, int batch, int My, const float2* tw,
                  const double* lamx, const double* lamy, const float* lamxf, const float* lamyf,
                  int fastd, double cutoff, float norm, float2* scratch, const float2* wbig,
                  const SideStreams* side, int paired, const float* dtab) {
  LinePeers peers;
  for (int i = 0; i < CFD_MAX_PEERS; ++i) peers.p[i] = T;
  return launch_xlines_peers(st, lm_x, peers, lm_x, 0, (size_t)batch * My, My, tw, lamx, lamy, lamxf,
                             lamyf, fastd, cutoff, norm, scratch, wbig, side, paired, dtab, nullptr);
}
int launch_divergence_generic(cudaStream_t st, const float* u, const float* v, const float* w, float* rhs,
                              int batch, int N0, int N1, int N2, float ih0, float ih1, float ih2);
int launch_correct_generic(cudaStream_t st, const float* us, const float* vs, const float* ws, const float* q,
                           float* uo, float* vo, float* wo, int batch, int N0, int N1, int N2, float ih0,
                           float ih1, float ih2);

int launch_divergence_2d(cudaStream_t st, const float* u, const float* v, float* rhs, int batch,
                         int Nx, int Ny, float inv_hx, float inv_hy) {
  if (Ny % 4) return launch_divergence_generic(st, u, v, nullptr, rhs, batch, Nx, Ny, 1, inv_hx, inv_hy, 0.f);
  const int threads = 128;
  dim3 grid((Ny / 4 + threads - 1) / threads, Nx, batch);
  divergence2d_kernel<<<grid, threads, 0, st>>>(u, v, rhs, Nx, Ny, inv_hx, inv_hy);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfd

namespace cfd {
int launch_irfft_rows(cudaStream_t st, int lm_row, const float2* T, float* q, int batch, int Nx,
                      const float2* tw, const float2* rtw, int paired) {
  CFD_DISPATCH_LM(lm_row, 4, 14, return launch_irfft_rows_t<LM_>(st, T, q, batch, Nx, tw, rtw, paired));
  return 0;
}
int launch_correct_2d(cudaStream_t st, const float* us, const float* vs, const float* q,
                      const float* qnext, float* uo, float* vo, int batch, int Nx, int Ny,
                      float inv_hx, float inv_hy) {
  if (Ny % 4) {
    if (qnext && qnext != q) return set_error_msg("internal: slab grids need rows of a multiple of 4 columns");
    return launch_correct_generic(st, us, vs, nullptr, q, uo, vo, nullptr, batch, Nx, Ny, 1, inv_hx, inv_hy, 0.f);
  }
  const int threads = 128;
  dim3 grid((Ny / 4 + threads - 1) / threads, Nx, batch);
  correct2d_kernel<<<grid, threads, 0, st>>>(us, vs, q, qnext ? qnext : q, uo, vo, Nx, Ny, inv_hx,
                                             inv_hy);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}
}  // namespace cfd
