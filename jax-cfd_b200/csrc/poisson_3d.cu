// 3-D fast-diagonalisation pressure solve  q = irfftn( D * rfftn(rhs) )  as five HBM sweeps
// (fast_diagonalization.py:199-225; pressure.py:115-157), each a batched shared-memory line FFT:
//
//   Z   rfft along z (axis 2, contiguous):   rhs[x][y][z]      -> T1[kz][x][y]   (kz = 0..N2/2)
//   Y   fft  along y, transposed store:      T1[kz][x][.]      -> T2[kz][ky][x]
//   X   fft along x * D * ifft along x:      T2[kz][ky][.]     in place
//   Yi  gather, ifft along y:                T2[kz][.][x]      -> T1[kz][x][y]
//   Zi  gather, irfft along z:               T1[.][x][y]       -> q[x][y][z]
//
// The half spectrum is stored UNPACKED here (N2/2 + 1 planes): the point-reflection partner needed
// to split a packed plane lives in a different line, and one extra plane in N2/2 is negligible.
// Scaling: Z stores 2X and Zi consumes 2Z, all passes unnormalised -> D carries 1/(2 N0 N1 N2).
#include <stdlib.h>

#include "common.cuh"
#include "fft_rows.cuh"
#include "fft_smem.cuh"

namespace cfd {

namespace {

// ---- Z: real rows -> T1[b][kz][row],  kz = 0..M,  row = x * N1 + y ---------------------------------
template <int LM, int ROWS>
__global__ void __launch_bounds__(ROWS * FftPlan<LM>::G)
rfft_rows3_kernel(const float* __restrict__ rhs, float2* __restrict__ T, int NR,
                  const float2* __restrict__ tw, const float2* __restrict__ rtw) {
  using P = FftPlan<LM>;
  constexpr int M = P::M, G = P::G, E = P::E;
  constexpr int RS = row_stride(M, ROWS);
  extern __shared__ float2 smem[];
  const int tid = threadIdx.x;
  const int row = tid / G, t = tid % G;
  const int r0 = blockIdx.x * ROWS;
  const size_t b = blockIdx.y;
  float2* s = smem + row * RS;
  float2 v[E];
  {
    const float2* src = reinterpret_cast<const float2*>(rhs + (b * NR + r0 + row) * (size_t)(2 * M));
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = __ldg(src + t + G * e);
  }
  FftRun<P, -1>::run(v, t, s, tw);
  __syncthreads();
#pragma unroll
  for (int e = 0; e < E; ++e) s[PAD(t + G * e)] = v[e];
  __syncthreads();
  for (int k = t; k <= M / 2; k += G) {
    if (k == 0) {
      const float2 z = s[0];
      s[0] = make_float2(2.f * (z.x + z.y), 0.f);
      s[PAD(M)] = make_float2(2.f * (z.x - z.y), 0.f);
    } else if (k == M / 2) {
      const float2 z = s[PAD(k)];
      s[PAD(k)] = make_float2(2.f * z.x, -2.f * z.y);
    } else {
      const float2 zk = s[PAD(k)], zm = s[PAD(M - k)];
      const float2 A = make_float2(zk.x + zm.x, zk.y - zm.y);
      const float2 B = make_float2(zk.x - zm.x, zk.y + zm.y);
      const float2 WB = cmul(__ldg(rtw + k), B);
      s[PAD(k)] = make_float2(A.x + WB.x, A.y + WB.y);
      s[PAD(M - k)] = make_float2(A.x - WB.x, -(A.y - WB.y));
    }
  }
  __syncthreads();
  float2* Tb = T + b * (size_t)(M + 1) * NR + r0;
  constexpr int NT = ROWS * G;
#pragma unroll 4
  for (int idx = tid; idx < ROWS * (M + 1); idx += NT) {
    const int r = idx % ROWS, kz = idx / ROWS;
    Tb[(size_t)kz * NR + r] = smem[r * RS + PAD(kz)];
  }
}

// ---- Zi: T1[b][kz][row] -> q rows -------------------------------------------------------------------
template <int LM, int ROWS>
__global__ void __launch_bounds__(ROWS * FftPlan<LM>::G)
irfft_rows3_kernel(const float2* __restrict__ T, float* __restrict__ q, int NR,
                   const float2* __restrict__ tw, const float2* __restrict__ rtw) {
  using P = FftPlan<LM>;
  constexpr int M = P::M, G = P::G, E = P::E;
  constexpr int RS = row_stride(M, ROWS);
  constexpr int NT = ROWS * G;
  extern __shared__ float2 smem[];
  const int tid = threadIdx.x;
  const int row = tid / G, t = tid % G;
  const int r0 = blockIdx.x * ROWS;
  const size_t b = blockIdx.y;
  float2* s = smem + row * RS;
  {
    const float2* Tb = T + b * (size_t)(M + 1) * NR + r0;
#pragma unroll 4
    for (int idx = tid; idx < ROWS * (M + 1); idx += NT) {
      const int r = idx % ROWS, kz = idx / ROWS;
      smem[r * RS + PAD(kz)] = __ldg(Tb + (size_t)kz * NR + r);
    }
  }
  __syncthreads();
  for (int k = t; k <= M / 2; k += G) {
    if (k == 0) {
      const float a = s[0].x, bb = s[PAD(M)].x;  // X[0], X[M] (real)
      s[0] = make_float2(a + bb, a - bb);
    } else if (k == M / 2) {
      const float2 x = s[PAD(k)];
      s[PAD(k)] = make_float2(2.f * x.x, -2.f * x.y);
    } else {
      const float2 xk = s[PAD(k)], xm = s[PAD(M - k)];
      const float2 A = make_float2(xk.x + xm.x, xk.y - xm.y);
      const float2 B = make_float2(xk.x - xm.x, xk.y + xm.y);
      const float2 WB = cmulc(B, __ldg(rtw + k));
      s[PAD(k)] = make_float2(A.x + WB.x, A.y + WB.y);
      s[PAD(M - k)] = make_float2(A.x - WB.x, -(A.y - WB.y));
    }
  }
  __syncthreads();
  float2 v[E];
  fft_load_regs<P>(v, t, s);
  FftRun<P, +1>::run(v, t, s, tw);
  float2* dst = reinterpret_cast<float2*>(q + (b * NR + r0 + row) * (size_t)(2 * M));
#pragma unroll
  for (int e = 0; e < E; ++e) dst[t + G * e] = v[e];
}

// ---- Y: contiguous complex lines A[plane][l][0..M) -> FFT -> B[plane][k][l] (transposed) --------------
template <int LM, int ROWS>
__global__ void __launch_bounds__(ROWS * FftPlan<LM>::G)
cfft_lines_scatter_kernel(const float2* __restrict__ A, float2* __restrict__ B, int NL,
                          const float2* __restrict__ tw) {
  using P = FftPlan<LM>;
  constexpr int M = P::M, G = P::G, E = P::E;
  constexpr int RS = row_stride(M, ROWS);
  constexpr int NT = ROWS * G;
  extern __shared__ float2 smem[];
  const int tid = threadIdx.x;
  const int row = tid / G, t = tid % G;
  const int l0 = blockIdx.x * ROWS;
  const size_t plane = blockIdx.y;
  float2* s = smem + row * RS;
  float2 v[E];
  {
    const float2* src = A + (plane * NL + l0 + row) * (size_t)M;
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = __ldg(src + t + G * e);
  }
  FftRun<P, -1>::run(v, t, s, tw);
  __syncthreads();
#pragma unroll
  for (int e = 0; e < E; ++e) s[PAD(t + G * e)] = v[e];
  __syncthreads();
  float2* Bp = B + plane * (size_t)M * NL + l0;
#pragma unroll 4
  for (int idx = tid; idx < ROWS * M; idx += NT) {
    const int r = idx % ROWS, k = idx / ROWS;
    Bp[(size_t)k * NL + r] = smem[r * RS + PAD(k)];
  }
}

// ---- Yi: gather B[plane][k][l] -> inverse FFT -> A[plane][l][0..M) -------------------------------------
template <int LM, int ROWS>
__global__ void __launch_bounds__(ROWS * FftPlan<LM>::G)
cfft_lines_gather_kernel(const float2* __restrict__ B, float2* __restrict__ A, int NL,
                         const float2* __restrict__ tw) {
  using P = FftPlan<LM>;
  constexpr int M = P::M, G = P::G, E = P::E;
  constexpr int RS = row_stride(M, ROWS);
  constexpr int NT = ROWS * G;
  extern __shared__ float2 smem[];
  const int tid = threadIdx.x;
  const int row = tid / G, t = tid % G;
  const int l0 = blockIdx.x * ROWS;
  const size_t plane = blockIdx.y;
  float2* s = smem + row * RS;
  {
    const float2* Bp = B + plane * (size_t)M * NL + l0;
#pragma unroll 4
    for (int idx = tid; idx < ROWS * M; idx += NT) {
      const int r = idx % ROWS, k = idx / ROWS;
      smem[r * RS + PAD(k)] = __ldg(Bp + (size_t)k * NL + r);
    }
  }
  __syncthreads();
  float2 v[E];
  fft_load_regs<P>(v, t, s);
  FftRun<P, +1>::run(v, t, s, tw);
  float2* dst = A + (plane * NL + l0 + row) * (size_t)M;
#pragma unroll
  for (int e = 0; e < E; ++e) dst[t + G * e] = v[e];
}

// ---- X: lines B[line][0..M) with line = (b * (Mz + 1) + kz) * N1 + ky : fwd * D * inv in place ---------
// (no minimum-blocks hint: measured at 512^3 with 512-point lines, 0.307 ms as is -- 127 registers,
// two CTAs per SM -- against 0.332 / 0.396 ms capped at 80 / 64 registers, which spill)
template <int LM, int LINES, bool FASTD>
__global__ void __launch_bounds__(LINES * FftPlan<LM>::G)
xlines3_kernel(LinePeers peers, int lnloc, size_t line_begin, int N1, int NZP, const float2* __restrict__ tw,
               const double* __restrict__ lamx, const double* __restrict__ lamy,
               const double* __restrict__ lamz, const float* __restrict__ lamxf,
               const float* __restrict__ lamyf, const float* __restrict__ lamzf, double cutoff,
               float norm, const float* __restrict__ dtab) {
  using P = FftPlan<LM>;
  constexpr int M = P::M, G = P::G, E = P::E;
  constexpr int RS = row_stride(M, 16);
  extern __shared__ float2 smem[];
  const int tid = threadIdx.x;
  const int ln = tid / G, t = tid % G;
  const size_t line = line_begin + (size_t)blockIdx.x * LINES + ln;
  const int ky = (int)(line % N1);
  const int kz = (int)((line / N1) % NZP);
  float2* s = smem + ln * RS;
  // One GPU (lnloc == LM): the line is contiguous.  Slab decomposition: element x of the line lives
  // in rank x >> lnloc, at [line][x & (Nloc - 1)] of its slab spectrum -- the all-to-all transposes
  // of the distributed FFT are this kernel's loads and stores (peer table in shared memory: a
  // dynamically indexed kernel parameter would live in local memory).
  __shared__ float2* s_peer[CFD_MAX_PEERS];
  const bool single = (lnloc == LM);
  if (!single) {
    if (tid < CFD_MAX_PEERS) s_peer[tid] = peers.p[tid];
    __syncthreads();
  }
  const size_t loff = line << lnloc;
  const int nloc_mask = (1 << lnloc) - 1;
  float2* const Tl = peers.p[0] + loff;
  auto elem = [&](int x) -> float2* {
    return single ? Tl + x : s_peer[x >> lnloc] + loff + (x & nloc_mask);
  };
  float2 v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) v[e] = *elem(t + G * e);
  FftRun<P, -1>::run(v, t, s, tw);
  if (FASTD) {
    const float lyz = __ldg(lamyf + ky) + __ldg(lamzf + kz);
    const bool mean_line = (ky == 0) && (kz == 0);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int kx = t + G * e;
      float d = norm * fast_rcp(__ldg(lamxf + kx) + lyz);
      if (mean_line && kx == 0) d = 0.f;
      v[e].x *= d;
      v[e].y *= d;
    }
  } else if (dtab) {
    // caller-supplied real diagonal in line layout dtab[(kz * N1 + ky) * M + kx] (cfd_transform)
    const float* trow = dtab + ((size_t)kz * N1 + ky) * M;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float d = norm * __ldg(trow + t + G * e);
      v[e].x *= d;
      v[e].y *= d;
    }
  } else {
    const double lyz = __ldg(lamy + ky) + __ldg(lamz + kz);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const double lam = __ldg(lamx + t + G * e) + lyz;
      const float d = (fabs(lam) > cutoff) ? norm * fast_rcp((float)lam) : 0.f;
      v[e].x *= d;
      v[e].y *= d;
    }
  }
  FftRun<P, +1>::run(v, t, s, tw);
#pragma unroll
  for (int e = 0; e < E; ++e) *elem(t + G * e) = v[e];
}

// ---- elementwise 3-D helpers -------------------------------------------------------------------------
// rhs = divergence(v)  (finite_differences.py:136-143)
// u_below: plane -1 of u (the previous rank's last plane on a slab-decomposed grid, multi_gpu_3d.cu;
// u + (N0 - 1) * plane -- the periodic wrap -- on one GPU)
__global__ void divergence3d_kernel(const float* __restrict__ u, const float* __restrict__ u_below,
                                    const float* __restrict__ v,
                                    const float* __restrict__ w, float* __restrict__ rhs, int N0,
                                    int N1, int N2, float ihx, float ihy, float ihz) {
  const size_t b = blockIdx.y;
  const int nchunk = (N2 / 4 + blockDim.x - 1) / blockDim.x;
  const int xy = blockIdx.x / nchunk;
  const int x = xy / N1, y = xy % N1;
  const int k = 4 * ((blockIdx.x % nchunk) * blockDim.x + threadIdx.x);
  if (k >= N2) return;
  const size_t plane = (size_t)N1 * N2, base = b * (size_t)N0 * plane;
  const int ym = y == 0 ? N1 - 1 : y - 1;
  const size_t off = base + x * plane + (size_t)y * N2 + k;
  const float4 u0 = ldg4(u + off);
  const float4 um = ldg4((x == 0 ? u_below + base : u + base + (x - 1) * plane) + (size_t)y * N2 + k);
  const float4 v0 = ldg4(v + off), vm = ldg4(v + base + x * plane + (size_t)ym * N2 + k);
  const float4 w0 = ldg4(w + off);
  const float wl = __ldg(w + base + x * plane + (size_t)y * N2 + (k == 0 ? N2 - 1 : k - 1));
  float4 d;
  d.x = (u0.x - um.x) * ihx + (v0.x - vm.x) * ihy + (w0.x - wl) * ihz;
  d.y = (u0.y - um.y) * ihx + (v0.y - vm.y) * ihy + (w0.y - w0.x) * ihz;
  d.z = (u0.z - um.z) * ihx + (v0.z - vm.z) * ihy + (w0.z - w0.y) * ihz;
  d.w = (u0.w - um.w) * ihx + (v0.w - vm.w) * ihy + (w0.w - w0.z) * ihz;
  stg4(rhs + off, d);
}

// v' = u* - forward_difference(q)   (pressure.py:194-196)
// q_above: plane N0 of q (the next rank's first plane on a slab-decomposed grid; q itself on one GPU)
__global__ void correct3d_kernel(const float* __restrict__ us, const float* __restrict__ vs,
                                 const float* __restrict__ ws, const float* __restrict__ q,
                                 const float* __restrict__ q_above, float* __restrict__ uo, float* __restrict__ vo,
                                 float* __restrict__ wo, int N0, int N1, int N2, float ihx,
                                 float ihy, float ihz) {
  const size_t b = blockIdx.y;
  const int nchunk = (N2 / 4 + blockDim.x - 1) / blockDim.x;
  const int xy = blockIdx.x / nchunk;
  const int x = xy / N1, y = xy % N1;
  const int k = 4 * ((blockIdx.x % nchunk) * blockDim.x + threadIdx.x);
  if (k >= N2) return;
  const size_t plane = (size_t)N1 * N2, base = b * (size_t)N0 * plane;
  const int yp = y == N1 - 1 ? 0 : y + 1;
  const size_t off = base + x * plane + (size_t)y * N2 + k;
  const float4 q0 = ldg4(q + off);
  const float4 qx = ldg4((x == N0 - 1 ? q_above + base : q + base + (x + 1) * plane) + (size_t)y * N2 + k);
  const float4 qy = ldg4(q + base + x * plane + (size_t)yp * N2 + k);
  const float qz = __ldg(q + base + x * plane + (size_t)y * N2 + (k + 4 == N2 ? 0 : k + 4));
  const float4 a = ldg4(us + off), bq = ldg4(vs + off), c = ldg4(ws + off);
  float4 ou, ov, ow;
  ou.x = a.x - (qx.x - q0.x) * ihx;
  ou.y = a.y - (qx.y - q0.y) * ihx;
  ou.z = a.z - (qx.z - q0.z) * ihx;
  ou.w = a.w - (qx.w - q0.w) * ihx;
  ov.x = bq.x - (qy.x - q0.x) * ihy;
  ov.y = bq.y - (qy.y - q0.y) * ihy;
  ov.z = bq.z - (qy.z - q0.z) * ihy;
  ov.w = bq.w - (qy.w - q0.w) * ihy;
  ow.x = c.x - (q0.y - q0.x) * ihz;
  ow.y = c.y - (q0.z - q0.y) * ihz;
  ow.z = c.z - (q0.w - q0.z) * ihz;
  ow.w = c.w - (qz - q0.w) * ihz;
  stg4(uo + off, ou);
  stg4(vo + off, ov);
  stg4(wo + off, ow);
}

// Rows per CTA of the 3-D row / line kernels are capped at CFD_FFT3_ROWS = 8 | 16 | 32.  Default 16:
// with rows_for's maximum of 32 a CTA of 512-point lines is 1024 threads and 139 KB, one per SM;
// 16 rows (512 threads, 70 KB) let three independent CTAs share an SM and overlap each other's
// load / exchange / store phases -- measured at 512^3: the four sweeps 1.12 -> 0.95 ms (inverse y
// sweep alone: 0.326 / 0.287 / 0.249 ms with 32 / 16 / 8 rows).
static int rows3_cap(int gather) {
  static const int v = [] {
    const char* e = getenv("CFD_FFT3_ROWS");
    return e ? atoi(e) : 0;
  }();
  if (v > 0) return v;
  return gather ? 8 : 16;  // the transposed gather of the inverse y sweep likes even smaller CTAs
}

// generic launcher: KERNEL<LM, ROWS> over (count / ROWS, planes) CTAs, ROWS from rows_for(LM)
#define CFD_ROWS_LAUNCH(KERNEL, CAP, count, planes, ...)                                              \
  do {                                                                                           \
    constexpr int ROWS_MAX = rows_for(LM);                                                       \
    using P = FftPlan<LM>;                                                                       \
    auto go = [&](auto rows_c) -> int {                                                          \
      constexpr int ROWS = decltype(rows_c)::value;                                              \
      constexpr size_t smem = (size_t)ROWS * row_stride(P::M, ROWS) * sizeof(float2);            \
      auto k = KERNEL<LM, ROWS>;                                                                 \
      if (int e = set_smem(k, smem)) return e;                                                   \
      k<<<dim3((unsigned)((count) / ROWS), (unsigned)(planes)), ROWS * P::G, smem, st>>>(__VA_ARGS__); \
      count_launch();                                                                            \
      CFD_CUDA_OK(cudaGetLastError());                                                           \
      return 0;                                                                                  \
    };                                                                                           \
    if ((count) % ROWS_MAX == 0 && (CAP) >= ROWS_MAX) return go(std::integral_constant<int, ROWS_MAX>{}); \
    if constexpr (ROWS_MAX > 8) {                                                                \
      if ((count) % 8 == 0 && (CAP) <= 8) return go(std::integral_constant<int, 8>{});     \
    }                                                                                            \
    if constexpr (ROWS_MAX > 16) {                                                               \
      if ((count) % 16 == 0) return go(std::integral_constant<int, 16>{});                       \
    }                                                                                            \
    return set_error_msg("3-D grid axis too small for the line FFT kernels (need >= 16)");       \
  } while (0)

template <int LM>
int launch_rfft_rows3_t(cudaStream_t st, const float* rhs, float2* T, int batch, int NR,
                        const float2* tw, const float2* rtw) {
  CFD_ROWS_LAUNCH(rfft_rows3_kernel, rows3_cap(0), NR, batch, rhs, T, NR, tw, rtw);
}
template <int LM>
int launch_irfft_rows3_t(cudaStream_t st, const float2* T, float* q, int batch, int NR,
                         const float2* tw, const float2* rtw) {
  CFD_ROWS_LAUNCH(irfft_rows3_kernel, rows3_cap(0), NR, batch, T, q, NR, tw, rtw);
}
template <int LM>
int launch_lines_scatter_t(cudaStream_t st, const float2* A, float2* B, int planes, int NL,
                           const float2* tw) {
  CFD_ROWS_LAUNCH(cfft_lines_scatter_kernel, rows3_cap(0), NL, planes, A, B, NL, tw);
}
template <int LM>
int launch_lines_gather_t(cudaStream_t st, const float2* B, float2* A, int planes, int NL,
                          const float2* tw) {
  CFD_ROWS_LAUNCH(cfft_lines_gather_kernel, rows3_cap(1), NL, planes, B, A, NL, tw);
}

template <int LM>
int launch_xlines3_t(cudaStream_t st, const LinePeers& peers, int lnloc, size_t line_begin, size_t nlines, int N1,
                     int NZP, const float2* tw,
                     const double* const* lam, const float* const* lamf, int fastd, double cutoff,
                     float norm, const float* dtab) {
  if (dtab) fastd = 0;
  constexpr int LINES = lines_for(LM);
  using P = FftPlan<LM>;
  constexpr size_t smem = (size_t)LINES * row_stride(P::M, 16) * sizeof(float2);
  if (nlines % LINES || line_begin % LINES) return set_error_msg("3-D line count not divisible by the lines per CTA");
  if (fastd) {
    auto k = xlines3_kernel<LM, LINES, true>;
    if (int e = set_smem(k, smem)) return e;
    k<<<(unsigned)(nlines / LINES), LINES * P::G, smem, st>>>(peers, lnloc, line_begin, N1, NZP, tw, lam[0], lam[1],
                                                            lam[2], lamf[0], lamf[1], lamf[2], cutoff, norm, dtab);
  } else {
    auto k = xlines3_kernel<LM, LINES, false>;
    if (int e = set_smem(k, smem)) return e;
    k<<<(unsigned)(nlines / LINES), LINES * P::G, smem, st>>>(peers, lnloc, line_begin, N1, NZP, tw, lam[0], lam[1],
                                                            lam[2], lamf[0], lamf[1], lamf[2], cutoff, norm, dtab);
  }
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

int launch_rfft_rows3(cudaStream_t st, int lm, const float* rhs, float2* T, int batch, int NR,
                      const float2* tw, const float2* rtw) {
  CFD_DISPATCH_LM(lm, 4, 14, return launch_rfft_rows3_t<LM_>(st, rhs, T, batch, NR, tw, rtw));
  return 0;
}
int launch_irfft_rows3(cudaStream_t st, int lm, const float2* T, float* q, int batch, int NR,
                       const float2* tw, const float2* rtw) {
  CFD_DISPATCH_LM(lm, 4, 14, return launch_irfft_rows3_t<LM_>(st, T, q, batch, NR, tw, rtw));
  return 0;
}
int launch_lines_scatter(cudaStream_t st, int lm, const float2* A, float2* B, int planes, int NL,
                         const float2* tw) {
  CFD_DISPATCH_LM(lm, 4, 14, return launch_lines_scatter_t<LM_>(st, A, B, planes, NL, tw));
  return 0;
}
int launch_lines_gather(cudaStream_t st, int lm, const float2* B, float2* A, int planes, int NL,
                        const float2* tw) {
  CFD_DISPATCH_LM(lm, 4, 14, return launch_lines_gather_t<LM_>(st, B, A, planes, NL, tw));
  return 0;
}
// lines [line_begin, line_begin + nlines) of the spectrum; element x of a line lives in peers.p[x >> lnloc]
int launch_xlines3_peers(cudaStream_t st, int lm, const LinePeers& peers, int lnloc, size_t line_begin,
                         size_t nlines, int N1, int NZP, const float2* tw, const double* const* lam,
                         const float* const* lamf, int fastd, double cutoff, float norm, const float* dtab) {
  CFD_DISPATCH_LM(lm, 4, 14,
                  return launch_xlines3_t<LM_>(st, peers, lnloc, line_begin, nlines, N1, NZP, tw, lam, lamf, fastd,
                                               cutoff, norm, dtab));
  return 0;
}
int launch_xlines3(cudaStream_t st, int lm, float2* T, size_t nlines, int N1, int NZP,
                   const float2* tw, const double* const* lam, const float* const* lamf, int fastd,
                   double cutoff, float norm, const float* dtab) {
  LinePeers peers;
  for (int i = 0; i < CFD_MAX_PEERS; ++i) peers.p[i] = T;
  return launch_xlines3_peers(st, lm, peers, lm, 0, nlines, N1, NZP, tw, lam, lamf, fastd, cutoff, norm, dtab);
}
int launch_divergence_generic(cudaStream_t st, const float* u, const float* v, const float* w, float* rhs,
                              int batch, int N0, int N1, int N2, float ih0, float ih1, float ih2);
int launch_correct_generic(cudaStream_t st, const float* us, const float* vs, const float* ws, const float* q,
                           float* uo, float* vo, float* wo, int batch, int N0, int N1, int N2, float ih0,
                           float ih1, float ih2);

int launch_divergence_3d_slab(cudaStream_t st, const float* u, const float* u_below, const float* v,
                              const float* w, float* rhs, int batch, int N0, int N1, int N2, float ihx,
                              float ihy, float ihz) {
  if (N2 % 4) return set_error_msg("internal: slab divergence needs N2 % 4 == 0");
  const int threads = N2 / 4 < 128 ? N2 / 4 : 128;
  dim3 grid(((N2 / 4 + threads - 1) / threads) * N0 * N1, batch);
  divergence3d_kernel<<<grid, threads, 0, st>>>(u, u_below, v, w, rhs, N0, N1, N2, ihx, ihy, ihz);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}
int launch_divergence_3d(cudaStream_t st, const float* u, const float* v, const float* w, float* rhs,
                         int batch, int N0, int N1, int N2, float ihx, float ihy, float ihz) {
  if (N2 % 4) return launch_divergence_generic(st, u, v, w, rhs, batch, N0, N1, N2, ihx, ihy, ihz);
  return launch_divergence_3d_slab(st, u, u + (size_t)(N0 - 1) * N1 * N2, v, w, rhs, batch, N0, N1, N2, ihx,
                                   ihy, ihz);
}
int launch_correct_3d_slab(cudaStream_t st, const float* us, const float* vs, const float* ws, const float* q,
                           const float* q_above, float* uo, float* vo, float* wo, int batch, int N0, int N1,
                           int N2, float ihx, float ihy, float ihz) {
  if (N2 % 4) return set_error_msg("internal: slab correction needs N2 % 4 == 0");
  const int threads = N2 / 4 < 128 ? N2 / 4 : 128;
  dim3 grid(((N2 / 4 + threads - 1) / threads) * N0 * N1, batch);
  correct3d_kernel<<<grid, threads, 0, st>>>(us, vs, ws, q, q_above, uo, vo, wo, N0, N1, N2, ihx, ihy, ihz);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}
int launch_correct_3d(cudaStream_t st, const float* us, const float* vs, const float* ws,
                      const float* q, float* uo, float* vo, float* wo, int batch, int N0, int N1,
                      int N2, float ihx, float ihy, float ihz) {
  if (N2 % 4) return launch_correct_generic(st, us, vs, ws, q, uo, vo, wo, batch, N0, N1, N2, ihx, ihy, ihz);
  return launch_correct_3d_slab(st, us, vs, ws, q, q, uo, vo, wo, batch, N0, N1, N2, ihx, ihy, ihz);
}

}  // namespace cfd
