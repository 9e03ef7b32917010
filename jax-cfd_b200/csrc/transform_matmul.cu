// Fast diagonalisation by matrix multiplication along each axis -- the `matmul` implementation of
// fast_diagonalization.transform (fast_diagonalization.py:129-165, `_hermitian_matmul_transform`):
//
//   out = (X_0 (x) X_1 [(x) X_2]) diag (X_0^T (x) X_1^T [(x) X_2^T]) rhs
//
// with X_j the eigenvectors of the operator along axis j (np.linalg.eigh on the host for arbitrary
// hermitian operators; the analytic real Fourier basis for the periodic Laplacian of the pressure
// solve, array_utils.py:168-173) and diag = func(sum of eigenvalues).  The reference selects it when
// the last axis is odd (fast_diagonalization.py:107-108) or on request; here it also serves every
// grid whose axes are not powers of two (the line-FFT kernels are radix-2).
//
// This is the only dense contraction on the path and the only place tensor cores are used:
// FP64 DMMA (mma.sync.m8n8k4.f64), operands widened to float64, float64 accumulation and float64
// intermediates between the per-axis products -- strictly more accurate than the reference's
// float32 `Precision.HIGHEST` tensordot; the result is rounded to float32 once at the end.
// Cost O(N) per cell and axis: meant for the small / odd grids of the reference's own tests
// (100^2, 40^3, 48x36), not for the BASELINE sizes, which take the O(log N) line FFTs.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "plan_struct.cuh"

namespace cfd {

namespace {

constexpr int kTile = 32;  // CTA tile (M and N) and K chunk
constexpr int kLd = kTile + 1;

__device__ __forceinline__ void dmma_m8n8k4(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double widen(float x) { return (double)x; }
__device__ __forceinline__ double widen(double x) { return x; }

// C[z] = A[z] B[z] (* scale), row-major, M x K times K x N; z = blockIdx.z with element strides
// sA / sB / sC (0 = shared operand).  TA / TB / TC: float or double.  `scale` (nullable): C[m][n] is
// multiplied by scale[(m % scale_rows) * N + n] -- the diagonal in the eigenbasis, fused into the
// product that completes the forward transform.
// 128 threads = 4 warps; every warp owns a 16 x 16 quarter of the 32 x 32 tile as 2 x 2 DMMA tiles.
template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(128)
dgemm_dmma_kernel(const TA* __restrict__ A, const TB* __restrict__ B, TC* __restrict__ C, int M,
                  int N, int K, size_t sA, size_t sB, size_t sC, const double* __restrict__ scale,
                  int scale_rows) {
  __shared__ double As[kTile][kLd];  // [m][k]
  __shared__ double Bs[kTile][kLd];  // [k][n]
  const size_t z = blockIdx.z;
  A += z * sA;
  B += z * sB;
  C += z * sC;
  const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 16;
  const int g = lane >> 2, t4 = lane & 3;
  double acc[2][2][2] = {};
  for (int k0 = 0; k0 < K; k0 += kTile) {
    for (int idx = tid; idx < kTile * kTile; idx += 128) {
      const int r = idx / kTile, c = idx % kTile;
      const int am = m0 + r, ak = k0 + c;
      As[r][c] = (am < M && ak < K) ? widen(A[(size_t)am * K + ak]) : 0.0;
      const int bk = k0 + r, bn = n0 + c;
      Bs[r][c] = (bk < K && bn < N) ? widen(B[(size_t)bk * N + bn]) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTile; kk += 4) {
      double a[2], b[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = As[wm + 8 * i + g][kk + t4];   // A fragment: row g, col t4
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = Bs[kk + t4][wn + 8 * j + g];   // B fragment: row t4, col g
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma_m8n8k4(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }
  // C fragment: row g, columns 2 * t4 + {0, 1}
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int m = m0 + wm + 8 * i + g, n = n0 + wn + 8 * j + 2 * t4 + e;
        if (m < M && n < N) {
          double v = acc[i][j][e];
          if (scale) v *= scale[(size_t)(m % scale_rows) * N + n];
          C[(size_t)m * N + n] = (TC)v;
        }
      }
}

template <typename TA, typename TB, typename TC>
int gemm(cudaStream_t st, const TA* A, const TB* B, TC* C, int M, int N, int K, size_t batch, size_t sA,
         size_t sB, size_t sC, const double* scale = nullptr, int scale_rows = 1) {
  // gridDim.z is limited to 65535: large batches are issued in slices
  for (size_t z0 = 0; z0 < batch; z0 += 65535) {
    const size_t nz = batch - z0 < 65535 ? batch - z0 : 65535;
    dim3 grid((N + kTile - 1) / kTile, (M + kTile - 1) / kTile, (unsigned)nz);
    dgemm_dmma_kernel<TA, TB, TC><<<grid, 128, 0, st>>>(A + z0 * sA, B + z0 * sB, C + z0 * sC, M, N, K, sA, sB,
                                                        sC, scale, scale_rows);
    count_launch();
    CFD_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

}  // namespace

// Orthonormal real eigenbasis of the periodic second-difference operator [-2, 1, 0, ..., 0, 1] / h^2
// (array_utils.py:168-173), built analytically in float64: column 0 constant, then cos / sin pairs
// of wavenumber m = 1 .. ceil(N/2) - 1, then (N even) the alternating vector.  lam[c] is the
// eigenvalue of column c.  V is row-major N x N: V[i * N + c].
void periodic_laplacian_eigenbasis(int N, double h, std::vector<double>* V, std::vector<double>* lam) {
  V->assign((size_t)N * N, 0.0);
  lam->assign(N, 0.0);
  const double s0 = 1.0 / sqrt((double)N), s = sqrt(2.0 / (double)N);
  int c = 0;
  for (int i = 0; i < N; ++i) (*V)[(size_t)i * N + c] = s0;
  (*lam)[c++] = 0.0;
  for (int m = 1; 2 * m < N; ++m) {
    const double l = (2.0 * cos(2.0 * M_PI * (double)m / (double)N) - 2.0) / (h * h);
    for (int i = 0; i < N; ++i) {
      const double ang = 2.0 * M_PI * (double)(((long long)m * i) % N) / (double)N;
      (*V)[(size_t)i * N + c] = s * cos(ang);
      (*V)[(size_t)i * N + c + 1] = s * sin(ang);
    }
    (*lam)[c] = (*lam)[c + 1] = l;
    c += 2;
  }
  if (N % 2 == 0) {
    for (int i = 0; i < N; ++i) (*V)[(size_t)i * N + c] = (i & 1) ? -s0 : s0;
    (*lam)[c++] = -4.0 / (h * h);
  }
}

// out = X diag X^T in   for `batch` members of an ndim-dimensional grid.  V[j] / Vt[j]: device
// N_j x N_j row-major eigenvectors and their transposes (float64); diag: device float64 table of the
// grid shape in the eigenbasis; w1 / w2: float64 workspaces of batch * cells elements.
int matmul_transform(cudaStream_t st, int ndim, const int64_t* shape, int batch, const float* in,
                     float* out, const double* const* V, const double* const* Vt, const double* diag,
                     double* w1, double* w2) {
  const int N0 = (int)shape[0], N1 = (int)shape[1], N2 = ndim == 3 ? (int)shape[2] : 1;
  const size_t cells = (size_t)N0 * N1 * N2;
  if (ndim == 2) {
    // forward: axis 0 (left product, per member), then axis 1 (right product over all rows) * diag
    if (int e = gemm(st, Vt[0], in, w1, N0, N1, N0, batch, 0, cells, cells)) return e;
    if (int e = gemm(st, w1, V[1], w2, batch * N0, N1, N1, 1, 0, 0, 0, diag, N0)) return e;
    // inverse: axis 1, then axis 0
    if (int e = gemm(st, w2, Vt[1], w1, batch * N0, N1, N1, 1, 0, 0, 0)) return e;
    return gemm(st, V[0], w1, out, N0, N1, N0, batch, 0, cells, cells);
  }
  const size_t plane = (size_t)N1 * N2;
  // forward: axis 0 (N0 x plane per member), axis 1 (N1 x N2 per (member, x)), axis 2 (rows) * diag
  if (int e = gemm(st, Vt[0], in, w1, N0, (int)plane, N0, batch, 0, cells, cells)) return e;
  if (int e = gemm(st, Vt[1], w1, w2, N1, N2, N1, (size_t)batch * N0, 0, plane, plane)) return e;
  if (int e = gemm(st, w2, V[2], w1, batch * N0 * N1, N2, N2, 1, 0, 0, 0, diag, N0 * N1)) return e;
  // inverse: axis 2, axis 1, axis 0
  if (int e = gemm(st, w1, Vt[2], w2, batch * N0 * N1, N2, N2, 1, 0, 0, 0)) return e;
  if (int e = gemm(st, V[1], w2, w1, N1, N2, N1, (size_t)batch * N0, 0, plane, plane)) return e;
  return gemm(st, V[0], w1, out, N0, (int)plane, N0, batch, 0, cells, cells);
}

}  // namespace cfd
