// Smagorinsky eddy-viscosity closure on 2-D grids (subgrid_models.py:40-134 is dimension
// agnostic; the reference's own tests run it at (100, 100), subgrid_models_test.py:117-217).
//
//   smag_nut2d_kernel:  nu_t = (cs * Delta)^2 * sqrt(2 * trace(S.S)) at cell centres, S = the strain
//                       rate interpolated to the centre (subgrid_models.py:88-93)
//   smag_add2d_kernel:  u* += (dt / rho) * acc  (or dv/dt += acc / rho),
//                       acc_i = -sum_j (tau_ij - S(tau_ij, -1, j)) / h_j,  tau_ij = -2 nu_ij s_ij
//                       with nu_t interpolated back to the offset of s_ij  (subgrid_models.py:94-97,
//                       124-134): the closure enters as the LAST forcing term (subgrid_models.py:188-213).
//
// One thread per cell, neighbours through the read-only cache: config #5 (the closure's benchmark) is
// 3-D (explicit_3d.cu has the tiled kernels); this path exists for coverage, not for speed.
#include "common.cuh"

namespace cfd {

namespace {

struct Idx2 {
  int i, j, N0, N1;
};
__device__ __forceinline__ float at2d(const float* __restrict__ f, const Idx2& ix, int di, int dj) {
  const int i = wrap_idx(ix.i + di, ix.N0), j = wrap_idx(ix.j + dj, ix.N1);
  return __ldg(f + (size_t)i * ix.N1 + j);
}

// s_IJ = 0.5 (D+_J v_I + D+_I v_J) at the cell shifted by (p0, p1)   (subgrid_models.py:124-128)
template <int I, int J>
__device__ __forceinline__ float strain2(const float* const (&v)[2], const Idx2& ix, int p0, int p1,
                                         const float* inv_h) {
  const float a = (at2d(v[I], ix, p0 + (J == 0), p1 + (J == 1)) - at2d(v[I], ix, p0, p1)) * inv_h[J];
  const float b = (at2d(v[J], ix, p0 + (I == 0), p1 + (I == 1)) - at2d(v[J], ix, p0, p1)) * inv_h[I];
  return 0.5f * (a + b);
}

// nu_t interpolated from the centres to the offset of s_IJ, at the cell shifted by (p0, p1):
// a pure shift on the diagonal, the four-point mean (axis 0 first, then axis 1: interpolation.linear
// applied along each axis in turn) off the diagonal
template <int I, int J>
__device__ __forceinline__ float nu_at2(const float* __restrict__ nut, const Idx2& ix, int p0, int p1) {
  if (I == J) return at2d(nut, ix, p0 + (I == 0), p1 + (I == 1));
  const float r0 = 0.5f * at2d(nut, ix, p0, p1) + 0.5f * at2d(nut, ix, p0 + 1, p1);
  const float r1 = 0.5f * at2d(nut, ix, p0, p1 + 1) + 0.5f * at2d(nut, ix, p0 + 1, p1 + 1);
  return 0.5f * r0 + 0.5f * r1;
}

template <int I, int J>
__device__ __forceinline__ float tau2(const float* const (&v)[2], const float* __restrict__ nut,
                                      const Idx2& ix, int p0, int p1, const float* inv_h) {
  return -2.f * nu_at2<I, J>(nut, ix, p0, p1) * strain2<I, J>(v, ix, p0, p1, inv_h);
}

template <int I>
__device__ __forceinline__ float smag_acc2(const float* const (&v)[2], const float* __restrict__ nut,
                                           const Idx2& ix, const float* inv_h) {
  float d = (tau2<I, 0>(v, nut, ix, 0, 0, inv_h) - tau2<I, 0>(v, nut, ix, -1, 0, inv_h)) * inv_h[0];
  d += (tau2<I, 1>(v, nut, ix, 0, 0, inv_h) - tau2<I, 1>(v, nut, ix, 0, -1, inv_h)) * inv_h[1];
  return -d;
}

__global__ void smag_nut2d_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                  float* __restrict__ nut, int N0, int N1, StepConsts c) {
  const size_t cells = (size_t)N0 * N1;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells) return;
  const size_t boff = (size_t)blockIdx.y * cells;
  const Idx2 ix = {(int)(gid / N1), (int)(gid % N1), N0, N1};
  const float* const vel[2] = {u + boff, v + boff};
  // strain interpolated to the centre: the diagonal is a pure shift; s_01 is the four-point mean
  // over (-1, 0) on both axes, axis 0 first
  const float s00 = strain2<0, 0>(vel, ix, -1, 0, c.inv_h);
  const float s11 = strain2<1, 1>(vel, ix, 0, -1, c.inv_h);
  const float lo = 0.5f * strain2<0, 1>(vel, ix, -1, -1, c.inv_h) + 0.5f * strain2<0, 1>(vel, ix, 0, -1, c.inv_h);
  const float hi = 0.5f * strain2<0, 1>(vel, ix, -1, 0, c.inv_h) + 0.5f * strain2<0, 1>(vel, ix, 0, 0, c.inv_h);
  const float s01 = 0.5f * lo + 0.5f * hi;
  // trace(S.S) row by row like np.trace(S.dot(S)); s_10 == s_01 bitwise
  const float r0 = s00 * s00 + s01 * s01;
  const float r1 = s01 * s01 + s11 * s11;
  nut[boff + gid] = c.smag_coef * sqrtf(2.f * (r0 + r1));
}

__global__ void smag_add2d_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                  const float* __restrict__ nut, float* __restrict__ us,
                                  float* __restrict__ vs, int N0, int N1, StepConsts c, int dvdt_mode) {
  const size_t cells = (size_t)N0 * N1;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells) return;
  const size_t boff = (size_t)blockIdx.y * cells;
  const Idx2 ix = {(int)(gid / N1), (int)(gid % N1), N0, N1};
  const float* const vel[2] = {u + boff, v + boff};
  const float scale = (dvdt_mode ? 1.f : c.dt) * c.inv_rho;
  us[boff + gid] += scale * smag_acc2<0>(vel, nut + boff, ix, c.inv_h);
  vs[boff + gid] += scale * smag_acc2<1>(vel, nut + boff, ix, c.inv_h);
}

}  // namespace

int launch_smag_nut_2d(cudaStream_t st, const float* u, const float* v, float* nut, int batch, int N0,
                       int N1, const StepConsts& c) {
  const size_t cells = (size_t)N0 * N1;
  dim3 grid((unsigned)((cells + 127) / 128), batch);
  smag_nut2d_kernel<<<grid, 128, 0, st>>>(u, v, nut, N0, N1, c);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_smag_add_2d(cudaStream_t st, const float* u, const float* v, const float* nut, float* us,
                       float* vs, int batch, int N0, int N1, const StepConsts& c, int dvdt_mode) {
  const size_t cells = (size_t)N0 * N1;
  dim3 grid((unsigned)((cells + 127) / 128), batch);
  smag_add2d_kernel<<<grid, 128, 0, st>>>(u, v, nut, us, vs, N0, N1, c, dvdt_mode);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfd
