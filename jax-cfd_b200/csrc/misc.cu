// Small elementwise / reduction kernels: RK stage combination and diagnostics.
#include "common.cuh"

namespace cfd {

namespace {

struct AxpyArgs {
  const float* y[4];
  float coef[4];
};

// out = x + sum_k coef[k] * y[k]   (time_stepping.py:96-101: u0 + dt * sum(a_ij k_j))
__global__ void axpy_kernel(const float* __restrict__ x, int nterms, AxpyArgs a,
                            float* __restrict__ out, size_t n4, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (i < n - 4 * n4) {  // tail of a field whose size is not a multiple of 4
    const size_t e = 4 * n4 + i;
    float acc = 0.f;
    for (int k = 0; k < nterms; ++k) acc = k == 0 ? a.coef[0] * a.y[0][e] : acc + a.coef[k] * a.y[k][e];
    out[e] = x[e] + acc;
  }
  for (; i < n4; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < nterms; ++k) {
      const float4 yk = ldg4(a.y[k] + 4 * i);
      const float ck = a.coef[k];
      if (k == 0) {
        acc = make_float4(ck * yk.x, ck * yk.y, ck * yk.z, ck * yk.w);
      } else {
        acc.x += ck * yk.x;
        acc.y += ck * yk.y;
        acc.z += ck * yk.z;
        acc.w += ck * yk.w;
      }
    }
    const float4 xv = ldg4(x + 4 * i);
    stg4(out + 4 * i, make_float4(xv.x + acc.x, xv.y + acc.y, xv.z + acc.z, xv.w + acc.w));
  }
}

// out = numer * x / denom, evaluated in that order in float32 (initial_conditions.py:118-121:
// `maximum_velocity * u / max_speed`)
__global__ void scale_kernel(const float* __restrict__ x, float numer, float denom, float* __restrict__ out,
                             size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __fdiv_rn(numer * x[i], denom);
}

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// NaN-propagating max (jnp.max semantics, equations.py:68: a blown-up state must not yield a
// finite time step); fmaxf would drop the NaN
__device__ __forceinline__ float nanmax(float a, float b) { return (a > b || a != a) ? a : b; }
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void atomic_max_double(double* addr, double val) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    const double cur = __longlong_as_double(assumed);
    if (cur != cur || (cur >= val && val == val)) break;  // NaN sticks
    old = atomicCAS(a, assumed, __double_as_longlong(val));
  } while (assumed != old);
}

// out4 = { sum 0.5(u^2+v^2), sum 0.5 w^2, max|div|, max(u^2+v^2) }
//   kinetic energy / enstrophy: data/xarray_utils.py:155-188;  divergence:
//   finite_differences.py:136-143;  max speed: equations.py:68.
__global__ void diag2d_kernel(const float* __restrict__ u, const float* __restrict__ v, int Nx,
                              int Ny, size_t total, float inv_hx, float inv_hy,
                              double* __restrict__ out4) {
  double ke = 0.0, ens = 0.0;
  float mdiv = 0.f, msp = 0.f;
  const size_t per = (size_t)Nx * Ny;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = idx / per, rem = idx % per;
    const int i = (int)(rem / Ny), j = (int)(rem % Ny);
    const float* ub = u + b * per;
    const float* vb = v + b * per;
    const int im = i == 0 ? Nx - 1 : i - 1, ip = i == Nx - 1 ? 0 : i + 1;
    const int jm = j == 0 ? Ny - 1 : j - 1, jp = j == Ny - 1 ? 0 : j + 1;
    const float u0 = ub[rem], v0 = vb[rem];
    const float div = (u0 - ub[(size_t)im * Ny + j]) * inv_hx + (v0 - vb[(size_t)i * Ny + jm]) * inv_hy;
    const float w = (vb[(size_t)ip * Ny + j] - v0) * inv_hx - (ub[(size_t)i * Ny + jp] - u0) * inv_hy;
    const float sp = u0 * u0 + v0 * v0;
    ke += 0.5 * (double)sp;
    ens += 0.5 * (double)w * (double)w;
    mdiv = nanmax(fabsf(div), mdiv);
    msp = nanmax(sp, msp);
  }
  ke = warp_sum(ke);
  ens = warp_sum(ens);
  mdiv = warp_max(mdiv);
  msp = warp_max(msp);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out4 + 0, ke);
    atomicAdd(out4 + 1, ens);
    atomic_max_double(out4 + 2, (double)mdiv);
    atomic_max_double(out4 + 3, (double)msp);
  }
}

}  // namespace

int launch_axpy(cudaStream_t st, const float* x, int nterms, const float* const* y,
                const float* coef, float* out, size_t n) {
  AxpyArgs a;
  for (int k = 0; k < 4; ++k) {
    a.y[k] = k < nterms ? y[k] : nullptr;
    a.coef[k] = k < nterms ? coef[k] : 0.f;
  }
  const size_t n4 = n / 4;
  const int threads = 256;
  const int blocks = (int)((n4 + threads - 1) / threads < 148 * 16 ? (n4 + threads - 1) / threads : 148 * 16);
  axpy_kernel<<<blocks > 0 ? blocks : 1, threads, 0, st>>>(x, nterms, a, out, n4, n);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_scale(cudaStream_t st, const float* x, float numer, float denom, float* out, size_t n) {
  const int threads = 256;
  size_t blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  scale_kernel<<<(int)blocks, threads, 0, st>>>(x, numer, denom, out, n);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_diag_2d(cudaStream_t st, const float* u, const float* v, int batch, int Nx, int Ny,
                   float inv_hx, float inv_hy, double* out4) {
  CFD_CUDA_OK(cudaMemsetAsync(out4, 0, 4 * sizeof(double), st));
  const size_t total = (size_t)batch * Nx * Ny;
  const int threads = 256;
  size_t blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  diag2d_kernel<<<(int)blocks, threads, 0, st>>>(u, v, Nx, Ny, total, inv_hx, inv_hy, out4);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfd
