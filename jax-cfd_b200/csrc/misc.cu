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

// resize.downsample_staggered_velocity_component (resize.py:38-74): of the fine faces normal to
// `direction` keep those that lie on a coarse face (index (I + 1) f - 1 along `direction`) and
// average the f^(d-1) of them that tile the coarse face.  One thread per coarse value; shapes are
// the FINE grid's (N2 = 1 in 2-D), row-major with a leading batch axis.  The sum runs over the
// block in row-major order in float32 and is divided by the count (jnp.mean).
__global__ void downsample_component_kernel(const float* __restrict__ in, float* __restrict__ out, int N0,
                                            int N1, int N2, int f, int direction, size_t total) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int M0 = N0 / f, M1 = N1 / f, M2 = N2 > 1 ? N2 / f : 1;
  const int K = (int)(idx % M2), J = (int)((idx / M2) % M1), I = (int)((idx / ((size_t)M2 * M1)) % M0);
  const size_t b = idx / ((size_t)M2 * M1 * M0);
  const float* src = in + b * (size_t)N0 * N1 * N2;
  const int f2 = N2 > 1 ? f : 1;
  const int lo[3] = {I * f, J * f, K * f2};
  int cnt[3] = {f, f, f2};
  int beg[3] = {lo[0], lo[1], lo[2]};
  beg[direction] = lo[direction] + (direction == 2 ? f2 : f) - 1;
  cnt[direction] = 1;
  float sum = 0.f;
  for (int a = 0; a < cnt[0]; ++a)
    for (int c = 0; c < cnt[1]; ++c)
      for (int d = 0; d < cnt[2]; ++d)
        sum += __ldg(src + ((size_t)(beg[0] + a) * N1 + (beg[1] + c)) * N2 + (beg[2] + d));
  out[idx] = __fdiv_rn(sum, (float)(cnt[0] * cnt[1] * cnt[2]));
}

// vorticity at offset (1, 1): (v[i+1][j] - v[i][j]) / dx - (u[i][j+1] - u[i][j]) / dy, periodic
// (data/xarray_utils.py:155-163)
__global__ void vorticity2d_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                   float* __restrict__ out, int N0, int N1, float dx, float dy, size_t total) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % N1), i = (int)((idx / N1) % N0);
  const size_t base = idx - (size_t)i * N1 - j;
  const int ip = i + 1 == N0 ? 0 : i + 1, jp = j + 1 == N1 ? 0 : j + 1;
  const float dv_dx = __fdiv_rn(__ldg(v + base + (size_t)ip * N1 + j) - __ldg(v + idx), dx);
  const float du_dy = __fdiv_rn(__ldg(u + base + (size_t)i * N1 + jp) - __ldg(u + idx), dy);
  out[idx] = dv_dx - du_dy;
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

int launch_downsample_component(cudaStream_t st, const float* in, float* out, int batch, int N0, int N1, int N2,
                                int factor, int direction) {
  const size_t total = (size_t)batch * (N0 / factor) * (N1 / factor) * (N2 > 1 ? N2 / factor : 1);
  if (total == 0) return 0;
  downsample_component_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, N0, N1, N2, factor,
                                                                               direction, total);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_vorticity_2d(cudaStream_t st, const float* u, const float* v, float* out, int batch, int N0, int N1,
                        float dx, float dy) {
  const size_t total = (size_t)batch * N0 * N1;
  vorticity2d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(u, v, out, N0, N1, dx, dy, total);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfd
