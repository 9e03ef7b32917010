// One-thread-per-cell kernels for grid shapes the tuned kernels do not take (last axis not a
// multiple of 4, odd axes, fewer than 3 rows): the same arithmetic per face (common.cuh: face_flux)
// with periodic index wraps and neighbours through the read-only cache.  The reference's own tests
// run such shapes ((33, 27), (30, 20), (100, 100), 40^3); the BASELINE sizes never come here.
//
//   explicit2d_generic_kernel  advection.py:387-395, interpolation.py:36-303, diffusion.py:35-37,
//                              forcings.py:35-129, equations.py:102-114, time_stepping.py:101
//   divergence / correct       finite_differences.py:136-143, pressure.py:194-196  (2-D and 3-D)
#include "common.cuh"

namespace cfd {

namespace {

struct Ix2 {
  int oi[5], oj[5];  // linear offsets of i-2..i+2 (times N1) and j-2..j+2, wrapped
};
__device__ __forceinline__ Ix2 make_ix2(int i, int j, int N0, int N1) {
  Ix2 r;
#pragma unroll
  for (int d = -2; d <= 2; ++d) {
    r.oi[d + 2] = wrap_idx(i + d, N0) * N1;
    r.oj[d + 2] = wrap_idx(j + d, N1);
  }
  return r;
}
__device__ __forceinline__ float at(const float* __restrict__ f, const Ix2& ix, int d0, int d1) {
  return __ldg(f + ix.oi[d0 + 2] + ix.oj[d1 + 2]);
}
template <int A, int B>
__device__ __forceinline__ float at2(const float* __restrict__ f, const Ix2& ix, int sA, int sB) {
  int d[2] = {0, 0};
  d[A] += sA;
  d[B] += sB;
  return at(f, ix, d[0], d[1]);
}

// (F_J(cell) - F_J(cell - e_J)) / h_J for component A advected along J   (advection.py:73-78)
template <int A, int J>
__device__ __forceinline__ float conv_dir2(const float* __restrict__ c, const float* __restrict__ vj,
                                           const Ix2& ix, float dth, float inv_h) {
  const float Up = 0.5f * (at2<J, A>(vj, ix, 0, 0) + at2<J, A>(vj, ix, 0, 1));   // interpolation.py:57-62
  const float Fp = face_flux(at2<J, J>(c, ix, -1, 0), at2<J, J>(c, ix, 0, 0), at2<J, J>(c, ix, 1, 0),
                             at2<J, J>(c, ix, 2, 0), Up, dth);
  const float Um = 0.5f * (at2<J, A>(vj, ix, -1, 0) + at2<J, A>(vj, ix, -1, 1));
  const float Fm = face_flux(at2<J, J>(c, ix, -2, 0), at2<J, J>(c, ix, -1, 0), at2<J, J>(c, ix, 0, 0),
                             at2<J, J>(c, ix, 1, 0), Um, dth);
  return (Fp - Fm) * inv_h;
}

template <int A>
__device__ __forceinline__ float explicit_comp2(const float* const (&vel)[2], const Ix2& ix,
                                                const StepConsts& c, int i, int j, size_t cell) {
  const float* cc = vel[A];
  const float c0 = at(cc, ix, 0, 0);
  float conv = conv_dir2<A, 0>(cc, vel[0], ix, c.dth[0], c.inv_h[0]);
  conv += conv_dir2<A, 1>(cc, vel[1], ix, c.dth[1], c.inv_h[1]);
  float dv = -conv;
  if (c.has_nu) {  // finite_differences.py:127-133
    float l = (-2.f * c0) * c.lap_sum;
    l += (at(cc, ix, -1, 0) + at(cc, ix, 1, 0)) * c.lap_s[0];
    l += (at(cc, ix, 0, -1) + at(cc, ix, 0, 1)) * c.lap_s[1];
    dv += c.nu * l;
  }
  if (c.n_terms > 0) {  // forcings.py:125-129 (left-to-right sum), equations.py:108-109
    float f = 0.f;
    for (int t = 0; t < c.n_terms; ++t) {
      const int kind = c.term_kind[t];
      if (kind == CFD_FORCE_SEPARABLE) {
        if (c.has_sep[A]) {
          float p = 1.f;
          if (c.sep_prof[A][0]) p = __ldg(c.sep_prof[A][0] + i);
          if (c.sep_prof[A][1]) p = p * __ldg(c.sep_prof[A][1] + j);
          f += p * c.sep_scale[A];
        }
      } else if (kind == CFD_FORCE_FIELD) {
        if (c.field[A]) f += __ldg(c.field[A] + cell);
      } else if (kind == CFD_FORCE_LINEAR) {
        f += c.linear_coef * c0;
      }  // CFD_FORCE_SMAGORINSKY: added by smag_add2d_kernel
    }
    dv = fmaf(f, c.inv_rho, dv);
  }
  return dv;
}

__global__ void __launch_bounds__(128)
explicit2d_generic_kernel(const float* __restrict__ u, const float* __restrict__ v,
                          float* __restrict__ us, float* __restrict__ vs, int N0, int N1, StepConsts c,
                          int dvdt_mode) {
  const size_t cells = (size_t)N0 * N1;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells) return;
  const size_t boff = (size_t)blockIdx.y * cells;
  const int i = (int)(gid / N1), j = (int)(gid % N1);
  const Ix2 ix = make_ix2(i, j, N0, N1);
  const float* const vel[2] = {u + boff, v + boff};
  const float d0 = explicit_comp2<0>(vel, ix, c, i, j, gid);
  const float d1 = explicit_comp2<1>(vel, ix, c, i, j, gid);
  us[boff + gid] = dvdt_mode ? d0 : __ldg(vel[0] + gid) + c.dt * d0;  // time_stepping.py:101
  vs[boff + gid] = dvdt_mode ? d1 : __ldg(vel[1] + gid) + c.dt * d1;
}

// rhs = sum_j (v_j - S(v_j, -1, j)) / h_j   for ndim = 2 (w == nullptr) or 3
__global__ void divergence_generic_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                          const float* __restrict__ w, float* __restrict__ rhs, int N0,
                                          int N1, int N2, float ih0, float ih1, float ih2) {
  const size_t cells = (size_t)N0 * N1 * N2;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells) return;
  const size_t boff = (size_t)blockIdx.y * cells;
  const int k = (int)(gid % N2), j = (int)((gid / N2) % N1), i = (int)(gid / ((size_t)N1 * N2));
  const size_t im = ((size_t)(i == 0 ? N0 - 1 : i - 1) * N1 + j) * N2 + k;
  const size_t jm = ((size_t)i * N1 + (j == 0 ? N1 - 1 : j - 1)) * N2 + k;
  float d = (__ldg(u + boff + gid) - __ldg(u + boff + im)) * ih0 + (__ldg(v + boff + gid) - __ldg(v + boff + jm)) * ih1;
  if (w != nullptr) {
    const size_t km = ((size_t)i * N1 + j) * N2 + (k == 0 ? N2 - 1 : k - 1);
    d += (__ldg(w + boff + gid) - __ldg(w + boff + km)) * ih2;
  }
  rhs[boff + gid] = d;
}

// v'_j = u*_j - (S(q, +1, j) - q) / h_j
__global__ void correct_generic_kernel(const float* __restrict__ us, const float* __restrict__ vs,
                                       const float* __restrict__ ws, const float* __restrict__ q,
                                       float* __restrict__ uo, float* __restrict__ vo,
                                       float* __restrict__ wo, int N0, int N1, int N2, float ih0,
                                       float ih1, float ih2) {
  const size_t cells = (size_t)N0 * N1 * N2;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells) return;
  const size_t boff = (size_t)blockIdx.y * cells;
  const int k = (int)(gid % N2), j = (int)((gid / N2) % N1), i = (int)(gid / ((size_t)N1 * N2));
  const size_t ip = ((size_t)(i == N0 - 1 ? 0 : i + 1) * N1 + j) * N2 + k;
  const size_t jp = ((size_t)i * N1 + (j == N1 - 1 ? 0 : j + 1)) * N2 + k;
  const float q0 = __ldg(q + boff + gid);
  uo[boff + gid] = __ldg(us + boff + gid) - (__ldg(q + boff + ip) - q0) * ih0;
  vo[boff + gid] = __ldg(vs + boff + gid) - (__ldg(q + boff + jp) - q0) * ih1;
  if (ws != nullptr) {
    const size_t kp = ((size_t)i * N1 + j) * N2 + (k == N2 - 1 ? 0 : k + 1);
    wo[boff + gid] = __ldg(ws + boff + gid) - (__ldg(q + boff + kp) - q0) * ih2;
  }
}

}  // namespace

int launch_explicit_2d_generic(cudaStream_t st, const float* u, const float* v, float* us, float* vs,
                               int batch, int N0, int N1, const StepConsts& c, int dvdt_mode) {
  const size_t cells = (size_t)N0 * N1;
  dim3 grid((unsigned)((cells + 127) / 128), batch);
  explicit2d_generic_kernel<<<grid, 128, 0, st>>>(u, v, us, vs, N0, N1, c, dvdt_mode);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_divergence_generic(cudaStream_t st, const float* u, const float* v, const float* w, float* rhs,
                              int batch, int N0, int N1, int N2, float ih0, float ih1, float ih2) {
  const size_t cells = (size_t)N0 * N1 * N2;
  dim3 grid((unsigned)((cells + 127) / 128), batch);
  divergence_generic_kernel<<<grid, 128, 0, st>>>(u, v, w, rhs, N0, N1, N2, ih0, ih1, ih2);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_correct_generic(cudaStream_t st, const float* us, const float* vs, const float* ws, const float* q,
                           float* uo, float* vo, float* wo, int batch, int N0, int N1, int N2, float ih0,
                           float ih1, float ih2) {
  const size_t cells = (size_t)N0 * N1 * N2;
  dim3 grid((unsigned)((cells + 127) / 128), batch);
  correct_generic_kernel<<<grid, 128, 0, st>>>(us, vs, ws, q, uo, vo, wo, N0, N1, N2, ih0, ih1, ih2);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfd
