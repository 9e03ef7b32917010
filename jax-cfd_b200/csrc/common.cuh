// Shared declarations for the CUDA core (no framework types).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <unordered_map>

#include "../../include/cfd_b200.h"

namespace cfd {

// Scalars of one step, rounded to float32 exactly where x64-disabled JAX rounds them
// (SURVEY.md Appendix A): (dt/h_j) and viscosity/density are formed in double then rounded once
// (interpolation.py:210, equations.py:107); the Laplacian scales are formed in float32
// (finite_differences.py:129).
struct StepConsts {
  float dt;                 // time_stepping.py:101
  float dth[CFD_MAX_DIM];   // (dt / h_j)
  float inv_h[CFD_MAX_DIM]; // 1 / h_j   (divisions by h_j become multiplications: <= 1 ulp)
  float lap_s[CFD_MAX_DIM]; // square(1 / float32(h_j))
  float lap_sum;            // float32 sum of lap_s
  float lap_m2sum;          // -2 * lap_sum (exact): (-2 c) * sum == c * (-2 sum) bitwise
  float nu;                 // viscosity / density
  float rho;                // density (forcing / rho, equations.py:109)
  float inv_rho;            // 1 / density
  int has_nu;
  // forcing
  int n_terms;
  int term_kind[CFD_MAX_FORCING_TERMS];
  float linear_coef;
  float smag_coef;          // (cs * cutoff)^2, cutoff = prod(h)^(1/d)   subgrid_models.py:91-93
  const float* sep_prof[CFD_MAX_DIM][CFD_MAX_DIM];
  float sep_scale[CFD_MAX_DIM];
  int has_sep[CFD_MAX_DIM];
  const float* field[CFD_MAX_DIM];
};

// One input field of a slab-decomposed grid: this rank's rows and the neighbouring ranks' buffers
// (peer-mapped); on a single GPU all three point at the same array.
struct SlabSrc {
  const float* prev;
  const float* own;
  const float* next;
};

// Slab-local spectrum buffers of all ranks of a distributed FFT (peer-mapped with CUDA IPC).
#define CFD_MAX_PEERS 8
struct LinePeers {
  float2* p[CFD_MAX_PEERS];
};

// Side streams + events used to overlap the chunks of the split 32768-point x pass.
struct SideStreams {
  int n = 0;
  cudaStream_t s[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t start = nullptr;
  cudaEvent_t done[4] = {nullptr, nullptr, nullptr, nullptr};
};

__device__ __forceinline__ int wrap_idx(int i, int n) {  // i in [-n, 2n)
  i = i < 0 ? i + n : i;
  return i >= n ? i - n : i;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming (evict-first) variants for data touched once per sweep
__device__ __forceinline__ float4 ldcs4(const float* p) {
  return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void stcs4(float* p, float4 v) {
  __stcs(reinterpret_cast<float4*>(p), v);
}

// MUFU.RCP (about 1 ulp); the IEEE __frcp_rn expands to a slow-path subroutine.
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One TVD-limited face flux  F = c_face * U :
//   upwind (interpolation.py:147-151), Lax-Wendroff (interpolation.py:210-217), van Leer limiter
//   on r+ / r- (interpolation.py:224-231, 287-297), flux = c * u (advection.py:73).
// For U <= 0 the stencil is mirrored so that one formula serves both signs:
//   A = upwind value, B = downwind value, L = the value behind the upwind one,
//   d = B - A, num = A - L, r = num / d, phi = 2r / (1 + r) for r > 0 else 0,
//   face = A + 0.5 (1 - |C|) d phi = A + (1 - |C|) * d * num / (d + num)      (r > 0 <=> d num > 0)
// which needs a single MUFU.RCP-based division; safe_div's "y == 0 -> 1" case gives 0 either way.
__device__ __forceinline__ float face_flux(float cL, float c0, float cR, float cRR, float U,
                                           float dth) {
  const bool pos = U > 0.f;
  const float A = pos ? c0 : cR;
  const float B = pos ? cR : c0;
  const float L = pos ? cL : cRR;
  const float d = B - A;
  const float num = A - L;
  const float g = d * num;
  const float w = fmaf(-dth, fabsf(U), 1.f);  // 1 - |C|,  C = (dt / h) U   interpolation.py:210
  const float half = (g > 0.f) ? w * (g * fast_rcp(d + num)) : 0.f;
  return (A + half) * U;
}

#define CFD_CUDA_OK(expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return cfd::set_error(#expr, _e, __FILE__, __LINE__); \
  } while (0)

int set_error(const char* what, cudaError_t e, const char* file, int line);
int set_error_msg(const char* msg);
void count_launch(int n = 1);

// Opt-in to `bytes` of dynamic shared memory for `kernel` on the CURRENT device, once per (kernel,
// device): the attribute is per device, a process may drive several, and the call is kept off the
// launch path (CUDA-graph captures then record plain launches).
template <typename K>
int opt_in_smem(K kernel, size_t bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, size_t> done[64];  // per device, keyed by kernel
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = done[dev & 63][reinterpret_cast<const void*>(kernel)];
  if (bytes > have) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(smem)", e, __FILE__, __LINE__);
    have = bytes;
  }
  return 0;
}

}  // namespace cfd
