// Pseudo-inverse scaling of one x line of the packed 2-D spectrum, shared by the x-line kernels
// (poisson_2d.cu, xlines15.cu).
#pragma once
#include "common.cuh"
#include "fft_smem.cuh"

namespace cfd {
namespace {

// Multiply the spectrum of one x line (held in registers, v[slot] <-> sub-index d = t + G*slot) by
// the pseudo-inverse table D(kx, ky) = norm / (lam_x[kx] + lam_y[ky]),  kx = kmul * d + kadd.
// (kmul, kadd) = (1, 0) for a whole line; (2, 0) / (2, 1) for the even / odd half-spectra of the
// split 32768-point transform.  The packed line ky = 0 (ky = 0 in Re, ky = Ny/2 in Im) is split
// into its two real sequences by symmetry: with C the line spectrum and C~ = conj C[N - kx],
//   C'[kx] = D(kx,0)/2 (C + C~) + D(kx,Ny/2)/2 (C - C~);   N - kx stays in the same half-spectrum.
template <class P, bool FASTD>
__device__ __forceinline__ void scale_line(float2 (&v)[P::E], int t, float2* s, int ky, int My,
                                           bool cta_has_packed, int kmul, int kadd,
                                           const double* __restrict__ lamx,
                                           const double* __restrict__ lamy,
                                           const float* __restrict__ lamxf,
                                           const float* __restrict__ lamyf, double cutoff,
                                           float norm, const float* __restrict__ dtab) {
  constexpr int M = P::M, G = P::G, E = P::E;
  // dtab != nullptr (only with FASTD == false): the diagonal comes from a caller-supplied table in
  // line layout, dtab[ky][kx] with My + 1 lines of kmul * M entries (cfd_transform: any real
  // func(eigenvalues), fast_diagonalization.py:28-126) instead of the pseudo-inverse
  const float* trow = dtab ? dtab + (size_t)ky * (size_t)(kmul * M) : nullptr;
  if (!cta_has_packed || ky != 0) {
    if (FASTD) {
      // only the mean mode is below the cutoff (checked on the host in f64): float eigenvalues,
      // |lam| >= min nonzero |lam_x|, |lam_y| > cutoff, so no per-element test is needed here
      const float ly = __ldg(lamyf + ky);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float d = norm * fast_rcp(__ldg(lamxf + kmul * (t + G * e) + kadd) + ly);
        v[e].x *= d;
        v[e].y *= d;
      }
    } else {
      if (trow) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const float d = norm * __ldg(trow + kmul * (t + G * e) + kadd);
          v[e].x *= d;
          v[e].y *= d;
        }
      } else {
        const double ly = __ldg(lamy + ky);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const double lam = __ldg(lamx + kmul * (t + G * e) + kadd) + ly;
          const float d = (fabs(lam) > cutoff) ? norm * fast_rcp((float)lam) : 0.f;
          v[e].x *= d;
          v[e].y *= d;
        }
      }
    }
  }
  if (cta_has_packed) {
    __syncthreads();
    if (ky == 0) {
#pragma unroll
      for (int e = 0; e < E; ++e) s[P::pad(t + G * e)] = v[e];
    }
    __syncthreads();
    if (ky == 0) {
      const double ly0 = __ldg(lamy + 0), lyM = __ldg(lamy + My);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int d = t + G * e;
        const int dp = kadd ? (M - 1 - d) : ((M - d) & (M - 1));  // sub-index of N - kx
        const float2 cp = s[P::pad(dp)];
        float d0, dM;
        if (dtab) {
          d0 = 0.5f * norm * __ldg(dtab + kmul * d + kadd);
          dM = 0.5f * norm * __ldg(dtab + (size_t)My * (size_t)(kmul * M) + kmul * d + kadd);
        } else {
          const double lx = __ldg(lamx + kmul * d + kadd);
          const double l0 = lx + ly0, lM = lx + lyM;
          d0 = (fabs(l0) > cutoff) ? 0.5f * norm * fast_rcp((float)l0) : 0.f;
          dM = (fabs(lM) > cutoff) ? 0.5f * norm * fast_rcp((float)lM) : 0.f;
        }
        const float2 c = v[e];
        const float2 sum = make_float2(c.x + cp.x, c.y - cp.y);
        const float2 dif = make_float2(c.x - cp.x, c.y + cp.y);
        v[e] = make_float2(d0 * sum.x + dM * dif.x, d0 * sum.y + dM * dif.y);
      }
    }
  }
}

}  // namespace
}  // namespace cfd
