// Sweep "E" (3-D): explicit terms + Euler update + divergence, and the Smagorinsky eddy viscosity.
//
// Two implementations with the same arithmetic per face (common.cuh: face_flux):
//   * explicit3d_march_kernel (+ smag_nut_march_kernel / smag_acc_march_kernel): 2.5-D blocking, a
//     CTA owns an 8 x 64 tile in (y, z) and marches along x with a ring of shared-memory planes --
//     used whenever the tile divides the grid;
//   * explicit3d_kernel (+ smag_nut3d_kernel / the strain-field kernels): one thread per cell,
//     neighbours through L1/L2 -- the fallback for other shapes and the first implementation.
//
//   advection      advection.py:387-395 -> 81-116 -> 34-78, interpolation.py:36-303
//   diffusion      diffusion.py:35-37, finite_differences.py:127-133
//   forcing        forcings.py:35-129 (separable / field / linear), equations.py:108-109
//   Smagorinsky    subgrid_models.py:40-98 (viscosity at cell centres),
//                  subgrid_models.py:101-134 (evm_model: -div(tau), tau_ij = -2 nu_ij s_ij)
//   update         time_stepping.py:101;   divergence   finite_differences.py:136-143
#include <stdlib.h>

#include "common.cuh"

namespace cfd {

namespace {

struct Idx3 {
  int ox[5], oy[5], oz[5];  // linear offsets of i-2..i+2 (wrapped) along each axis
};

__device__ __forceinline__ Idx3 make_idx(int i, int j, int k, int N0, int N1, int N2) {
  Idx3 r;
#pragma unroll
  for (int d = -2; d <= 2; ++d) {
    r.ox[d + 2] = wrap_idx(i + d, N0) * N1 * N2;
    r.oy[d + 2] = wrap_idx(j + d, N1) * N2;
    r.oz[d + 2] = wrap_idx(k + d, N2);
  }
  return r;
}

// value of f at the cell shifted by (d0, d1, d2)
__device__ __forceinline__ float at(const float* __restrict__ f, const Idx3& ix, int d0, int d1, int d2) {
  return __ldg(f + ix.ox[d0 + 2] + ix.oy[d1 + 2] + ix.oz[d2 + 2]);
}
// shift sA along axis A plus sB along axis B (A may equal B)
template <int A, int B>
__device__ __forceinline__ float at2(const float* __restrict__ f, const Idx3& ix, int sA, int sB) {
  int d[3] = {0, 0, 0};
  d[A] += sA;
  d[B] += sB;
  return at(f, ix, d[0], d[1], d[2]);
}

// (F_J(cell) - F_J(cell - e_J)) / h_J for component A advected along J   (advection.py:73-78)
template <int A, int J>
__device__ __forceinline__ float conv_dir(const float* __restrict__ c, const float* __restrict__ vj,
                                          const Idx3& ix, float dth, float inv_h) {
  const float Up = 0.5f * (at2<J, A>(vj, ix, 0, 0) + at2<J, A>(vj, ix, 0, 1));
  const float Fp = face_flux(at2<J, J>(c, ix, -1, 0), at2<J, J>(c, ix, 0, 0), at2<J, J>(c, ix, 1, 0),
                             at2<J, J>(c, ix, 2, 0), Up, dth);
  const float Um = 0.5f * (at2<J, A>(vj, ix, -1, 0) + at2<J, A>(vj, ix, -1, 1));
  const float Fm = face_flux(at2<J, J>(c, ix, -2, 0), at2<J, J>(c, ix, -1, 0), at2<J, J>(c, ix, 0, 0),
                             at2<J, J>(c, ix, 1, 0), Um, dth);
  return (Fp - Fm) * inv_h;
}

struct Vel3 {
  const float* f[3];
};

// forward difference of component C along axis AX at the cell shifted by (p0,p1,p2)
template <int AX>
__device__ __forceinline__ float fwd(const float* __restrict__ f, const Idx3& ix, int p0, int p1,
                                     int p2, float inv_h) {
  int d[3] = {p0, p1, p2};
  const float a = at(f, ix, d[0], d[1], d[2]);
  d[AX] += 1;
  return (at(f, ix, d[0], d[1], d[2]) - a) * inv_h;
}
// strain rate s_IJ at the cell shifted by p   (subgrid_models.py:124-128)
template <int I, int J>
__device__ __forceinline__ float strain(const Vel3& v, const Idx3& ix, int p0, int p1, int p2,
                                        const float* inv_h) {
  return 0.5f * (fwd<J>(v.f[I], ix, p0, p1, p2, inv_h[J]) + fwd<I>(v.f[J], ix, p0, p1, p2, inv_h[I]));
}
// s_IJ interpolated to the cell centre (subgrid_models.py:88-89 via interpolation.linear)
template <int I, int J>
__device__ __forceinline__ float strain_center(const Vel3& v, const Idx3& ix, const float* inv_h) {
  if (I == J) {
    int d[3] = {0, 0, 0};
    d[I] = -1;
    return strain<I, I>(v, ix, d[0], d[1], d[2], inv_h);
  }
  constexpr int A = I < J ? I : J, B = I < J ? J : I;  // ascending axis order
  float r[2];
#pragma unroll
  for (int sb = 0; sb < 2; ++sb) {  // sb = 0: shifted -1 along B, 1: unshifted
    int d0[3] = {0, 0, 0}, d1[3] = {0, 0, 0};
    d0[B] = d1[B] = sb - 1;
    d0[A] = -1;
    const float lo = strain<I, J>(v, ix, d0[0], d0[1], d0[2], inv_h);
    const float hi = strain<I, J>(v, ix, d1[0], d1[1], d1[2], inv_h);
    r[sb] = 0.5f * lo + 0.5f * hi;
  }
  return 0.5f * r[0] + 0.5f * r[1];
}

// nu_t at cell centres   (subgrid_models.py:91-93)
__global__ void smag_nut3d_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                  const float* __restrict__ w, float* __restrict__ nut, int N0,
                                  int N1, int N2, StepConsts c) {
  const size_t cells = (size_t)N0 * N1 * N2;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells) return;
  const size_t boff = (size_t)blockIdx.y * cells;
  const int k = (int)(gid % N2), j = (int)((gid / N2) % N1), i = (int)(gid / ((size_t)N1 * N2));
  const Idx3 ix = make_idx(i, j, k, N0, N1, N2);
  Vel3 vel = {{u + boff, v + boff, w + boff}};
  const float s00 = strain_center<0, 0>(vel, ix, c.inv_h), s11 = strain_center<1, 1>(vel, ix, c.inv_h),
              s22 = strain_center<2, 2>(vel, ix, c.inv_h);
  const float s01 = strain_center<0, 1>(vel, ix, c.inv_h), s02 = strain_center<0, 2>(vel, ix, c.inv_h),
              s12 = strain_center<1, 2>(vel, ix, c.inv_h);
  // trace(S.S) row by row like np.trace(S.dot(S)); s_ji == s_ij bitwise
  const float r0 = s00 * s00 + s01 * s01 + s02 * s02;
  const float r1 = s01 * s01 + s11 * s11 + s12 * s12;
  const float r2 = s02 * s02 + s12 * s12 + s22 * s22;
  const float tr = (r0 + r1) + r2;
  nut[boff + gid] = c.smag_coef * sqrtf(2.f * tr);
}

// nu interpolated from centres to the offset of s_IJ, at the cell shifted by p  (subgrid_models.py:94-97)
template <int I, int J>
__device__ __forceinline__ float nu_at(const float* __restrict__ nut, const Idx3& ix, int p0, int p1,
                                       int p2) {
  if (I == J) {
    int d[3] = {p0, p1, p2};
    d[I] += 1;
    return at(nut, ix, d[0], d[1], d[2]);
  }
  constexpr int A = I < J ? I : J, B = I < J ? J : I;
  float r[2];
#pragma unroll
  for (int sb = 0; sb < 2; ++sb) {
    int d0[3] = {p0, p1, p2}, d1[3] = {p0, p1, p2};
    d0[B] += sb;
    d1[B] += sb;
    d1[A] += 1;
    r[sb] = 0.5f * at(nut, ix, d0[0], d0[1], d0[2]) + 0.5f * at(nut, ix, d1[0], d1[1], d1[2]);
  }
  return 0.5f * r[0] + 0.5f * r[1];
}
// tau_IJ = -2 nu_IJ s_IJ at the cell shifted by p
template <int I, int J>
__device__ __forceinline__ float tau(const Vel3& v, const float* __restrict__ nut, const Idx3& ix,
                                     int p0, int p1, int p2, const float* inv_h) {
  return -2.f * nu_at<I, J>(nut, ix, p0, p1, p2) * strain<I, J>(v, ix, p0, p1, p2, inv_h);
}
// -(divergence of row I of tau)   (subgrid_models.py:131-134)
template <int I>
__device__ __forceinline__ float smag_acc(const Vel3& v, const float* __restrict__ nut, const Idx3& ix,
                                          const float* inv_h) {
  float d = (tau<I, 0>(v, nut, ix, 0, 0, 0, inv_h) - tau<I, 0>(v, nut, ix, -1, 0, 0, inv_h)) * inv_h[0];
  d += (tau<I, 1>(v, nut, ix, 0, 0, 0, inv_h) - tau<I, 1>(v, nut, ix, 0, -1, 0, inv_h)) * inv_h[1];
  d += (tau<I, 2>(v, nut, ix, 0, 0, 0, inv_h) - tau<I, 2>(v, nut, ix, 0, 0, -1, inv_h)) * inv_h[2];
  return -d;
}

template <int A>
__device__ __forceinline__ float explicit_comp(const Vel3& vel, const float* __restrict__ nut,
                                               const Idx3& ix, const StepConsts& c, int i, int j,
                                               int k, size_t cell) {
  const float* cc = vel.f[A];
  const float c0 = at(cc, ix, 0, 0, 0);
  float conv = conv_dir<A, 0>(cc, vel.f[0], ix, c.dth[0], c.inv_h[0]);
  conv += conv_dir<A, 1>(cc, vel.f[1], ix, c.dth[1], c.inv_h[1]);
  conv += conv_dir<A, 2>(cc, vel.f[2], ix, c.dth[2], c.inv_h[2]);
  float dv = -conv;
  if (c.has_nu) {
    float l = (-2.f * c0) * c.lap_sum;
    l += (at(cc, ix, -1, 0, 0) + at(cc, ix, 1, 0, 0)) * c.lap_s[0];
    l += (at(cc, ix, 0, -1, 0) + at(cc, ix, 0, 1, 0)) * c.lap_s[1];
    l += (at(cc, ix, 0, 0, -1) + at(cc, ix, 0, 0, 1)) * c.lap_s[2];
    dv += c.nu * l;
  }
  if (c.n_terms > 0) {
    float f = 0.f;
    for (int t = 0; t < c.n_terms; ++t) {
      const int kind = c.term_kind[t];
      if (kind == CFD_FORCE_SEPARABLE) {
        if (c.has_sep[A]) {
          float p = 1.f;
          if (c.sep_prof[A][0]) p = __ldg(c.sep_prof[A][0] + i);
          if (c.sep_prof[A][1]) p = p * __ldg(c.sep_prof[A][1] + j);
          if (c.sep_prof[A][2]) p = p * __ldg(c.sep_prof[A][2] + k);
          f += p * c.sep_scale[A];
        }
      } else if (kind == CFD_FORCE_FIELD) {
        if (c.field[A]) f += __ldg(c.field[A] + cell);
      } else if (kind == CFD_FORCE_LINEAR) {
        f += c.linear_coef * c0;
      } else if (kind == CFD_FORCE_SMAGORINSKY) {
        f += smag_acc<A>(vel, nut, ix, c.inv_h);
      }
    }
    dv = fmaf(f, c.inv_rho, dv);
  }
  return dv;
}

// one thread per cell; writes u* (or dv/dt)
__global__ void __launch_bounds__(128)
explicit3d_kernel(const float* __restrict__ u, const float* __restrict__ v,
                  const float* __restrict__ w, const float* __restrict__ nut,
                  float* __restrict__ us, float* __restrict__ vs, float* __restrict__ ws, int N0,
                  int N1, int N2, StepConsts c, int dvdt_mode) {
  const size_t cells = (size_t)N0 * N1 * N2;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells) return;
  const size_t boff = (size_t)blockIdx.y * cells;
  const int k = (int)(gid % N2), j = (int)((gid / N2) % N1), i = (int)(gid / ((size_t)N1 * N2));
  const Idx3 ix = make_idx(i, j, k, N0, N1, N2);
  Vel3 vel = {{u + boff, v + boff, w + boff}};
  const float* nu_b = nut ? nut + boff : nullptr;
  const float d0 = explicit_comp<0>(vel, nu_b, ix, c, i, j, k, gid);
  const float d1 = explicit_comp<1>(vel, nu_b, ix, c, i, j, k, gid);
  const float d2 = explicit_comp<2>(vel, nu_b, ix, c, i, j, k, gid);
  if (dvdt_mode) {
    us[boff + gid] = d0;
    vs[boff + gid] = d1;
    ws[boff + gid] = d2;
  } else {
    us[boff + gid] = __ldg(vel.f[0] + gid) + c.dt * d0;
    vs[boff + gid] = __ldg(vel.f[1] + gid) + c.dt * d1;
    ws[boff + gid] = __ldg(vel.f[2] + gid) + c.dt * d2;
  }
}

// ---------------------------------------------------------------------------------------------
// Marching 3-D explicit kernel (2.5-D blocking): a CTA of 8 warps owns a tile of BY = 8 rows (y) x
// BZ = 64 columns (z; 2 per lane, 64-bit accesses) and marches along x.  The +-2 x-stencil of the
// advected component lives in registers (5-plane window per column); y / z neighbours and the
// cross-component face velocities come from a 4-slot ring of shared-memory planes (tile + halo 2,
// all three components), one __syncthreads per plane.  x-face fluxes are carried to the next plane
// in registers, z-face fluxes are shared inside the lane (3 faces for 2 cells), y-face fluxes are
// evaluated on both sides (13.5 face evaluations per cell vs 18 in the one-thread-per-cell kernel,
// and ~2 global loads per cell instead of ~75).  Smagorinsky is added by smag_add3d_kernel.
constexpr int kBY = 8, kBZ = 64, kHY = kBY + 4, kSlots = 4;

// Plane i in [-N0, 2 N0) of a field that is slab-decomposed along axis 0 (multi_gpu_3d.cu): planes
// below 0 / from N0 on live in the previous / next rank's buffer, read over NVLink through the
// CUDA-IPC mapping.  On one GPU prev = next = own, which is the periodic wrap.
__device__ __forceinline__ const float* slab_plane(const SlabSrc& s, int i, int N0, size_t planeN) {
  const float* base = s.own;
  if (i < 0) {
    base = s.prev;
    i += N0;
  } else if (i >= N0) {
    base = s.next;
    i -= N0;
  }
  return base + (size_t)i * planeN;
}
constexpr int kHZ = kBZ + 8;  // row: [pad pad h h | 64 interior (16-byte aligned) | h h pad pad]
constexpr int kZ0 = 4;        // column of the first interior cell
constexpr int kPlane = kHY * kHZ;  // floats per component per slot

// Readers of the shared-memory plane rings (velocity, nu_t) and the strain / eddy-viscosity samples
// of subgrid_models.py:40-134 expressed through them (used by the Smagorinsky kernels below and by
// the fused stencil).
template <int AX, class R>
__device__ __forceinline__ float fwd_r(const R& rd, int comp, int p0, int p1, int p2, float inv_h) {
  int d[3] = {p0, p1, p2};
  const float a = rd(comp, d[0], d[1], d[2]);
  d[AX] += 1;
  return (rd(comp, d[0], d[1], d[2]) - a) * inv_h;
}
template <int I, int J, class R>
__device__ __forceinline__ float strain_r(const R& rd, int p0, int p1, int p2, const float* inv_h) {
  return 0.5f * (fwd_r<J>(rd, I, p0, p1, p2, inv_h[J]) + fwd_r<I>(rd, J, p0, p1, p2, inv_h[I]));
}
template <int I, int J, class R>
__device__ __forceinline__ float strain_center_r(const R& rd, const float* inv_h) {
  if (I == J) {
    int d[3] = {0, 0, 0};
    d[I] = -1;
    return strain_r<I, I>(rd, d[0], d[1], d[2], inv_h);
  }
  constexpr int A = I < J ? I : J, B = I < J ? J : I;
  float r[2];
#pragma unroll
  for (int sb = 0; sb < 2; ++sb) {
    int d0[3] = {0, 0, 0}, d1[3] = {0, 0, 0};
    d0[B] = d1[B] = sb - 1;
    d0[A] = -1;
    const float lo = strain_r<I, J>(rd, d0[0], d0[1], d0[2], inv_h);
    const float hi = strain_r<I, J>(rd, d1[0], d1[1], d1[2], inv_h);
    r[sb] = 0.5f * lo + 0.5f * hi;
  }
  return 0.5f * r[0] + 0.5f * r[1];
}
template <int I, int J, class NR>
__device__ __forceinline__ float nu_at_r(const NR& nr, int p0, int p1, int p2) {
  if (I == J) {
    int d[3] = {p0, p1, p2};
    d[I] += 1;
    return nr(d[0], d[1], d[2]);
  }
  constexpr int A = I < J ? I : J, B = I < J ? J : I;
  float r[2];
#pragma unroll
  for (int sb = 0; sb < 2; ++sb) {
    int d0[3] = {p0, p1, p2}, d1[3] = {p0, p1, p2};
    d0[B] += sb;
    d1[B] += sb;
    d1[A] += 1;
    r[sb] = 0.5f * nr(d0[0], d0[1], d0[2]) + 0.5f * nr(d1[0], d1[1], d1[2]);
  }
  return 0.5f * r[0] + 0.5f * r[1];
}

// Readers over three consecutive planes (x-1, x, x+1) of a shared-memory ring
struct VelPlanes {
  const float* pl[3];
  int own;
  __device__ __forceinline__ float operator()(int comp, int d0, int d1, int d2) const {
    return pl[d0 + 1][comp * kPlane + own + d1 * kHZ + d2];
  }
};
struct NutPlanes {
  const float* pl[3];
  int own;
  __device__ __forceinline__ float operator()(int d0, int d1, int d2) const {
    return pl[d0 + 1][own + d1 * kHZ + d2];
  }
};

// HAS_FORCE = false: no forcing term other than Smagorinsky (which the smag kernels add) -- the
// runtime walk over the term list, 6 times per plane, is compiled out.
//
// SMAG = true: the Smagorinsky acceleration -div tau (subgrid_models.py:101-134, the LAST forcing term,
// subgrid_models.py:188-213) is added in the same kernel from a second ring of nu_t planes: the
// separate smag_acc_march_kernel re-loads the three velocity planes and read-modify-writes u*
// (-36 B/cell) and is latency bound on exactly those loads.  Same arithmetic, same operation order
// (u* = (c0 + dt dv) + scale (-d)): bit-identical to the two-kernel path (CFD_SMAG_FUSED=0).
// Two CTAs per SM (117 registers, no spills): measured at 512^3, stencil + closure 3.51 ms as two
// kernels, 4.00 / 3.91 / 2.71 ms fused with 3 / 4 / 2 CTAs per SM (85 / 64 / 128-register caps).
#ifndef CFD_SMAG_MINB
#define CFD_SMAG_MINB 2
#endif
template <bool HAS_FORCE, bool SMAG = false>
__global__ void __launch_bounds__(256, SMAG ? CFD_SMAG_MINB : 4)
explicit3d_march_kernel(SlabSrc su, SlabSrc sv, SlabSrc sw, SlabSrc snut, float* __restrict__ us,
                        float* __restrict__ vs, float* __restrict__ ws, int N0, int N1, int N2,
                        StepConsts c, int dvdt_mode, int TX, int row0) {
  extern __shared__ __align__(16) float sm3[];  // [slot][comp][kHY][kHZ] (+ SMAG: [slot][kHY][kHZ] of nu_t)
  float* const smn = sm3 + kSlots * 3 * kPlane;
  const int tid = threadIdx.x, lane = tid & 31, ty = tid >> 5;
  const int tilesz = N2 / kBZ, tilesy = N1 / kBY;
  const int tz = blockIdx.x % tilesz, tyb = (blockIdx.x / tilesz) % tilesy;
  const int xb = blockIdx.x / (tilesz * tilesy);
  const size_t cells = (size_t)N0 * N1 * N2;
  const size_t boff = (size_t)blockIdx.y * cells;
  const SlabSrc f[3] = {su, sv, sw};
  float* o[3] = {us + boff, vs + boff, ws + boff};
  const int j0 = tyb * kBY, k0 = tz * kBZ;
  const int i0 = xb * TX, iend = min(i0 + TX, N0);
  const size_t planeN = (size_t)N1 * N2;
  const int own = (ty + 2) * kHZ + 2 * lane + kZ0;  // own position inside the halo tile
  // face velocities are carried as 2U = a + b; the exact factor 0.5 lives in dt/h and 1/h instead
  // (bit-identical, one multiply less per face -- see explicit_2d.cu)
  const float dthh[3] = {0.5f * c.dth[0], 0.5f * c.dth[1], 0.5f * c.dth[2]};
  const float ihh[3] = {0.5f * c.inv_h[0], 0.5f * c.inv_h[1], 0.5f * c.inv_h[2]};

  // Plane loader: threads 0..191 move one aligned float4 of the interior (12 rows x 16), threads
  // 192..239 one halo element (12 rows x 4); source offsets inside a plane are loop invariant.
  int ld_src = -1, ld_dst = 0;
  if (tid < 192) {
    const int r = tid >> 4, m = tid & 15;
    int j = j0 + r - 2;
    j = j < 0 ? j + N1 : (j >= N1 ? j - N1 : j);
    ld_src = j * N2 + k0 + 4 * m;
    ld_dst = r * kHZ + kZ0 + 4 * m;
  } else if (tid < 240) {
    const int h = tid - 192, r = h >> 2, hc = h & 3;
    int j = j0 + r - 2;
    j = j < 0 ? j + N1 : (j >= N1 ? j - N1 : j);
    int k = hc < 2 ? k0 - 2 + hc : k0 + kBZ + (hc - 2);
    k = k < 0 ? k + N2 : (k >= N2 ? k - N2 : k);
    ld_src = j * N2 + k;
    ld_dst = r * kHZ + (hc < 2 ? kZ0 - 2 + hc : kZ0 + kBZ + (hc - 2));
  }
  auto slot_base = [&](int plane) { return sm3 + ((plane + kSlots) & (kSlots - 1)) * (3 * kPlane); };
  auto nslot_base = [&](int plane) { return smn + ((plane + kSlots) & (kSlots - 1)) * kPlane; };
  auto load_plane = [&](int i) {
    float* dst = slot_base(i);
    if (tid < 192) {
#pragma unroll
      for (int comp = 0; comp < 3; ++comp)
        *reinterpret_cast<float4*>(dst + comp * kPlane + ld_dst) =
            ldg4(slab_plane(f[comp], i, N0, planeN) + boff + ld_src);
      if (SMAG)
        *reinterpret_cast<float4*>(nslot_base(i) + ld_dst) = ldg4(slab_plane(snut, i, N0, planeN) + boff + ld_src);
    } else if (tid < 240) {
#pragma unroll
      for (int comp = 0; comp < 3; ++comp)
        dst[comp * kPlane + ld_dst] = __ldg(slab_plane(f[comp], i, N0, planeN) + boff + ld_src);
      if (SMAG) nslot_base(i)[ld_dst] = __ldg(slab_plane(snut, i, N0, planeN) + boff + ld_src);
    }
  };
  // value of component `comp` at plane (pa: i, pb: i+1, pm: i-1), row j + dy, column k + dz
  const float* pa = nullptr;
  const float* pb = nullptr;
  const float* pm = nullptr;
  auto g = [&](int comp, int dx, int dy, int dz) -> float {
    const float* base = dx == 0 ? pa : (dx > 0 ? pb : pm);
    return base[comp * kPlane + own + dy * kHZ + dz];
  };

  // x-window of own columns: X[comp][plane offset + 2][col]
  float X[3][5][2];
  float Fxp[3][2];  // x-face flux at (i-1 | i), carried
  // prologue: planes i0-1, i0, i0+1 into the ring; own values of i0-2 straight from global
  load_plane(i0 - 1);
  load_plane(i0);
  load_plane(i0 + 1);
  {
#pragma unroll
    for (int comp = 0; comp < 3; ++comp) {
      const float2 t2 = __ldg(reinterpret_cast<const float2*>(
          slab_plane(f[comp], i0 - 2, N0, planeN) + boff + (size_t)(j0 + ty) * N2 + k0 + 2 * lane));
      X[comp][1][0] = t2.x;  // will become offset -2 after the first shift
      X[comp][1][1] = t2.y;
    }
  }
  __syncthreads();
  pm = slot_base(i0 - 1);
  pa = slot_base(i0);
  pb = slot_base(i0 + 1);
#pragma unroll
  for (int comp = 0; comp < 3; ++comp)
#pragma unroll
    for (int col = 0; col < 2; ++col) {
      X[comp][2][col] = g(comp, -1, 0, col);
      X[comp][3][col] = g(comp, 0, 0, col);
      X[comp][4][col] = g(comp, 1, 0, col);
    }
  // carried x-face fluxes at face (i0-1 | i0): stencil planes i0-2 .. i0+1, velocities at plane i0-1
#pragma unroll
  for (int A = 0; A < 3; ++A)
#pragma unroll
    for (int col = 0; col < 2; ++col) {
      const float ua = g(0, -1, 0, col);
      const float ub = A == 0 ? g(0, 0, 0, col) : (A == 1 ? g(0, -1, 1, col) : g(0, -1, 0, col + 1));
      const float U = ua + ub;
      Fxp[A][col] = face_flux(X[A][1][col], X[A][2][col], X[A][3][col], X[A][4][col], U, dthh[0]);
    }

  const float smag_scale = (dvdt_mode ? 1.f : c.dt) * c.inv_rho;
  for (int i = i0; i < iend; ++i) {
    pa = slot_base(i);
    pb = slot_base(i + 1);
    // SMAG reads plane i-1 as well, whose slot the load of plane i+3 reuses: nobody may start the
    // next plane's load before everyone has finished this plane
    if (SMAG && i > i0) __syncthreads();
    load_plane(i + 2);
    // shift the register window: offsets -2..+1 <- old -1..+2
#pragma unroll
    for (int comp = 0; comp < 3; ++comp)
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        X[comp][0][col] = X[comp][1][col];
        X[comp][1][col] = X[comp][2][col];
        X[comp][2][col] = X[comp][3][col];
        X[comp][3][col] = X[comp][4][col];
      }
    __syncthreads();
#pragma unroll
    for (int comp = 0; comp < 3; ++comp)
#pragma unroll
      for (int col = 0; col < 2; ++col) X[comp][4][col] = slot_base(i + 2)[comp * kPlane + own + col];

    const int jg = j0 + ty, kg = k0 + 2 * lane;
    const size_t cell0 = (size_t)i * planeN + (size_t)jg * N2 + kg;
    float sd[3][2];  // SMAG: div tau per component and column
    if (SMAG) {
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const VelPlanes rd = {{slot_base(i - 1), pa, pb}, own + col};
        const NutPlanes nr = {{nslot_base(i - 1), nslot_base(i), nslot_base(i + 1)}, own + col};
#define TAU(I, J, p0, p1, p2) (-2.f * nu_at_r<I, J>(nr, p0, p1, p2) * strain_r<I, J>(rd, p0, p1, p2, c.inv_h))
        // every distinct tau sample once (tau_ij == tau_ji), same order as smag_acc_march_kernel
        const float t00 = TAU(0, 0, 0, 0, 0), t00m = TAU(0, 0, -1, 0, 0);
        const float t11 = TAU(1, 1, 0, 0, 0), t11m = TAU(1, 1, 0, -1, 0);
        const float t22 = TAU(2, 2, 0, 0, 0), t22m = TAU(2, 2, 0, 0, -1);
        const float t01 = TAU(0, 1, 0, 0, 0), t01x = TAU(0, 1, -1, 0, 0), t01y = TAU(0, 1, 0, -1, 0);
        const float t02 = TAU(0, 2, 0, 0, 0), t02x = TAU(0, 2, -1, 0, 0), t02z = TAU(0, 2, 0, 0, -1);
        const float t12 = TAU(1, 2, 0, 0, 0), t12y = TAU(1, 2, 0, -1, 0), t12z = TAU(1, 2, 0, 0, -1);
#undef TAU
        float d0 = (t00 - t00m) * c.inv_h[0];
        d0 += (t01 - t01y) * c.inv_h[1];
        d0 += (t02 - t02z) * c.inv_h[2];
        float d1 = (t01 - t01x) * c.inv_h[0];
        d1 += (t11 - t11m) * c.inv_h[1];
        d1 += (t12 - t12z) * c.inv_h[2];
        float d2 = (t02 - t02x) * c.inv_h[0];
        d2 += (t12 - t12y) * c.inv_h[1];
        d2 += (t22 - t22m) * c.inv_h[2];
        sd[0][col] = d0;
        sd[1][col] = d1;
        sd[2][col] = d2;
      }
    }
#pragma unroll
    for (int A = 0; A < 3; ++A) {
      // unit vector of A for the face-velocity interpolation (interpolation.py:57-62)
      constexpr int ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1};
      // z faces shared by the lane's two cells: face fz between columns fz-1 and fz
      float Fz[3];
#pragma unroll
      for (int fz = 0; fz < 3; ++fz) {
        const float U = g(2, 0, 0, fz - 1) + g(2, ex[A], ey[A], fz - 1 + ez[A]);
        Fz[fz] = face_flux(g(A, 0, 0, fz - 2), g(A, 0, 0, fz - 1), g(A, 0, 0, fz), g(A, 0, 0, fz + 1), U,
                           dthh[2]);
      }
      float out[2];
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const float c0 = X[A][2][col];
        // x direction: new face (i | i+1), carried face (i-1 | i)
        const float Ux = g(0, 0, 0, col) + g(0, ex[A], ey[A], col + ez[A]);
        const float Fx = face_flux(X[A][1][col], X[A][2][col], X[A][3][col], X[A][4][col], Ux, dthh[0]);
        // y direction: faces (j | j+1) and (j-1 | j)
        const float Uyp = g(1, 0, 0, col) + g(1, ex[A], ey[A], col + ez[A]);
        const float Fyp = face_flux(g(A, 0, -1, col), c0, g(A, 0, 1, col), g(A, 0, 2, col), Uyp, dthh[1]);
        const float Uym = g(1, 0, -1, col) + g(1, ex[A], ey[A] - 1, col + ez[A]);
        const float Fym = face_flux(g(A, 0, -2, col), g(A, 0, -1, col), c0, g(A, 0, 1, col), Uym, dthh[1]);
        float conv = (Fx - Fxp[A][col]) * ihh[0];
        conv += (Fyp - Fym) * ihh[1];
        conv += (Fz[col + 1] - Fz[col]) * ihh[2];
        Fxp[A][col] = Fx;
        float dv = -conv;
        if (c.has_nu) {
          float l = c0 * c.lap_m2sum;  // (-2 c) * sum_j s_j
          l += (X[A][1][col] + X[A][3][col]) * c.lap_s[0];
          l += (g(A, 0, -1, col) + g(A, 0, 1, col)) * c.lap_s[1];
          l += (g(A, 0, 0, col - 1) + g(A, 0, 0, col + 1)) * c.lap_s[2];
          dv += c.nu * l;
        }
        if (HAS_FORCE) {
          float fsum = 0.f;
          for (int t = 0; t < c.n_terms; ++t) {
            const int kind = c.term_kind[t];
            if (kind == CFD_FORCE_SEPARABLE) {
              if (c.has_sep[A]) {
                float p = 1.f;
                if (c.sep_prof[A][0]) p = __ldg(c.sep_prof[A][0] + row0 + i);  // GLOBAL row of a slab
                if (c.sep_prof[A][1]) p = p * __ldg(c.sep_prof[A][1] + jg);
                if (c.sep_prof[A][2]) p = p * __ldg(c.sep_prof[A][2] + kg + col);
                fsum += p * c.sep_scale[A];
              }
            } else if (kind == CFD_FORCE_FIELD) {
              if (c.field[A]) fsum += __ldg(c.field[A] + cell0 + col);
            } else if (kind == CFD_FORCE_LINEAR) {
              fsum += c.linear_coef * c0;
            }  // CFD_FORCE_SMAGORINSKY: added by smag_add3d_kernel
          }
          dv = fmaf(fsum, c.inv_rho, dv);
        }
        out[col] = dvdt_mode ? dv : c0 + c.dt * dv;
        if (SMAG) out[col] += smag_scale * (-sd[A][col]);
      }
      *reinterpret_cast<float2*>(o[A] + cell0) = make_float2(out[0], out[1]);
    }
  }
}

// u* += dt * acc / rho  (or dv/dt += acc / rho): the Smagorinsky acceleration as the last forcing
// term (subgrid_models.py:188-213), evaluated from the PROJECTED input velocity.
__global__ void smag_add3d_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                  const float* __restrict__ w, const float* __restrict__ nut,
                                  float* __restrict__ us, float* __restrict__ vs,
                                  float* __restrict__ ws, int N0, int N1, int N2, StepConsts c,
                                  int dvdt_mode) {
  const size_t cells = (size_t)N0 * N1 * N2;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells) return;
  const size_t boff = (size_t)blockIdx.y * cells;
  const int k = (int)(gid % N2), j = (int)((gid / N2) % N1), i = (int)(gid / ((size_t)N1 * N2));
  const Idx3 ix = make_idx(i, j, k, N0, N1, N2);
  Vel3 vel = {{u + boff, v + boff, w + boff}};
  const float scale = (dvdt_mode ? 1.f : c.dt) * c.inv_rho;
  us[boff + gid] += scale * smag_acc<0>(vel, nut + boff, ix, c.inv_h);
  vs[boff + gid] += scale * smag_acc<1>(vel, nut + boff, ix, c.inv_h);
  ws[boff + gid] += scale * smag_acc<2>(vel, nut + boff, ix, c.inv_h);
}

// ---------------------------------------------------------------------------------------------
// Smagorinsky through strain FIELDS (what the reference itself does, subgrid_models.py:124-134):
// three streaming kernels instead of recomputing every strain sample from velocities per cell.
//   S[0..5] = s00, s11, s22, s01, s02, s12 at their natural offsets;  nu_t at centres;
//   acc_i = -sum_j (tau_ij - S(tau_ij, -1, j)) / h_j  with tau_ij = -2 nu_ij s_ij.
struct Sf {
  const float* s[6];
};
struct SfOut {
  float* s[6];
};
__device__ __forceinline__ int sidx(int i, int j) {  // field index of s_ij (symmetric)
  return i == j ? i : (i + j + 2);                   // (0,1)->3 (0,2)->4 (1,2)->5
}

__global__ void smag_strain3d_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                     const float* __restrict__ w, SfOut out, int N0, int N1, int N2,
                                     StepConsts c) {
  const size_t cells = (size_t)N0 * N1 * N2;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells) return;
  const size_t boff = (size_t)blockIdx.y * cells;
  const int k = (int)(gid % N2), j = (int)((gid / N2) % N1), i = (int)(gid / ((size_t)N1 * N2));
  const Idx3 ix = make_idx(i, j, k, N0, N1, N2);
  Vel3 vel = {{u + boff, v + boff, w + boff}};
  out.s[0][boff + gid] = strain<0, 0>(vel, ix, 0, 0, 0, c.inv_h);
  out.s[1][boff + gid] = strain<1, 1>(vel, ix, 0, 0, 0, c.inv_h);
  out.s[2][boff + gid] = strain<2, 2>(vel, ix, 0, 0, 0, c.inv_h);
  out.s[3][boff + gid] = strain<0, 1>(vel, ix, 0, 0, 0, c.inv_h);
  out.s[4][boff + gid] = strain<0, 2>(vel, ix, 0, 0, 0, c.inv_h);
  out.s[5][boff + gid] = strain<1, 2>(vel, ix, 0, 0, 0, c.inv_h);
}

// s_IJ field interpolated to the cell centre (same operation order as strain_center)
template <int I, int J>
__device__ __forceinline__ float sfield_center(const float* __restrict__ S, const Idx3& ix) {
  if (I == J) {
    int d[3] = {0, 0, 0};
    d[I] = -1;
    return at(S, ix, d[0], d[1], d[2]);
  }
  constexpr int A = I < J ? I : J, B = I < J ? J : I;
  float r[2];
#pragma unroll
  for (int sb = 0; sb < 2; ++sb) {
    int d0[3] = {0, 0, 0}, d1[3] = {0, 0, 0};
    d0[B] = d1[B] = sb - 1;
    d0[A] = -1;
    r[sb] = 0.5f * at(S, ix, d0[0], d0[1], d0[2]) + 0.5f * at(S, ix, d1[0], d1[1], d1[2]);
  }
  return 0.5f * r[0] + 0.5f * r[1];
}

__global__ void smag_nut_fields_kernel(Sf S, float* __restrict__ nut, int N0, int N1, int N2,
                                       StepConsts c) {
  const size_t cells = (size_t)N0 * N1 * N2;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells) return;
  const size_t boff = (size_t)blockIdx.y * cells;
  const int k = (int)(gid % N2), j = (int)((gid / N2) % N1), i = (int)(gid / ((size_t)N1 * N2));
  const Idx3 ix = make_idx(i, j, k, N0, N1, N2);
  const float s00 = sfield_center<0, 0>(S.s[0] + boff, ix), s11 = sfield_center<1, 1>(S.s[1] + boff, ix),
              s22 = sfield_center<2, 2>(S.s[2] + boff, ix);
  const float s01 = sfield_center<0, 1>(S.s[3] + boff, ix), s02 = sfield_center<0, 2>(S.s[4] + boff, ix),
              s12 = sfield_center<1, 2>(S.s[5] + boff, ix);
  const float r0 = s00 * s00 + s01 * s01 + s02 * s02;
  const float r1 = s01 * s01 + s11 * s11 + s12 * s12;
  const float r2 = s02 * s02 + s12 * s12 + s22 * s22;
  nut[boff + gid] = c.smag_coef * sqrtf(2.f * ((r0 + r1) + r2));
}

template <int I, int J>
__device__ __forceinline__ float tau_f(const Sf& S, size_t boff, const float* __restrict__ nut,
                                       const Idx3& ix, int p0, int p1, int p2) {
  return -2.f * nu_at<I, J>(nut, ix, p0, p1, p2) * at(S.s[I == J ? I : (I + J + 2)] + boff, ix, p0, p1, p2);
}

__global__ void smag_acc_fields_kernel(Sf S, const float* __restrict__ nut, float* __restrict__ us,
                                       float* __restrict__ vs, float* __restrict__ ws, int N0, int N1,
                                       int N2, StepConsts c, int dvdt_mode) {
  const size_t cells = (size_t)N0 * N1 * N2;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells) return;
  const size_t boff = (size_t)blockIdx.y * cells;
  const int k = (int)(gid % N2), j = (int)((gid / N2) % N1), i = (int)(gid / ((size_t)N1 * N2));
  const Idx3 ix = make_idx(i, j, k, N0, N1, N2);
  const float* nu = nut + boff;
  // every distinct tau sample once (tau_ij == tau_ji)
  const float t00 = tau_f<0, 0>(S, boff, nu, ix, 0, 0, 0), t00m = tau_f<0, 0>(S, boff, nu, ix, -1, 0, 0);
  const float t11 = tau_f<1, 1>(S, boff, nu, ix, 0, 0, 0), t11m = tau_f<1, 1>(S, boff, nu, ix, 0, -1, 0);
  const float t22 = tau_f<2, 2>(S, boff, nu, ix, 0, 0, 0), t22m = tau_f<2, 2>(S, boff, nu, ix, 0, 0, -1);
  const float t01 = tau_f<0, 1>(S, boff, nu, ix, 0, 0, 0), t01x = tau_f<0, 1>(S, boff, nu, ix, -1, 0, 0),
              t01y = tau_f<0, 1>(S, boff, nu, ix, 0, -1, 0);
  const float t02 = tau_f<0, 2>(S, boff, nu, ix, 0, 0, 0), t02x = tau_f<0, 2>(S, boff, nu, ix, -1, 0, 0),
              t02z = tau_f<0, 2>(S, boff, nu, ix, 0, 0, -1);
  const float t12 = tau_f<1, 2>(S, boff, nu, ix, 0, 0, 0), t12y = tau_f<1, 2>(S, boff, nu, ix, 0, -1, 0),
              t12z = tau_f<1, 2>(S, boff, nu, ix, 0, 0, -1);
  float d0 = (t00 - t00m) * c.inv_h[0];
  d0 += (t01 - t01y) * c.inv_h[1];
  d0 += (t02 - t02z) * c.inv_h[2];
  float d1 = (t01 - t01x) * c.inv_h[0];
  d1 += (t11 - t11m) * c.inv_h[1];
  d1 += (t12 - t12z) * c.inv_h[2];
  float d2 = (t02 - t02x) * c.inv_h[0];
  d2 += (t12 - t12y) * c.inv_h[1];
  d2 += (t22 - t22m) * c.inv_h[2];
  const float scale = (dvdt_mode ? 1.f : c.dt) * c.inv_rho;
  us[boff + gid] += scale * (-d0);
  vs[boff + gid] += scale * (-d1);
  ws[boff + gid] += scale * (-d2);
}

// ---------------------------------------------------------------------------------------------
// Smagorinsky straight from the velocity planes (2.5-D blocking, same tile / plane ring as
// explicit3d_march_kernel): two kernels, 16 + 40 B/cell of HBM traffic instead of the 116 B/cell
// of the three strain-field kernels above, and no six-field scratch.
//   smag_nut_march_kernel:  nu_t at cell centres            (subgrid_models.py:40-98)
//   smag_acc_march_kernel:  u* += scale * (-div tau)        (subgrid_models.py:101-134, 188-213)
// The arithmetic (operation order included) is that of strain<> / strain_center<> / nu_at<> above,
// read through shared-memory planes instead of global memory.
// Per-thread source / destination offsets of the plane loader (see explicit3d_march_kernel)
__device__ __forceinline__ void plane_loader_offsets(int tid, int j0, int k0, int N1, int N2,
                                                     int* ld_src, int* ld_dst) {
  *ld_src = -1;
  *ld_dst = 0;
  if (tid < 192) {
    const int r = tid >> 4, m = tid & 15;
    int j = j0 + r - 2;
    j = j < 0 ? j + N1 : (j >= N1 ? j - N1 : j);
    *ld_src = j * N2 + k0 + 4 * m;
    *ld_dst = r * kHZ + kZ0 + 4 * m;
  } else if (tid < 240) {
    const int h = tid - 192, r = h >> 2, hc = h & 3;
    int j = j0 + r - 2;
    j = j < 0 ? j + N1 : (j >= N1 ? j - N1 : j);
    int k = hc < 2 ? k0 - 2 + hc : k0 + kBZ + (hc - 2);
    k = k < 0 ? k + N2 : (k >= N2 ? k - N2 : k);
    *ld_src = j * N2 + k;
    *ld_dst = r * kHZ + (hc < 2 ? kZ0 - 2 + hc : kZ0 + kBZ + (hc - 2));
  }
}
template <int NCOMP>
__device__ __forceinline__ void load_plane_to(float* dst, const SlabSrc* f, int i, int N0, size_t planeN,
                                              size_t boff, int tid, int ld_src, int ld_dst) {
  if (tid < 192) {
#pragma unroll
    for (int comp = 0; comp < NCOMP; ++comp)
      *reinterpret_cast<float4*>(dst + comp * kPlane + ld_dst) =
          ldg4(slab_plane(f[comp], i, N0, planeN) + boff + ld_src);
  } else if (tid < 240) {
#pragma unroll
    for (int comp = 0; comp < NCOMP; ++comp)
      dst[comp * kPlane + ld_dst] = __ldg(slab_plane(f[comp], i, N0, planeN) + boff + ld_src);
  }
}

__global__ void __launch_bounds__(256)
smag_nut_march_kernel(SlabSrc su, SlabSrc sv, SlabSrc sw, float* __restrict__ nut, int N0, int N1, int N2,
                      StepConsts c, int TX) {
  extern __shared__ __align__(16) float sm3[];  // [slot][comp][kHY][kHZ]
  const int tid = threadIdx.x, lane = tid & 31, ty = tid >> 5;
  const int tilesz = N2 / kBZ, tilesy = N1 / kBY;
  const int tz = blockIdx.x % tilesz, tyb = (blockIdx.x / tilesz) % tilesy;
  const int xb = blockIdx.x / (tilesz * tilesy);
  const size_t cells = (size_t)N0 * N1 * N2;
  const size_t boff = (size_t)blockIdx.y * cells;
  const SlabSrc f[3] = {su, sv, sw};
  const int j0 = tyb * kBY, k0 = tz * kBZ;
  const int i0 = xb * TX, iend = min(i0 + TX, N0);
  const size_t planeN = (size_t)N1 * N2;
  const int own = (ty + 2) * kHZ + 2 * lane + kZ0;
  int ld_src, ld_dst;
  plane_loader_offsets(tid, j0, k0, N1, N2, &ld_src, &ld_dst);
  auto slot = [&](int plane) { return sm3 + ((plane + kSlots) & (kSlots - 1)) * (3 * kPlane); };
  auto load = [&](int i) { load_plane_to<3>(slot(i), f, i, N0, planeN, boff, tid, ld_src, ld_dst); };
  load(i0 - 1);
  load(i0);
  load(i0 + 1);
  for (int i = i0; i < iend; ++i) {
    __syncthreads();  // planes <= i+1 are visible; everyone is done with plane i-2
    load(i + 2);
    float out[2];
#pragma unroll
    for (int col = 0; col < 2; ++col) {
      const VelPlanes rd = {{slot(i - 1), slot(i), slot(i + 1)}, own + col};
      const float s00 = strain_center_r<0, 0>(rd, c.inv_h), s11 = strain_center_r<1, 1>(rd, c.inv_h),
                  s22 = strain_center_r<2, 2>(rd, c.inv_h);
      const float s01 = strain_center_r<0, 1>(rd, c.inv_h), s02 = strain_center_r<0, 2>(rd, c.inv_h),
                  s12 = strain_center_r<1, 2>(rd, c.inv_h);
      const float r0 = s00 * s00 + s01 * s01 + s02 * s02;
      const float r1 = s01 * s01 + s11 * s11 + s12 * s12;
      const float r2 = s02 * s02 + s12 * s12 + s22 * s22;
      out[col] = c.smag_coef * sqrtf(2.f * ((r0 + r1) + r2));
    }
    const size_t cell0 = (size_t)i * planeN + (size_t)(j0 + ty) * N2 + k0 + 2 * lane;
    *reinterpret_cast<float2*>(nut + boff + cell0) = make_float2(out[0], out[1]);
  }
}

__global__ void __launch_bounds__(256)
smag_acc_march_kernel(SlabSrc su, SlabSrc sv, SlabSrc sw, SlabSrc snut,
                      float* __restrict__ us, float* __restrict__ vs, float* __restrict__ ws, int N0,
                      int N1, int N2, StepConsts c, int dvdt_mode, int TX) {
  extern __shared__ __align__(16) float sm3[];  // velocity ring [slot][3][plane], then nu_t ring
  float* smn = sm3 + kSlots * 3 * kPlane;
  const int tid = threadIdx.x, lane = tid & 31, ty = tid >> 5;
  const int tilesz = N2 / kBZ, tilesy = N1 / kBY;
  const int tz = blockIdx.x % tilesz, tyb = (blockIdx.x / tilesz) % tilesy;
  const int xb = blockIdx.x / (tilesz * tilesy);
  const size_t cells = (size_t)N0 * N1 * N2;
  const size_t boff = (size_t)blockIdx.y * cells;
  const SlabSrc f[3] = {su, sv, sw};
  const SlabSrc fn[1] = {snut};
  float* o[3] = {us + boff, vs + boff, ws + boff};
  const int j0 = tyb * kBY, k0 = tz * kBZ;
  const int i0 = xb * TX, iend = min(i0 + TX, N0);
  const size_t planeN = (size_t)N1 * N2;
  const int own = (ty + 2) * kHZ + 2 * lane + kZ0;
  int ld_src, ld_dst;
  plane_loader_offsets(tid, j0, k0, N1, N2, &ld_src, &ld_dst);
  auto slot = [&](int plane) { return sm3 + ((plane + kSlots) & (kSlots - 1)) * (3 * kPlane); };
  auto nslot = [&](int plane) { return smn + ((plane + kSlots) & (kSlots - 1)) * kPlane; };
  auto load = [&](int i) {
    load_plane_to<3>(slot(i), f, i, N0, planeN, boff, tid, ld_src, ld_dst);
    load_plane_to<1>(nslot(i), fn, i, N0, planeN, boff, tid, ld_src, ld_dst);
  };
  load(i0 - 1);
  load(i0);
  load(i0 + 1);
  const float scale = (dvdt_mode ? 1.f : c.dt) * c.inv_rho;
  for (int i = i0; i < iend; ++i) {
    __syncthreads();
    load(i + 2);
    const size_t cell0 = (size_t)i * planeN + (size_t)(j0 + ty) * N2 + k0 + 2 * lane;
    float2 acc[3];
#pragma unroll
    for (int A = 0; A < 3; ++A) acc[A] = *reinterpret_cast<const float2*>(o[A] + cell0);
    float d[3][2];
#pragma unroll
    for (int col = 0; col < 2; ++col) {
      const VelPlanes rd = {{slot(i - 1), slot(i), slot(i + 1)}, own + col};
      const NutPlanes nr = {{nslot(i - 1), nslot(i), nslot(i + 1)}, own + col};
#define TAU(I, J, p0, p1, p2) (-2.f * nu_at_r<I, J>(nr, p0, p1, p2) * strain_r<I, J>(rd, p0, p1, p2, c.inv_h))
      // every distinct tau sample once (tau_ij == tau_ji), same order as smag_acc_fields_kernel
      const float t00 = TAU(0, 0, 0, 0, 0), t00m = TAU(0, 0, -1, 0, 0);
      const float t11 = TAU(1, 1, 0, 0, 0), t11m = TAU(1, 1, 0, -1, 0);
      const float t22 = TAU(2, 2, 0, 0, 0), t22m = TAU(2, 2, 0, 0, -1);
      const float t01 = TAU(0, 1, 0, 0, 0), t01x = TAU(0, 1, -1, 0, 0), t01y = TAU(0, 1, 0, -1, 0);
      const float t02 = TAU(0, 2, 0, 0, 0), t02x = TAU(0, 2, -1, 0, 0), t02z = TAU(0, 2, 0, 0, -1);
      const float t12 = TAU(1, 2, 0, 0, 0), t12y = TAU(1, 2, 0, -1, 0), t12z = TAU(1, 2, 0, 0, -1);
#undef TAU
      float d0 = (t00 - t00m) * c.inv_h[0];
      d0 += (t01 - t01y) * c.inv_h[1];
      d0 += (t02 - t02z) * c.inv_h[2];
      float d1 = (t01 - t01x) * c.inv_h[0];
      d1 += (t11 - t11m) * c.inv_h[1];
      d1 += (t12 - t12z) * c.inv_h[2];
      float d2 = (t02 - t02x) * c.inv_h[0];
      d2 += (t12 - t12y) * c.inv_h[1];
      d2 += (t22 - t22m) * c.inv_h[2];
      d[0][col] = d0;
      d[1][col] = d1;
      d[2][col] = d2;
    }
#pragma unroll
    for (int A = 0; A < 3; ++A) {
      acc[A].x += scale * (-d[A][0]);
      acc[A].y += scale * (-d[A][1]);
      *reinterpret_cast<float2*>(o[A] + cell0) = acc[A];
    }
  }
}

// diagnostics: sum 0.5|v|^2, sum 0.5|curl|^2, max|div|, max|v|^2
__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float nanmax_f(float a, float b) { return (a > b || a != a) ? a : b; }  // jnp.max
__device__ __forceinline__ float warp_max_f(float v) {
  for (int o = 16; o > 0; o >>= 1) v = nanmax_f(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void atomic_max_d(double* addr, double val) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    const double cur = __longlong_as_double(assumed);
    if (cur != cur || (cur >= val && val == val)) break;  // NaN sticks
    old = atomicCAS(a, assumed, __double_as_longlong(val));
  } while (assumed != old);
}
__global__ void diag3d_kernel(const float* __restrict__ u, const float* __restrict__ v,
                              const float* __restrict__ w, int N0, int N1, int N2, size_t total,
                              float ihx, float ihy, float ihz, double* __restrict__ out4) {
  const size_t cells = (size_t)N0 * N1 * N2;
  double ke = 0.0, ens = 0.0;
  float mdiv = 0.f, msp = 0.f;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = idx / cells, gid = idx % cells;
    const int k = (int)(gid % N2), j = (int)((gid / N2) % N1), i = (int)(gid / ((size_t)N1 * N2));
    const Idx3 ix = make_idx(i, j, k, N0, N1, N2);
    const float* ub = u + b * cells;
    const float* vb = v + b * cells;
    const float* wb = w + b * cells;
    const float u0 = at(ub, ix, 0, 0, 0), v0 = at(vb, ix, 0, 0, 0), w0 = at(wb, ix, 0, 0, 0);
    const float div = (u0 - at(ub, ix, -1, 0, 0)) * ihx + (v0 - at(vb, ix, 0, -1, 0)) * ihy +
                      (w0 - at(wb, ix, 0, 0, -1)) * ihz;
    // curl by forward differences (finite_differences.curl_3d convention)
    const float cx = (at(wb, ix, 0, 1, 0) - w0) * ihy - (at(vb, ix, 0, 0, 1) - v0) * ihz;
    const float cy = (at(ub, ix, 0, 0, 1) - u0) * ihz - (at(wb, ix, 1, 0, 0) - w0) * ihx;
    const float cz = (at(vb, ix, 1, 0, 0) - v0) * ihx - (at(ub, ix, 0, 1, 0) - u0) * ihy;
    const float sp = u0 * u0 + v0 * v0 + w0 * w0;
    ke += 0.5 * (double)sp;
    ens += 0.5 * ((double)cx * cx + (double)cy * cy + (double)cz * cz);
    mdiv = nanmax_f(fabsf(div), mdiv);
    msp = nanmax_f(sp, msp);
  }
  ke = warp_sum_d(ke);
  ens = warp_sum_d(ens);
  mdiv = warp_max_f(mdiv);
  msp = warp_max_f(msp);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out4 + 0, ke);
    atomicAdd(out4 + 1, ens);
    atomic_max_d(out4 + 2, (double)mdiv);
    atomic_max_d(out4 + 3, (double)msp);
  }
}

}  // namespace

bool explicit_3d_uses_march(int N0, int N1, int N2);

// x rows per CTA of the marching kernels: long enough to amortise the 3-plane prologue, short
// enough to fill the GPU
static int march_tx(int batch, int N0, int N1, int N2) {
  static const int forced = [] { const char* e = getenv("CFD_MARCH_TX"); return e ? atoi(e) : 0; }();
  if (forced >= 4) return forced;
  int TX = 32;
  while (TX > 8 && (long)(N1 / kBY) * (N2 / kBZ) * ((N0 + TX - 1) / TX) * batch < 148L * 4) TX /= 2;
  return TX;
}

// CFD_SMAG_TILED=0 selects the strain-field kernels instead of the plane-marching ones
bool smag_uses_tiles(int N0, int N1, int N2) {
  static const int v = [] {
    const char* e = getenv("CFD_SMAG_TILED");
    return e ? atoi(e) : 1;
  }();
  return v && explicit_3d_uses_march(N0, N1, N2);
}

// nu_t of a slab (or, with prev = next = own, of a whole periodic grid) by the plane-marching kernel
int launch_smag_nut_3d_slab(cudaStream_t st, SlabSrc su, SlabSrc sv, SlabSrc sw, float* nut, int batch, int N0,
                            int N1, int N2, const StepConsts& c) {
  if (!explicit_3d_uses_march(N0, N1, N2)) return set_error_msg("slab needs N1 % 8 == 0, N2 % 64 == 0, >= 4 planes");
  const int TX = march_tx(batch, N0, N1, N2);
  constexpr size_t smem = (size_t)kSlots * 3 * kPlane * sizeof(float);
  if (int e = opt_in_smem(smag_nut_march_kernel, smem)) return e;
  dim3 g((unsigned)((N1 / kBY) * (N2 / kBZ) * ((N0 + TX - 1) / TX)), batch);
  smag_nut_march_kernel<<<g, 256, smem, st>>>(su, sv, sw, nut, N0, N1, N2, c, TX);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

// sfield != nullptr: 6 strain fields of batch * cells floats each (strain-field path)
int launch_smag_nut_3d(cudaStream_t st, const float* u, const float* v, const float* w, float* nut,
                       float* sfield, int batch, int N0, int N1, int N2, const StepConsts& c) {
  const size_t cells = (size_t)N0 * N1 * N2;
  if (!sfield && smag_uses_tiles(N0, N1, N2)) {
    const SlabSrc su = {u, u, u}, sv = {v, v, v}, sw = {w, w, w};
    return launch_smag_nut_3d_slab(st, su, sv, sw, nut, batch, N0, N1, N2, c);
  }
  dim3 grid((unsigned)((cells + 127) / 128), batch);
  if (sfield) {
    SfOut so;
    Sf si;
    for (int q = 0; q < 6; ++q) {
      so.s[q] = sfield + (size_t)q * batch * cells;
      si.s[q] = so.s[q];
    }
    smag_strain3d_kernel<<<grid, 128, 0, st>>>(u, v, w, so, N0, N1, N2, c);
    count_launch();
    smag_nut_fields_kernel<<<grid, 128, 0, st>>>(si, nut, N0, N1, N2, c);
    count_launch();
    CFD_CUDA_OK(cudaGetLastError());
    return 0;
  }
  smag_nut3d_kernel<<<grid, 128, 0, st>>>(u, v, w, nut, N0, N1, N2, c);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

bool explicit_3d_uses_march(int N0, int N1, int N2) {
  static const int use_march = [] {
    const char* e = getenv("CFD_EXPLICIT3D_MARCH");
    return e ? atoi(e) : 1;
  }();
  return use_march && N1 % kBY == 0 && N2 % kBZ == 0 && N0 >= 4;
}

// CFD_SMAG_FUSED=0: -div tau by the separate smag_acc_march_kernel instead of inside the stencil
bool smag_fused() {
  static const int v = [] {
    const char* e = getenv("CFD_SMAG_FUSED");
    return e ? atoi(e) : 1;
  }();
  return v != 0;
}

// the marching stencil on a slab (row0 = first GLOBAL plane of the slab, for the forcing profiles);
// snut != nullptr: the Smagorinsky acceleration is added in the same kernel from these nu_t planes
int launch_explicit_3d_slab(cudaStream_t st, SlabSrc su, SlabSrc sv, SlabSrc sw, float* us, float* vs, float* ws,
                            int batch, int N0, int N1, int N2, const StepConsts& c, int dvdt_mode, int row0,
                            const SlabSrc* snut) {
  if (!explicit_3d_uses_march(N0, N1, N2)) return set_error_msg("slab needs N1 % 8 == 0, N2 % 64 == 0, >= 4 planes");
  const int TX = march_tx(batch, N0, N1, N2);
  constexpr size_t smem = (size_t)kSlots * 3 * kPlane * sizeof(float);
  constexpr size_t smem_smag = (size_t)kSlots * 4 * kPlane * sizeof(float);
  dim3 grid((unsigned)((N1 / kBY) * (N2 / kBZ) * ((N0 + TX - 1) / TX)), batch);
  bool has_force = false;  // a non-zero forcing sum (an all-Smagorinsky list adds +0 here)
  for (int t = 0; t < c.n_terms; ++t) has_force = has_force || c.term_kind[t] != CFD_FORCE_SMAGORINSKY;
  const SlabSrc none = {nullptr, nullptr, nullptr};
  auto ready = [&](auto k, size_t bytes) { return opt_in_smem(k, bytes); };
  if (int e = snut ? (has_force ? ready(explicit3d_march_kernel<true, true>, smem_smag)
                               : ready(explicit3d_march_kernel<false, true>, smem_smag))
                   : (has_force ? ready(explicit3d_march_kernel<true, false>, smem)
                               : ready(explicit3d_march_kernel<false, false>, smem)))
    return e;
  if (snut) {
    if (has_force)
      explicit3d_march_kernel<true, true><<<grid, 256, smem_smag, st>>>(su, sv, sw, *snut, us, vs, ws, N0, N1, N2, c,
                                                                        dvdt_mode, TX, row0);
    else
      explicit3d_march_kernel<false, true><<<grid, 256, smem_smag, st>>>(su, sv, sw, *snut, us, vs, ws, N0, N1, N2, c,
                                                                         dvdt_mode, TX, row0);
  } else if (has_force) {
    explicit3d_march_kernel<true, false><<<grid, 256, smem, st>>>(su, sv, sw, none, us, vs, ws, N0, N1, N2, c,
                                                                  dvdt_mode, TX, row0);
  } else {
    explicit3d_march_kernel<false, false><<<grid, 256, smem, st>>>(su, sv, sw, none, us, vs, ws, N0, N1, N2, c,
                                                                   dvdt_mode, TX, row0);
  }
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

// u* += scale * (-div tau) on a slab: velocity and nu_t halo planes come from the neighbouring ranks
int launch_smag_acc_3d_slab(cudaStream_t st, SlabSrc su, SlabSrc sv, SlabSrc sw, SlabSrc snut, float* us,
                            float* vs, float* ws, int batch, int N0, int N1, int N2, const StepConsts& c,
                            int dvdt_mode) {
  if (!explicit_3d_uses_march(N0, N1, N2)) return set_error_msg("slab needs N1 % 8 == 0, N2 % 64 == 0, >= 4 planes");
  const int TX = march_tx(batch, N0, N1, N2);
  constexpr size_t smem2 = (size_t)kSlots * 4 * kPlane * sizeof(float);
  if (int e = opt_in_smem(smag_acc_march_kernel, smem2)) return e;
  dim3 grid((unsigned)((N1 / kBY) * (N2 / kBZ) * ((N0 + TX - 1) / TX)), batch);
  smag_acc_march_kernel<<<grid, 256, smem2, st>>>(su, sv, sw, snut, us, vs, ws, N0, N1, N2, c, dvdt_mode, TX);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_explicit_3d(cudaStream_t st, const float* u, const float* v, const float* w,
                       const float* nut, const float* sfield, float* us, float* vs, float* ws,
                       int batch, int N0, int N1, int N2, const StepConsts& c, int dvdt_mode) {
  const size_t cells = (size_t)N0 * N1 * N2;
  if (explicit_3d_uses_march(N0, N1, N2)) {
    const SlabSrc su = {u, u, u}, sv = {v, v, v}, sw = {w, w, w};
    const SlabSrc sn = {nut, nut, nut};
    const bool tiled_smag = nut && !sfield && smag_uses_tiles(N0, N1, N2);
    const bool fused = tiled_smag && smag_fused();
    if (int e = launch_explicit_3d_slab(st, su, sv, sw, us, vs, ws, batch, N0, N1, N2, c, dvdt_mode, 0,
                                        fused ? &sn : nullptr))
      return e;
    if (fused) {
      // -div tau was added inside the stencil
    } else if (tiled_smag) {
      if (int e = launch_smag_acc_3d_slab(st, su, sv, sw, sn, us, vs, ws, batch, N0, N1, N2, c, dvdt_mode)) return e;
    } else if (nut && sfield) {
      dim3 g2((unsigned)((cells + 127) / 128), batch);
      Sf si;
      for (int q = 0; q < 6; ++q) si.s[q] = sfield + (size_t)q * batch * cells;
      smag_acc_fields_kernel<<<g2, 128, 0, st>>>(si, nut, us, vs, ws, N0, N1, N2, c, dvdt_mode);
      count_launch();
      CFD_CUDA_OK(cudaGetLastError());
    } else if (nut) {
      dim3 g2((unsigned)((cells + 127) / 128), batch);
      smag_add3d_kernel<<<g2, 128, 0, st>>>(u, v, w, nut, us, vs, ws, N0, N1, N2, c, dvdt_mode);
      count_launch();
      CFD_CUDA_OK(cudaGetLastError());
    }
    return 0;
  }
  dim3 grid((unsigned)((cells + 127) / 128), batch);
  explicit3d_kernel<<<grid, 128, 0, st>>>(u, v, w, nut, us, vs, ws, N0, N1, N2, c, dvdt_mode);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_diag_3d(cudaStream_t st, const float* u, const float* v, const float* w, int batch, int N0,
                   int N1, int N2, float ihx, float ihy, float ihz, double* out4) {
  CFD_CUDA_OK(cudaMemsetAsync(out4, 0, 4 * sizeof(double), st));
  const size_t total = (size_t)batch * N0 * N1 * N2;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  diag3d_kernel<<<(int)blocks, 256, 0, st>>>(u, v, w, N0, N1, N2, total, ihx, ihy, ihz, out4);
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfd
