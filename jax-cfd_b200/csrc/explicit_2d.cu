// Sweep "E" (2-D): fused explicit terms of the momentum equation + forward-Euler update + divergence.
//
//   u* = v + dt * ( -div(flux_vanleer) + (nu/rho) lap(v) + forcing/rho ),   rhs = div(u*)
//
// Replaces, in one pass over HBM (8 B/cell read, 12 B/cell written):
//   advection.advect_van_leer_using_limiters  advection.py:387-395 -> 81-116 -> 34-78
//   interpolation.linear / upwind / lax_wendroff / apply_tvd_limiter  interpolation.py:36-303
//   diffusion.diffuse -> finite_differences.laplacian   diffusion.py:35-37, finite_differences.py:127-133
//   forcings.{kolmogorov,taylor_green,linear,sum}_forcing   forcings.py:35-129
//   equations.navier_stokes_explicit_terms   equations.py:102-114   (order conv + diff + force/rho)
//   time_stepping.navier_stokes_rk   time_stepping.py:101     (u* = u0 + dt * k0)
//   finite_differences.divergence   finite_differences.py:136-143   (rhs of pressure.py:147)
//
// Layout: a warp owns a strip of 32*C consecutive columns (C = 4 or 2 per lane, 128/64-bit
// coalesced loads) and marches down the rows keeping a 5-row window of u and v in REGISTERS; the
// +-2 column halo of the current row comes from the neighbouring lanes by warp shuffle.  Lanes 0
// and 31 are halo lanes (their results are not stored), so a warp produces 30*C columns.  x-face
// fluxes are computed once and carried to the next row in registers; y-face fluxes are computed
// once per lane (C+1 faces for C cells).  No shared memory, no block-level synchronisation.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace cfd {

namespace {

constexpr int kWarpsPerCta = 4;

// C consecutive columns owned by one lane
template <int C>
struct FC {
  float a[C];
};
template <int C>
__device__ __forceinline__ FC<C> ldrow(const float* p);
template <>
__device__ __forceinline__ FC<4> ldrow<4>(const float* p) {
  const float4 v = ldg4(p);
  return FC<4>{{v.x, v.y, v.z, v.w}};
}
template <>
__device__ __forceinline__ FC<2> ldrow<2>(const float* p) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return FC<2>{{v.x, v.y}};
}
template <>
__device__ __forceinline__ FC<8> ldrow<8>(const float* p) {
  const float4 a = ldg4(p), b = ldg4(p + 4);
  return FC<8>{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ void strow(float* p, const FC<8>& f) {
  stg4(p, make_float4(f.a[0], f.a[1], f.a[2], f.a[3]));
  stg4(p + 4, make_float4(f.a[4], f.a[5], f.a[6], f.a[7]));
}
__device__ __forceinline__ void strow(float* p, const FC<4>& f) {
  stg4(p, make_float4(f.a[0], f.a[1], f.a[2], f.a[3]));
}
__device__ __forceinline__ void strow(float* p, const FC<2>& f) {
  *reinterpret_cast<float2*>(p) = make_float2(f.a[0], f.a[1]);
}

#define FULLMASK 0xffffffffu

// PATTERN encodes the (compile-time) sequence of forcing terms, 2 bits per term in summation
// order: 0 = end, 1 = separable, 2 = field, 3 = linear.
//
// LAZY: the inputs are the UNPROJECTED state of the previous step (u*, v*) plus its pressure q;
// the projection  v = u* - forward_difference(q)  (pressure.py:194-196) is applied while the
// window rows are loaded, so that chained steps never materialise the projected state in HBM.
//
// Slab decomposition (multi-GPU): Nx is the number of LOCAL rows; rows -3..-1 and Nx..Nx+2 are read
// straight from the neighbouring ranks' buffers over NVLink (SlabSrc::prev / next, peer-mapped
// with CUDA IPC) -- the halo exchange is fused into the stencil loads.  On one GPU prev = next =
// own, which is the periodic wrap.
//
// WRAP: the row is exactly 32 * C columns long, so ONE warp owns whole rows and the periodic column
// halo is the wrap-around of the lane index -- no halo lanes, no redundant columns, every y-face
// flux evaluated by a lane is used.  With C = 8 this is the 256-column row of the ensemble members
// (BASELINE config #3), where halo lanes waste a third of the machine (3 x 128 lane-columns for 256
// columns).  The four warps of a CTA then take four consecutive row tiles.
template <int PATTERN, int C, bool LAZY, bool WRAP = false>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
explicit2d_kernel(SlabSrc su, SlabSrc sv, SlabSrc sq, float* __restrict__ us,
                  float* __restrict__ vs, float* __restrict__ rhs, int Nx, int Ny, int row0,
                  int nx_global, StepConsts c, int dvdt_mode, int TX, int tile_begin) {
  using Row = FC<C>;
  // Halo lanes: the divergence needs v* one column to the left of the first stored column, whose
  // own stencil reaches 2 further columns: 4 columns = 1 lane (C = 4) or 2 lanes (C = 2) on the
  // left, 1 lane on the right.
  // (LAZY adds one column on the right: projecting v[j] needs q[j+1]; C = 2 then needs 2 lanes.)
  static_assert(WRAP || C <= 4, "more than 4 columns per lane only in WRAP mode");
  constexpr int kHaloL = WRAP ? 0 : 4 / C;
  constexpr int kHaloR = WRAP ? 0 : ((C == 2) ? 2 : 1);
  constexpr int kWarpCols = (32 - kHaloL - kHaloR) * C;  // stored columns per warp: 120 or 56 (WRAP: all)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int strip = WRAP ? 0 : blockIdx.x * kWarpsPerCta + warp;
  if (strip * kWarpCols >= Ny) return;
  const int jbase = strip * kWarpCols - kHaloL * C + C * lane;
  int jg = jbase % Ny;
  if (jg < 0) jg += Ny;
  const bool store_ok = WRAP || ((lane >= kHaloL) && (lane < 32 - kHaloR) && (jbase < Ny));
  const size_t boff = (size_t)blockIdx.z * (size_t)Nx * (size_t)Ny;
  // row tiles [tile_begin, tile_begin + tiles of this launch); WRAP: one tile per warp
  const int i0 = ((WRAP ? blockIdx.y * kWarpsPerCta + warp : blockIdx.y) + tile_begin) * TX;
  if (WRAP && i0 >= Nx) return;
  const int iend = min(i0 + TX, Nx);  // exclusive
  // value of the previous / next lane; WRAP: around the row
  auto from_left = [&](float x) {
    return WRAP ? __shfl_sync(FULLMASK, x, (lane + 31) & 31) : __shfl_up_sync(FULLMASK, x, 1);
  };
  auto from_right = [&](float x) {
    return WRAP ? __shfl_sync(FULLMASK, x, (lane + 1) & 31) : __shfl_down_sync(FULLMASK, x, 1);
  };

  // row i in [-Nx, 2 Nx) of a slab-decomposed field
  auto rowptr = [&](const SlabSrc& f, int i) -> const float* {
    const float* base = f.own;
    if (i < 0) {
      base = f.prev;
      i += Nx;
    } else if (i >= Nx) {
      base = f.next;
      i -= Nx;
    }
    return base + boff + (size_t)i * Ny + jg;
  };

  // The face velocity U = 0.5 (a + b) (interpolation.py:57-62) is carried as 2U = a + b: the factor
  // 0.5 is a power of two, so folding it into dt/h (Courant number) and into 1/h (flux divergence)
  // gives bit-identical results (away from underflow) and saves one multiply per face.
  const float dth0h = 0.5f * c.dth[0], dth1h = 0.5f * c.dth[1];
  const float ih0h = 0.5f * c.inv_h[0], ih1h = 0.5f * c.inv_h[1];

  // window rows i-2 .. i+2 for the first processed row i = i0 - 1
  Row ua[5], va[5];
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    ua[r] = ldrow<C>(rowptr(su, i0 - 3 + r));
    va[r] = ldrow<C>(rowptr(sv, i0 - 3 + r));
  }
  // lazily projected input: row r of (u, v) = (u*, v*)[r] - grad q, needs q rows r and r+1 and
  // the column to the right (from lane+1; lane 31's last column is never used by lane 30)
  auto project_row = [&](Row& ur, Row& vr, const Row& q0, const Row& q1) {
    const float qR = from_right(q0.a[0]);
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const float qright = (k == C - 1) ? qR : q0.a[(k + 1) % C];
      ur.a[k] = ur.a[k] - (q1.a[k] - q0.a[k]) * c.inv_h[0];
      vr.a[k] = vr.a[k] - (qright - q0.a[k]) * c.inv_h[1];
    }
  };
  Row qkeep;  // q row i+3 relative to the first loop iteration (= row i0+2)
#pragma unroll
  for (int k = 0; k < C; ++k) qkeep.a[k] = 0.f;
  if (LAZY) {
    Row qa[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) qa[r] = ldrow<C>(rowptr(sq, i0 - 3 + r));
#pragma unroll
    for (int r = 0; r < 5; ++r) project_row(ua[r], va[r], qa[r], qa[r + 1]);
    qkeep = qa[5];
  }
  // x-face fluxes at face (i0-2 | i0-1), needed by row i0-1:  stencil rows i0-3 .. i0
  Row f0u_prev, f0v_prev;
  {
    const float uR1 = from_right(ua[1].a[0]);
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const float Uu = ua[1].a[k] + ua[2].a[k];
      f0u_prev.a[k] = face_flux(ua[0].a[k], ua[1].a[k], ua[2].a[k], ua[3].a[k], Uu, dth0h);
      const float unext = (k < C - 1) ? ua[1].a[(k + 1) % C] : uR1;
      const float Uv = ua[1].a[k] + unext;
      f0v_prev.a[k] = face_flux(va[0].a[k], va[1].a[k], va[2].a[k], va[3].a[k], Uv, dth0h);
    }
  }
  Row us_prev;
#pragma unroll
  for (int k = 0; k < C; ++k) us_prev.a[k] = 0.f;
  float vL1_cur = from_left(va[2].a[C - 1]);  // v[i][-1] for i = i0-1


  // forcing tables: the column profiles of the separable term are fixed per thread
  constexpr bool kHasSep = ((PATTERN & 3) == 1) || (((PATTERN >> 2) & 3) == 1) || (((PATTERN >> 4) & 3) == 1);
  constexpr bool kHasField = ((PATTERN & 3) == 2) || (((PATTERN >> 2) & 3) == 2) || (((PATTERN >> 4) & 3) == 2);
  const float* px_u = c.sep_prof[0][0];
  const float* px_v = c.sep_prof[1][0];
  float pyu[C], pyv[C];
#pragma unroll
  for (int k = 0; k < C; ++k) pyu[k] = pyv[k] = 1.f;
  float scale_u = 0.f, scale_v = 0.f;
  if (kHasSep) {
#pragma unroll
    for (int k = 0; k < C; ++k) {
      if (c.sep_prof[0][1]) pyu[k] = __ldg(c.sep_prof[0][1] + jg + k);
      if (c.sep_prof[1][1]) pyv[k] = __ldg(c.sep_prof[1][1] + jg + k);
    }
    scale_u = c.has_sep[0] ? c.sep_scale[0] : 0.f;
    scale_v = c.has_sep[1] ? c.sep_scale[1] : 0.f;
    // a component the term does not force contributes +0
    if (!c.has_sep[0]) {
#pragma unroll
      for (int k = 0; k < C; ++k) pyu[k] = 0.f;
      px_u = nullptr;
    }
    if (!c.has_sep[1]) {
#pragma unroll
      for (int k = 0; k < C; ++k) pyv[k] = 0.f;
      px_v = nullptr;
    }
  }
  // Row profiles of the separable term for the rows of this block, one row per lane (3 x 32 >= TX
  // + 1), broadcast by shuffle inside the loop: no scalar global loads on the critical path.
  const bool has_px = kHasSep && (px_u != nullptr || px_v != nullptr);
  float pxu_tab[3] = {1.f, 1.f, 1.f}, pxv_tab[3] = {1.f, 1.f, 1.f};
  if (has_px) {
#pragma unroll
    for (int sgm = 0; sgm < 3; ++sgm) {
      int iw = (row0 + i0 - 1 + 32 * sgm + lane) % nx_global;  // GLOBAL row of the profile
      if (iw < 0) iw += nx_global;
      if (px_u) pxu_tab[sgm] = __ldg(px_u + iw);
      if (px_v) pxv_tab[sgm] = __ldg(px_v + iw);
    }
  }
  // constant-field forcing: rows are prefetched one iteration ahead
  Row nfu, nfv;
#pragma unroll
  for (int k = 0; k < C; ++k) nfu.a[k] = nfv.a[k] = 0.f;
  if (kHasField) {
    // (field forcing is single-GPU only: rows wrap inside the local array)
    const int iwf = (i0 - 1 + Nx) % Nx;
    if (c.field[0]) nfu = ldrow<C>(c.field[0] + (size_t)iwf * Ny + jg);
    if (c.field[1]) nfv = ldrow<C>(c.field[1] + (size_t)iwf * Ny + jg);
  }

  // two rows per trip in the lazy kernel: half of the window-rotation moves disappear (423 -> 407 us
  // at 8192^2; the plain kernel gets slower with it, 407 -> 421 us, and keeps one row per trip)
#ifndef CFD_WRAP_UNROLL
#define CFD_WRAP_UNROLL 1
#endif
  constexpr int kRowUnroll = LAZY ? (C > 4 ? CFD_WRAP_UNROLL : 2) : 1;
#pragma unroll kRowUnroll
  for (int i = i0 - 1; i < iend; ++i) {
    // ---- column halos of row i (and v[i+1][-1]) from neighbouring lanes
    float ue[C + 4], ve[C + 4];
    ue[0] = from_left(ua[2].a[C - 2]);
    ue[1] = from_left(ua[2].a[C - 1]);
    ue[C + 2] = from_right(ua[2].a[0]);
    ue[C + 3] = from_right(ua[2].a[1]);
    ve[0] = from_left(va[2].a[C - 2]);
    ve[1] = vL1_cur;
    ve[C + 2] = from_right(va[2].a[0]);
    ve[C + 3] = from_right(va[2].a[1]);
    const float vnL1 = from_left(va[3].a[C - 1]);  // v[i+1][-1]
#pragma unroll
    for (int k = 0; k < C; ++k) {
      ue[2 + k] = ua[2].a[k];
      ve[2 + k] = va[2].a[k];
    }
    // Prefetch row i+3 (the new window row of the NEXT iteration) only now, after the shuffles:
    // the loads then do not occupy scoreboard slots while the shuffles need them, and the rest of
    // the body (~500 instructions) covers their latency.
    Row nu, nv, nq;
#pragma unroll
    for (int k = 0; k < C; ++k) nu.a[k] = nv.a[k] = nq.a[k] = 0.f;
    if (i + 1 < iend) {
      nu = ldrow<C>(rowptr(su, i + 3));
      nv = ldrow<C>(rowptr(sv, i + 3));
      if (LAZY) nq = ldrow<C>(rowptr(sq, i + 4));
    }

    // ---- face fluxes
    Row f0u, f0v;                  // x-faces (i | i+1) at the lane's columns
    float f1u[C + 1], f1v[C + 1];  // y-faces (j' | j'+1), j' = -1 .. C-1
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const float Uu = ua[2].a[k] + ua[3].a[k];  // 2 U, interpolation.py:57-62
      f0u.a[k] = face_flux(ua[1].a[k], ua[2].a[k], ua[3].a[k], ua[4].a[k], Uu, dth0h);
      const float Uv = ue[2 + k] + ue[3 + k];
      f0v.a[k] = face_flux(va[1].a[k], va[2].a[k], va[3].a[k], va[4].a[k], Uv, dth0h);
    }
#pragma unroll
    for (int f = 0; f < C + 1; ++f) {  // face j' = f - 1 ; stencil columns j'-1..j'+2 = ext[f .. f+3]
      const float vnext = (f == 0) ? vnL1 : va[3].a[(f + C - 1) % C];  // v[i+1][j']
      const float Uu = ve[f + 1] + vnext;
      f1u[f] = face_flux(ue[f], ue[f + 1], ue[f + 2], ue[f + 3], Uu, dth1h);
      const float Uv = ve[f + 1] + ve[f + 2];
      f1v[f] = face_flux(ve[f], ve[f + 1], ve[f + 2], ve[f + 3], Uv, dth1h);
    }

    // ---- assemble
    const int iw = i < 0 ? i + Nx : i;  // local row (the warm-up row -1 is never stored)
    Row us_cur, vs_cur;
    float pxu_row = 1.f, pxv_row = 1.f;
    if (has_px) {
      const int ridx = i - (i0 - 1);
      const int sgm = ridx >> 5;
      const float tu = sgm == 0 ? pxu_tab[0] : (sgm == 1 ? pxu_tab[1] : pxu_tab[2]);
      const float tv = sgm == 0 ? pxv_tab[0] : (sgm == 1 ? pxv_tab[1] : pxv_tab[2]);
      pxu_row = __shfl_sync(FULLMASK, tu, ridx & 31);
      pxv_row = __shfl_sync(FULLMASK, tv, ridx & 31);
    }
    const Row fld_u = nfu, fld_v = nfv;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const float u0 = ua[2].a[k], v0 = va[2].a[k];
      // -divergence(flux)   advection.py:78, finite_differences.py:136-143
      float du = -((f0u.a[k] - f0u_prev.a[k]) * ih0h + (f1u[k + 1] - f1u[k]) * ih1h);
      float dv = -((f0v.a[k] - f0v_prev.a[k]) * ih0h + (f1v[k + 1] - f1v[k]) * ih1h);
      if (c.has_nu) {  // finite_differences.py:127-133, diffusion.py:35-37
        float lu = u0 * c.lap_m2sum;  // (-2 u) * sum_j s_j
        lu += (ua[1].a[k] + ua[3].a[k]) * c.lap_s[0];
        lu += (ue[1 + k] + ue[3 + k]) * c.lap_s[1];
        float lv = v0 * c.lap_m2sum;
        lv += (va[1].a[k] + va[3].a[k]) * c.lap_s[0];
        lv += (ve[1 + k] + ve[3 + k]) * c.lap_s[1];
        du += c.nu * lu;
        dv += c.nu * lv;
      }
      if (PATTERN != 0) {  // forcings.py:125-129 (left-to-right sum), equations.py:108-109
        float fu = 0.f, fv = 0.f;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          constexpr int kSh[3] = {0, 2, 4};
          const int kind = (PATTERN >> kSh[t]) & 3;
          if (kind == 1) {  // (px * py) * scale, absent profile == 1 (exact)
            fu += (pxu_row * pyu[k]) * scale_u;
            fv += (pxv_row * pyv[k]) * scale_v;
          } else if (kind == 2) {
            fu += fld_u.a[k];
            fv += fld_v.a[k];
          } else if (kind == 3) {  // forcings.py:111-113
            fu += c.linear_coef * u0;
            fv += c.linear_coef * v0;
          }
        }
        // forcing / rho (equations.py:109) as a multiplication by 1/rho (exact for rho = 1,
        // <= 1 ulp of the forcing otherwise)
        du = fmaf(fu, c.inv_rho, du);
        dv = fmaf(fv, c.inv_rho, dv);
      }
      us_cur.a[k] = dvdt_mode ? du : u0 + c.dt * du;  // time_stepping.py:101
      vs_cur.a[k] = dvdt_mode ? dv : v0 + c.dt * dv;
    }

    const float vsL = from_left(vs_cur.a[C - 1]);  // v*[i][-1]
    if (i >= i0 && store_ok) {
      const size_t off = boff + (size_t)iw * Ny + jg;
      strow(us + off, us_cur);
      strow(vs + off, vs_cur);
      if (rhs != nullptr) {  // finite_differences.py:136-143 on u*
        Row d;
#pragma unroll
        for (int k = 0; k < C; ++k) {
          const float vleft = (k == 0) ? vsL : vs_cur.a[(k + C - 1) % C];
          d.a[k] = (us_cur.a[k] - us_prev.a[k]) * c.inv_h[0] + (vs_cur.a[k] - vleft) * c.inv_h[1];
        }
        strow(rhs + off, d);
      }
    }

    // ---- slide the window down one row
    us_prev = us_cur;
    f0u_prev = f0u;
    f0v_prev = f0v;
    vL1_cur = vnL1;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      ua[r] = ua[r + 1];
      va[r] = va[r + 1];
    }
    if (LAZY) {
      project_row(nu, nv, qkeep, nq);
      qkeep = nq;
    }
    ua[4] = nu;
    va[4] = nv;
    if (kHasField && i + 1 < iend) {
      if (c.field[0]) nfu = ldrow<C>(c.field[0] + (size_t)(i + 1) * Ny + jg);
      if (c.field[1]) nfv = ldrow<C>(c.field[1] + (size_t)(i + 1) * Ny + jg);
    }
  }
}

}  // namespace

// sq.own == nullptr: (u, v) is a projected state.  sq.own != nullptr: LAZY mode, (u, v) = (u*, v*)
// of the previous step and q its pressure.  Nx = local rows; row0 / nx_global place the slab.
// rows per warp: 64 normally (5 warm-up rows = 8 % redundant work); 16 on small grids, where 64
// would leave most SMs without a warp (2048^2: 576 warps for 148 SMs)
int explicit_2d_tile_rows(int batch, int Nx, int Ny) {
  const int cols0 = 120;
  const long warps64 = (long)((Ny + cols0 - 1) / cols0) * ((Nx + 63) / 64) * batch;
  return warps64 >= 148L * 12 ? 64 : 16;
}

// Rows per warp of a launch that covers the WHOLE grid (the slab-decomposed step launches ranges of
// explicit_2d_tile_rows tiles and keeps those).  All warps of the launch do the same work and 16 of
// them fit an SM (128 registers), so the launch runs in waves of 148 x 16 warps and its last wave
// is only as full as the warp count allows: at 8192^2, 64-row tiles are 8832 warps = 3.73 waves.
// Pick the tile height in [48, 95] (the row profile table holds 96 rows) that fills its last wave
// best, preferring taller tiles (5 warm-up rows each).  CFD_EXPLICIT_TX=n overrides.
static int explicit_2d_tile_rows_full(int batch, int Nx, int Ny, int warp_cols) {
  static const int forced = [] {
    const char* e = getenv("CFD_EXPLICIT_TX");
    return e ? atoi(e) : 0;
  }();
  const int base = explicit_2d_tile_rows(batch, Nx, Ny);
  if (forced >= 2 && forced <= 95) return forced;
  if (forced < 0) return base;
  const long strips = (Ny + warp_cols - 1) / warp_cols;
  if (base != 64) {
    // small grids are latency bound (256^2 with 16-row tiles: 80 warps on 148 SMs, each marching
    // 21 rows): shorter tiles until there are a few warps per SM -- 256^2: 4 rows, stencil 16.4 ->
    // 9.1 us; 2048^2 keeps 16 rows
    int tx = base;
    while (tx > 4 && strips * ((Nx + tx - 1) / tx) * batch < 148L * 4) tx /= 2;
    return tx;
  }
  const long cta_per_tile = (strips + kWarpsPerCta - 1) / kWarpsPerCta;
  const double slots = 148.0 * (16 / kWarpsPerCta);  // CTAs resident at once
  int best = base;
  double best_cost = 1e30;
  for (int tx = 48; tx <= 95; ++tx) {
    const long tiles = (Nx + tx - 1) / tx;
    const double waves = (double)(tiles * cta_per_tile * batch) / slots;
    // time ~ full waves (rounded up) x rows marched per warp, warm-up rows included
    const double cost = ceil(waves - 1e-9) * (tx + 5);
    if (cost < best_cost - 1e-9 || (fabs(cost - best_cost) <= 1e-9 && tx > best)) {
      best_cost = cost;
      best = tx;
    }
  }
  return best;
}

// tile_begin / tile_count: the range of row tiles (explicit_2d_tile_rows rows each) this launch
// covers; tile_count < 0 = all of them.  The slab-decomposed step launches the stencil block by
// block so that the row FFT and the NVLink transfer of a block overlap the stencil of the next.
int launch_explicit_2d_slab(cudaStream_t stream, SlabSrc su, SlabSrc sv, SlabSrc sq, float* us,
                            float* vs, float* rhs, int batch, int Nx, int Ny, int row0,
                            int nx_global, const StepConsts& c, int dvdt_mode, int tile_begin,
                            int tile_count) {
  const float* qprev = sq.own;
  if (Nx < 3) return set_error_msg("a slab needs at least 3 rows");
  static const int forced_cols = [] {  // tuning knob: CFD_EXPLICIT_COLS=2|4 columns per lane
    const char* e = getenv("CFD_EXPLICIT_COLS");
    return (e && (e[0] == '2' || e[0] == '4')) ? e[0] - '0' : 0;
  }();
  // columns per lane: the choice that wastes fewer lanes on this row length (ties -> 4).  When the
  // launch fills the GPU anyway, 4 columns per lane win unless they waste a quarter more lanes (2
  // columns cost 1.5 instead of 1.25 y-face fluxes per cell): 1024 x 256^2, 384 vs 320 lane-columns
  // per row, measured 0.551 ms (4) vs 0.574 ms (2).
  const int work4 = ((Ny + 119) / 120) * 128, work2 = ((Ny + 55) / 56) * 64;
  const bool fills_gpu = (long)((Ny + 119) / 120) * ((Nx + 63) / 64) * batch >= 148L * 16;
  const int cols = forced_cols ? forced_cols
                               : (fills_gpu ? (4 * work2 < 3 * work4 ? 2 : 4) : (work2 < work4 ? 2 : 4));
  const int warp_cols = cols == 4 ? 120 : 56;
  const bool full_launch = tile_count < 0;
  const int TX = full_launch ? explicit_2d_tile_rows_full(batch, Nx, Ny, warp_cols)
                             : explicit_2d_tile_rows(batch, Nx, Ny);
  const int tiles_all = (Nx + TX - 1) / TX;
  if (tile_count < 0) {
    tile_begin = 0;
    tile_count = tiles_all;
  }
  if (tile_begin < 0 || tile_begin + tile_count > tiles_all) return set_error_msg("internal: bad stencil tile range");
  const int strips = (Ny + warp_cols - 1) / warp_cols;
  dim3 grid((strips + kWarpsPerCta - 1) / kWarpsPerCta, tile_count, batch);
  // whole-row warps (WRAP, 8 columns per lane) for GPU-filling batches of 256-column rows
  static const int wrap_ok = [] { const char* e = getenv("CFD_EXPLICIT_WRAP"); return e ? atoi(e) : 1; }();
  const bool wrap8 = wrap_ok && !forced_cols && Ny == 256 && fills_gpu && full_launch;
  if (wrap8) grid = dim3(1, (tile_count + kWarpsPerCta - 1) / kWarpsPerCta, batch);
  int pattern = 0, nt = 0;
  for (int t = 0; t < c.n_terms; ++t) {
    const int kind = c.term_kind[t];
    if (kind == CFD_FORCE_SMAGORINSKY) continue;  // added by smag_add2d_kernel (smagorinsky_2d.cu)
    int code = kind == CFD_FORCE_SEPARABLE ? 1 : kind == CFD_FORCE_FIELD ? 2 : kind == CFD_FORCE_LINEAR ? 3 : -1;
    if (code < 0 || nt >= 3) return set_error_msg("unsupported forcing term for the 2-D kernel");
    for (int q = 0; q < nt; ++q)
      if (((pattern >> (2 * q)) & 3) == code) return set_error_msg("each forcing kind may appear once");
    pattern |= code << (2 * nt++);
  }
#define CFD_EXPL_LAUNCH(P, CC, LZ, WR)                                                      \
  explicit2d_kernel<P, CC, LZ, WR><<<grid, 32 * kWarpsPerCta, 0, stream>>>(                    \
      su, sv, sq, us, vs, rhs, Nx, Ny, row0, nx_global, c, dvdt_mode, TX, tile_begin)
#define CFD_EXPL_CASE(P)                           \
  case P:                                          \
    if (wrap8) {                                   \
      if (qprev) CFD_EXPL_LAUNCH(P, 8, true, true);   \
      else CFD_EXPL_LAUNCH(P, 8, false, true);     \
    } else if (cols == 2) {                        \
      if (qprev) CFD_EXPL_LAUNCH(P, 2, true, false);  \
      else CFD_EXPL_LAUNCH(P, 2, false, false);    \
    } else {                                       \
      if (qprev) CFD_EXPL_LAUNCH(P, 4, true, false);  \
      else CFD_EXPL_LAUNCH(P, 4, false, false);    \
    }                                              \
    break;
  switch (pattern) {
    CFD_EXPL_CASE(0)
    CFD_EXPL_CASE(1) CFD_EXPL_CASE(2) CFD_EXPL_CASE(3)
    CFD_EXPL_CASE(1 | (2 << 2)) CFD_EXPL_CASE(1 | (3 << 2)) CFD_EXPL_CASE(2 | (1 << 2))
    CFD_EXPL_CASE(2 | (3 << 2)) CFD_EXPL_CASE(3 | (1 << 2)) CFD_EXPL_CASE(3 | (2 << 2))
    CFD_EXPL_CASE(1 | (2 << 2) | (3 << 4)) CFD_EXPL_CASE(1 | (3 << 2) | (2 << 4))
    CFD_EXPL_CASE(2 | (1 << 2) | (3 << 4)) CFD_EXPL_CASE(2 | (3 << 2) | (1 << 4))
    CFD_EXPL_CASE(3 | (1 << 2) | (2 << 4)) CFD_EXPL_CASE(3 | (2 << 2) | (1 << 4))
    default:
      return set_error_msg("internal: forcing pattern not instantiated");
  }
#undef CFD_EXPL_CASE
#undef CFD_EXPL_LAUNCH
  count_launch();
  CFD_CUDA_OK(cudaGetLastError());
  return 0;
}


int launch_explicit_2d(cudaStream_t stream, const float* u, const float* v, const float* qprev,
                       float* us, float* vs, float* rhs, int batch, int Nx, int Ny,
                       const StepConsts& c, int dvdt_mode) {
  const SlabSrc su = {u, u, u}, sv = {v, v, v}, sq = {qprev, qprev, qprev};
  return launch_explicit_2d_slab(stream, su, sv, sq, us, vs, rhs, batch, Nx, Ny, 0, Nx, c, dvdt_mode, 0, -1);
}

}  // namespace cfd
