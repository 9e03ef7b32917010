/* C restatement of the 2-D explicit step + projection stencils.  TEST INFRASTRUCTURE ONLY.
 *
 * Used (a) as a second, independent statement of the algorithm checked against
 * oracle/cfd_oracle.py and the golden vectors, and (b) as the multi-threaded CPU baseline that
 * bench.py times beside the GPU numbers (kind "port": JAX is not installed in this image, so the
 * reference's jitted CPU path cannot run; this is the fused, OpenMP-parallel equivalent, with the
 * FFTs done by scipy.fft (pocketfft) on all cores from oracle/cpu_baseline.py).
 * Never linked into or called by the product.
 *
 * Reference lines restated (paths under /root/reference/jax_cfd/base/):
 *   face flux:   interpolation.py:57-62 (linear), 147-151 (upwind), 210-217 (lax_wendroff),
 *                224-231 + 287-297 (van Leer TVD limiter), advection.py:73 (flux = c u)
 *   divergence:  finite_differences.py:95-102, 136-143;  advection.py:78 (negated)
 *   laplacian:   finite_differences.py:127-133;  diffusion.py:35-37
 *   update:      equations.py:105-109, time_stepping.py:101
 *   correction:  pressure.py:194-196
 */
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Thread count of the stencil loops.  Launchers such as torchrun export OMP_NUM_THREADS=1 to
 * their workers; the CPU baseline asks for all cores explicitly (cpu_baseline.CpuStep). */
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static inline float safe_div(float x, float y) { return x / (y != 0.0f ? y : 1.0f); }
static inline float van_leer(float r) { return r > 0.0f ? safe_div(2.0f * r, 1.0f + r) : 0.0f; }

static inline float face_flux(float cL, float c0, float cR, float cRR, float U, float dth) {
  const float C = dth * U;
  const float d = cR - c0;
  float low, high, phi;
  if (U > 0.0f) {
    low = c0;
    high = c0 + 0.5f * (1.0f - C) * d;
    phi = van_leer(safe_div(c0 - cL, d));
  } else {
    low = cR;
    high = cR - 0.5f * (1.0f + C) * d;
    phi = van_leer(safe_div(cRR - cR, d));
  }
  return (low - (low - high) * phi) * U;
}

#define IDX(i, j) ((size_t)(i) * ny + (j))

/* us, vs = v + dt * (conv + nu lap + (fconst + lin * v) / rho);  rhs = div(us, vs) if rhs != NULL.
 * fu / fv: constant forcing fields or NULL. */
void oracle_explicit_2d(const float* u, const float* v, float* us, float* vs, int nx, int ny,
                        float dt, float dthx, float dthy, float hx, float hy, int has_nu,
                        float nu, float rho, const float* fu, const float* fv, int has_lin,
                        float lin) {
  const float sx = (1.0f / hx) * (1.0f / hx), sy = (1.0f / hy) * (1.0f / hy);
  const float ssum = sx + sy;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nx; ++i) {
    const int im1 = (i + nx - 1) % nx, im2 = (i + nx - 2) % nx, ip1 = (i + 1) % nx, ip2 = (i + 2) % nx;
    for (int j = 0; j < ny; ++j) {
      const int jm1 = (j + ny - 1) % ny, jm2 = (j + ny - 2) % ny, jp1 = (j + 1) % ny, jp2 = (j + 2) % ny;
      /* component u (axis a = 0) */
      float Ux1 = 0.5f * u[IDX(i, j)] + 0.5f * u[IDX(ip1, j)];
      float Ux0 = 0.5f * u[IDX(im1, j)] + 0.5f * u[IDX(i, j)];
      float Fx1 = face_flux(u[IDX(im1, j)], u[IDX(i, j)], u[IDX(ip1, j)], u[IDX(ip2, j)], Ux1, dthx);
      float Fx0 = face_flux(u[IDX(im2, j)], u[IDX(im1, j)], u[IDX(i, j)], u[IDX(ip1, j)], Ux0, dthx);
      float Uy1 = 0.5f * v[IDX(i, j)] + 0.5f * v[IDX(ip1, j)];
      float Uy0 = 0.5f * v[IDX(i, jm1)] + 0.5f * v[IDX(ip1, jm1)];
      float Fy1 = face_flux(u[IDX(i, jm1)], u[IDX(i, j)], u[IDX(i, jp1)], u[IDX(i, jp2)], Uy1, dthy);
      float Fy0 = face_flux(u[IDX(i, jm2)], u[IDX(i, jm1)], u[IDX(i, j)], u[IDX(i, jp1)], Uy0, dthy);
      float du = -((Fx1 - Fx0) / hx + (Fy1 - Fy0) / hy);
      /* component v (axis a = 1) */
      float Vx1 = 0.5f * u[IDX(i, j)] + 0.5f * u[IDX(i, jp1)];
      float Vx0 = 0.5f * u[IDX(im1, j)] + 0.5f * u[IDX(im1, jp1)];
      float Gx1 = face_flux(v[IDX(im1, j)], v[IDX(i, j)], v[IDX(ip1, j)], v[IDX(ip2, j)], Vx1, dthx);
      float Gx0 = face_flux(v[IDX(im2, j)], v[IDX(im1, j)], v[IDX(i, j)], v[IDX(ip1, j)], Vx0, dthx);
      float Vy1 = 0.5f * v[IDX(i, j)] + 0.5f * v[IDX(i, jp1)];
      float Vy0 = 0.5f * v[IDX(i, jm1)] + 0.5f * v[IDX(i, j)];
      float Gy1 = face_flux(v[IDX(i, jm1)], v[IDX(i, j)], v[IDX(i, jp1)], v[IDX(i, jp2)], Vy1, dthy);
      float Gy0 = face_flux(v[IDX(i, jm2)], v[IDX(i, jm1)], v[IDX(i, j)], v[IDX(i, jp1)], Vy0, dthy);
      float dv = -((Gx1 - Gx0) / hx + (Gy1 - Gy0) / hy);
      if (has_nu) {
        float lu = -2.0f * u[IDX(i, j)] * ssum;
        lu += (u[IDX(im1, j)] + u[IDX(ip1, j)]) * sx;
        lu += (u[IDX(i, jm1)] + u[IDX(i, jp1)]) * sy;
        float lv = -2.0f * v[IDX(i, j)] * ssum;
        lv += (v[IDX(im1, j)] + v[IDX(ip1, j)]) * sx;
        lv += (v[IDX(i, jm1)] + v[IDX(i, jp1)]) * sy;
        du += nu * lu;
        dv += nu * lv;
      }
      float f0 = 0.0f, f1 = 0.0f;
      int any = 0;
      if (fu) { f0 += fu[IDX(i, j)]; any = 1; }
      if (fv) { f1 += fv[IDX(i, j)]; any = 1; }
      if (has_lin) { f0 += lin * u[IDX(i, j)]; f1 += lin * v[IDX(i, j)]; any = 1; }
      if (any) { du += f0 / rho; dv += f1 / rho; }
      us[IDX(i, j)] = u[IDX(i, j)] + dt * du;
      vs[IDX(i, j)] = v[IDX(i, j)] + dt * dv;
    }
  }
}

void oracle_divergence_2d(const float* u, const float* v, float* rhs, int nx, int ny, float hx,
                          float hy) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nx; ++i) {
    const int im1 = (i + nx - 1) % nx;
    for (int j = 0; j < ny; ++j) {
      const int jm1 = (j + ny - 1) % ny;
      rhs[IDX(i, j)] = (u[IDX(i, j)] - u[IDX(im1, j)]) / hx + (v[IDX(i, j)] - v[IDX(i, jm1)]) / hy;
    }
  }
}

void oracle_correct_2d(const float* us, const float* vs, const float* q, float* uo, float* vo,
                       int nx, int ny, float hx, float hy) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nx; ++i) {
    const int ip1 = (i + 1) % nx;
    for (int j = 0; j < ny; ++j) {
      const int jp1 = (j + 1) % ny;
      uo[IDX(i, j)] = us[IDX(i, j)] - (q[IDX(ip1, j)] - q[IDX(i, j)]) / hx;
      vo[IDX(i, j)] = vs[IDX(i, j)] - (q[IDX(i, jp1)] - q[IDX(i, j)]) / hy;
    }
  }
}
