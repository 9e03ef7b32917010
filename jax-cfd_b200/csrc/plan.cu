// Plan management + the extern "C" ABI declared in include/cfd_b200.h.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "fft_smem.cuh"
#include "fft_rows.cuh"

namespace cfd {

// ---- kernels implemented in the other translation units ------------------------------------
int launch_explicit_2d(cudaStream_t, const float* u, const float* v, const float* qprev, float* us,
                       float* vs, float* rhs, int batch, int Nx, int Ny, const StepConsts& c,
                       int dvdt_mode);
int launch_irfft_rows(cudaStream_t, int lm_row, const float2* T, float* q, int batch, int Nx,
                      const float2* tw, const float2* rtw, int paired);
int launch_correct_2d(cudaStream_t, const float* us, const float* vs, const float* q,
                      const float* qnext, float* uo, float* vo, int batch, int Nx, int Ny,
                      float inv_hx, float inv_hy);
int launch_rfft_rows(cudaStream_t, int lm_row, const float* rhs, float2* T, int batch, int Nx,
                     const float2* tw, const float2* rtw, int paired);
int launch_xlines(cudaStream_t, int lm_x, float2* T, int batch, int My, const float2* tw,
                  const double* lamx, const double* lamy, const float* lamxf, const float* lamyf,
                  int fastd, double cutoff, float norm, float2* scratch, const float2* wbig,
                  const SideStreams* side, int paired, const float* dtab);
int launch_divergence_2d(cudaStream_t, const float* u, const float* v, float* rhs, int batch,
                         int Nx, int Ny, float inv_hx, float inv_hy);
int launch_axpy(cudaStream_t, const float* x, int nterms, const float* const* y, const float* coef,
                float* out, size_t n);
int launch_diag_2d(cudaStream_t, const float* u, const float* v, int batch, int Nx, int Ny,
                   float inv_hx, float inv_hy, double* out4);
int launch_scale(cudaStream_t, const float* x, float numer, float denom, float* out, size_t n);
int launch_downsample_component(cudaStream_t st, const float* in, float* out, int batch, int N0, int N1, int N2,
                                int factor, int direction);
int launch_vorticity_2d(cudaStream_t st, const float* u, const float* v, float* out, int batch, int N0, int N1,
                        float dx, float dy);
int launch_explicit_2d_generic(cudaStream_t, const float* u, const float* v, float* us, float* vs, int batch,
                               int N0, int N1, const StepConsts& c, int dvdt_mode);
void periodic_laplacian_eigenbasis(int N, double h, std::vector<double>* V, std::vector<double>* lam);
int matmul_transform(cudaStream_t st, int ndim, const int64_t* shape, int batch, const float* in, float* out,
                     const double* const* V, const double* const* Vt, const double* diag, double* w1,
                     double* w2);
int launch_smag_nut_2d(cudaStream_t, const float* u, const float* v, float* nut, int batch, int N0,
                       int N1, const StepConsts& c);
int launch_smag_add_2d(cudaStream_t, const float* u, const float* v, const float* nut, float* us,
                       float* vs, int batch, int N0, int N1, const StepConsts& c, int dvdt_mode);

int launch_rfft_rows3(cudaStream_t, int lm, const float* rhs, float2* T, int batch, int NR,
                      const float2* tw, const float2* rtw);
int launch_irfft_rows3(cudaStream_t, int lm, const float2* T, float* q, int batch, int NR,
                       const float2* tw, const float2* rtw);
int launch_lines_scatter(cudaStream_t, int lm, const float2* A, float2* B, int planes, int NL,
                         const float2* tw);
int launch_lines_gather(cudaStream_t, int lm, const float2* B, float2* A, int planes, int NL,
                        const float2* tw);
int launch_xlines3(cudaStream_t, int lm, float2* T, size_t nlines, int N1, int NZP, const float2* tw,
                   const double* const* lam, const float* const* lamf, int fastd, double cutoff,
                   float norm, const float* dtab);
int launch_divergence_3d(cudaStream_t, const float* u, const float* v, const float* w, float* rhs,
                         int batch, int N0, int N1, int N2, float ihx, float ihy, float ihz);
int launch_correct_3d(cudaStream_t, const float* us, const float* vs, const float* ws, const float* q,
                      float* uo, float* vo, float* wo, int batch, int N0, int N1, int N2, float ihx,
                      float ihy, float ihz);
int launch_smag_nut_3d(cudaStream_t, const float* u, const float* v, const float* w, float* nut,
                       float* sfield, int batch, int N0, int N1, int N2, const StepConsts& c);
bool smag_uses_tiles(int N0, int N1, int N2);
int launch_explicit_3d(cudaStream_t, const float* u, const float* v, const float* w, const float* nut,
                       const float* sfield, float* us, float* vs, float* ws, int batch, int N0, int N1,
                       int N2, const StepConsts& c, int dvdt_mode);
bool explicit_3d_uses_march(int N0, int N1, int N2);
int launch_diag_3d(cudaStream_t, const float* u, const float* v, const float* w, int batch, int N0,
                   int N1, int N2, float ihx, float ihy, float ihz, double* out4);

// ---- errors / counters ------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

int set_error(const char* what, cudaError_t e, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  g_err = buf;
  return 1;
}
int set_error_msg(const char* msg) {
  g_err = msg;
  return 1;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static int ilog2(int64_t n) {
  int l = 0;
  while ((int64_t(1) << l) < n) ++l;
  return l;
}
static bool is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

static std::vector<float2> build_twiddles(int lm, int lrmax = 4, bool bal = false) {
  return fft_build_twiddles(lm, lrmax, bal);
}
// ... for the x lines of a 2-D plan: the x-line kernel picks its own schedule per length
static std::vector<float2> build_xline_twiddles(int lm_x) {
  const int lm = lm_x == 15 ? 14 : lm_x;  // 32768-point lines are two 16384-point transforms
  const bool bal = xlines_balanced(lm);
  return build_twiddles(lm, bal ? 5 : 4, bal);
}

}  // namespace cfd

using namespace cfd;

#include "plan_struct.cuh"

namespace {

template <typename T>
int upload(T** dst, const std::vector<T>& src) {
  CFD_CUDA_OK(cudaMalloc((void**)dst, src.size() * sizeof(T)));
  CFD_CUDA_OK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace

namespace cfd {
void prof_mark(cfd_plan* p, cudaStream_t st, const char* name) {
  if (!p->profiling) return;
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  cudaEventRecord(ev, st);
  p->prof_events.push_back(ev);
  p->prof_names.push_back(name);
}

int make_consts(const cfd_plan* p, const cfd_params* prm, StepConsts* c) {
  memset(c, 0, sizeof *c);
  const int d = p->ndim;
  for (int j = 0; j < d; ++j)  // periodic stencils of reach 2 (transform-only plans may be smaller)
    if (p->shape[j] < 2) return set_error_msg("the time step needs at least 2 cells along every axis");
  c->dt = (float)prm->dt;
  float lap_sum = 0.f;
  // the dt of the Lax-Wendroff Courant number (equations.py:127-128 closes `convect` over the
  // builder's dt) may differ from the dt of the time stepper (time_stepping.py:101)
  const double cdt = prm->convect_dt > 0 ? prm->convect_dt : prm->dt;
  for (int j = 0; j < d; ++j) {
    c->dth[j] = (float)(cdt / p->step[j]);
    c->inv_h[j] = (float)(1.0 / p->step[j]);
    const float hf = (float)p->step[j];
    const float inv = 1.0f / hf;
    c->lap_s[j] = inv * inv;
  }
  // np.sum over a float32 array of <= 3 entries: sequential float32 adds
  for (int j = 0; j < d; ++j) lap_sum = (j == 0) ? c->lap_s[0] : lap_sum + c->lap_s[j];
  c->lap_sum = lap_sum;
  c->lap_m2sum = -2.f * lap_sum;
  c->has_nu = prm->has_viscosity ? 1 : 0;
  c->nu = (float)(prm->viscosity / prm->density);
  c->rho = (float)prm->density;
  c->inv_rho = (float)(1.0 / prm->density);
  if (prm->n_terms < 0 || prm->n_terms > CFD_MAX_FORCING_TERMS)
    return set_error_msg("cfd_params.n_terms out of range");
  c->n_terms = prm->n_terms;
  for (int t = 0; t < prm->n_terms; ++t) {
    const int k = prm->term_kind[t];
    if (k < CFD_FORCE_SEPARABLE || k > CFD_FORCE_SMAGORINSKY)
      return set_error_msg("cfd_params.term_kind: unknown forcing kind");
    c->term_kind[t] = k;
  }
  c->linear_coef = (float)prm->linear_coef;
  double prod = 1.0;
  for (int j = 0; j < d; ++j) prod *= p->step[j];
  const double cutoff = pow(prod, 1.0 / d);
  c->smag_coef = (float)((prm->smagorinsky_cs * cutoff) * (prm->smagorinsky_cs * cutoff));
  for (int a = 0; a < CFD_MAX_DIM; ++a) {
    for (int j = 0; j < CFD_MAX_DIM; ++j) c->sep_prof[a][j] = prm->sep_prof[a][j];
    c->sep_scale[a] = prm->sep_scale[a];
    c->has_sep[a] = prm->has_sep[a];
    c->field[a] = prm->field[a];
  }
  return 0;
}

// 32768-point x lines: w_N^m table and the scratch of the split transform (nlines lines per launch)
int big_line_tables(cfd_plan* p, size_t nlines) {
  const int half = 1 << 14;
  std::vector<float2> w(half);
  for (int m = 0; m < half; ++m) {
    const double ang = -2.0 * M_PI * (double)m / (double)(2 * half);
    w[m] = make_float2((float)cos(ang), (float)sin(ang));
  }
  cudaFree(p->wbig);
  cudaFree(p->xscratch);
  p->wbig = nullptr;
  p->xscratch = nullptr;
  int err = upload(&p->wbig, w);
  if (cudaMalloc((void**)&p->xscratch, nlines * (size_t)(2 * half) * sizeof(float2)) != cudaSuccess)
    err |= set_error_msg("scratch allocation for 32768-point lines failed");
  if (p->side.n == 0) {
    for (int i = 0; i < 4; ++i) {
      if (cudaStreamCreateWithFlags(&p->side.s[i], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&p->side.done[i], cudaEventDisableTiming) != cudaSuccess)
        return set_error_msg("side stream creation failed");
    }
    if (cudaEventCreateWithFlags(&p->side.start, cudaEventDisableTiming) != cudaSuccess)
      return set_error_msg("side stream creation failed");
    p->side.n = 4;
  }
  return err;
}

// Which layout the 2-D spectrum uses (poisson_2d.cu): pairs of ky lines interleaved when the row
// kernels hold only one or two rows per CTA (rows of >= 16384 reals: the transposed chunks would be
// 8 / 16 bytes) and the x lines are ones the stride-2 access is built for -- 4096 / 8192 points in
// the kernel itself, 32768 through the pair-wise split / merge.  Measured on one B200:
//   4096 x 16384:  rfft 318 -> 218, x lines 205 -> 226, irfft 293 -> 268 us   (paired wins)
//   8192 x 8192:   rfft 190 -> 183, x lines 223 -> 257, irfft 237 -> 214 us   (a wash: plain)
// Several GPUs: only with the 32768-point split transform, whose peer accesses stay full 16-byte
// pieces; a half-used sector in the x-line kernel would double the NVLink traffic.
// CFD_T_PAIRED=0|1 overrides where the layout is supported.
static int choose_t_paired(int ndim, int lm_row, int lm_x, int world) {
  if (ndim != 2) return 0;
  const bool supported = (lm_x == 12 || lm_x == 13 || lm_x == 15);
  int want = supported && lm_row >= 13 && (world <= 1 || lm_x == 15);
  if (const char* e = getenv("CFD_T_PAIRED")) want = supported && atoi(e) != 0 && (world <= 1 || lm_x == 15);
  return want ? 1 : 0;
}

int plan_tables_create(cfd_plan* p, int ndim, const int64_t* shape, const double* step) {
  const int Nxg = (int)shape[0];
  cudaFree(p->tw_x);
  cudaFree(p->lam[0]);
  cudaFree(p->lamf[0]);
  p->tw_x = nullptr;
  p->lam[0] = nullptr;
  p->lamf[0] = nullptr;
  p->lm_x = ilog2(Nxg);
  p->t_paired = choose_t_paired(ndim, p->lm_row, p->lm_x, p->world);
  int err = upload(&p->tw_x, ndim == 2 ? build_xline_twiddles(p->lm_x) : build_twiddles(p->lm_x));
  if (p->lm_x == 15) err |= big_line_tables(p, (size_t)(shape[1] / 2) / (p->world > 0 ? p->world : 1));
  std::vector<double> lam(Nxg);
  for (int k = 0; k < Nxg; ++k)
    lam[k] = (2.0 * cos(2.0 * M_PI * (double)k / (double)Nxg) - 2.0) / (step[0] * step[0]);
  lam[0] = 0.0;
  err |= upload(&p->lam[0], lam);
  std::vector<float> lamf(lam.begin(), lam.end());
  err |= upload(&p->lamf[0], lamf);
  double minabs = 1e300;
  for (int j = 0; j < ndim; ++j) {
    const double l1 = fabs((2.0 * cos(2.0 * M_PI / (double)shape[j]) - 2.0) / (step[j] * step[j]));
    if (l1 < minabs) minabs = l1;
  }
  p->fastd = (minabs > 4.0 * p->cutoff) ? 1 : 0;
  double cells = 1.0;
  for (int j = 0; j < ndim; ++j) cells *= (double)shape[j];
  p->norm = (float)(1.0 / (2.0 * cells));
  return err;
}
}  // namespace cfd

namespace {

int check_plan(const cfd_plan* p) {
  if (p == nullptr) return set_error_msg("null plan");
  CFD_CUDA_OK(cudaSetDevice(p->device));
  return 0;
}

// Every entry point runs on the plan's device and leaves the caller's current device untouched.
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { cudaGetDevice(&prev); }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// q = pinv(rhs): rfft rows -> x lines (fwd * D * inv) -> irfft rows
int solve_matmul(cfd_plan* p, cudaStream_t st, const float* rhs, float* q, const double* diag) {
  if (int e = matmul_transform(st, p->ndim, p->shape, p->batch, rhs, q, p->mm_V, p->mm_Vt, diag, p->mm_w1,
                               p->mm_w2))
    return e;
  prof_mark(p, st, "matmul_transform");
  return 0;
}

// out = irfftn(D * rfftn(in)) with D the pseudo-inverse of the Laplacian (dtab == nullptr) or a
// caller-supplied real diagonal in line layout (cfd_transform_rfft)
int transform_2d(cfd_plan* p, cudaStream_t st, const float* in, float* q, const float* dtab) {
  const int Nx = (int)p->shape[0], Ny = (int)p->shape[1];
  if (int e = launch_rfft_rows(st, p->lm_row, in, p->T, p->batch, Nx, p->tw_row, p->rtw, p->t_paired))
    return e;
  prof_mark(p, st, "rfft_rows");
  if (int e = launch_xlines(st, p->lm_x, p->T, p->batch, Ny / 2, p->tw_x, p->lam[0], p->lam[1],
                            p->lamf[0], p->lamf[1], p->fastd, p->cutoff, p->norm, p->xscratch, p->wbig,
                            &p->side, p->t_paired, dtab))
    return e;
  prof_mark(p, st, "xlines");
  if (int e = launch_irfft_rows(st, p->lm_row, p->T, q, p->batch, Nx, p->tw_row, p->rtw, p->t_paired))
    return e;
  prof_mark(p, st, "irfft_rows");
  return 0;
}
int solve_2d(cfd_plan* p, cudaStream_t st, float* q) {
  if (p->impl == 1) return solve_matmul(p, st, p->rhs, q, p->mm_diag);
  return transform_2d(p, st, p->rhs, q, nullptr);
}

int correct_2d(cfd_plan* p, cudaStream_t st, const float* us, const float* vs, const float* q,
               float* uo, float* vo) {
  if (int e = launch_correct_2d(st, us, vs, q, nullptr, uo, vo, p->batch, (int)p->shape[0], (int)p->shape[1],
                                (float)(1.0 / p->step[0]), (float)(1.0 / p->step[1])))
    return e;
  prof_mark(p, st, "correct");
  return 0;
}

// 3-D: q = pinv(rhs) in five sweeps (poisson_3d.cu)
int transform_3d(cfd_plan* p, cudaStream_t st, const float* in, float* q, const float* dtab) {
  const int N0 = (int)p->shape[0], N1 = (int)p->shape[1], N2 = (int)p->shape[2];
  const int NZP = N2 / 2 + 1;
  if (int e = launch_rfft_rows3(st, p->lm_row, in, p->T, p->batch, N0 * N1, p->tw_row, p->rtw)) return e;
  prof_mark(p, st, "rfft_z");
  if (int e = launch_lines_scatter(st, p->lm_y, p->T, p->T2, p->batch * NZP, N0, p->tw_y)) return e;
  prof_mark(p, st, "fft_y");
  const float norm = (float)(1.0 / (2.0 * (double)p->cells));
  if (int e = launch_xlines3(st, p->lm_x, p->T2, (size_t)p->batch * NZP * N1, N1, NZP, p->tw_x, p->lam,
                             p->lamf, p->fastd, p->cutoff, norm, dtab))
    return e;
  prof_mark(p, st, "xlines3");
  if (int e = launch_lines_gather(st, p->lm_y, p->T2, p->T, p->batch * NZP, N0, p->tw_y)) return e;
  prof_mark(p, st, "ifft_y");
  if (int e = launch_irfft_rows3(st, p->lm_row, p->T, q, p->batch, N0 * N1, p->tw_row, p->rtw)) return e;
  prof_mark(p, st, "irfft_z");
  return 0;
}
int solve_3d(cfd_plan* p, cudaStream_t st, float* q) {
  if (p->impl == 1) return solve_matmul(p, st, p->rhs, q, p->mm_diag);
  return transform_3d(p, st, p->rhs, q, nullptr);
}

// Smagorinsky workspace: nu_t always; the six strain fields when the marching kernel (and with it
// the strain-field path) is used
int smag_buffers(cfd_plan* p, const StepConsts& c, float** nut, float** sfield) {
  *nut = nullptr;
  *sfield = nullptr;
  bool smag = false;
  for (int t = 0; t < c.n_terms; ++t) smag = smag || c.term_kind[t] == CFD_FORCE_SMAGORINSKY;
  if (!smag) return 0;
  const size_t fbytes = (size_t)p->batch * p->cells * sizeof(float);
  if (!p->nut) CFD_CUDA_OK(cudaMalloc((void**)&p->nut, fbytes));
  *nut = p->nut;
  if (explicit_3d_uses_march((int)p->shape[0], (int)p->shape[1], (int)p->shape[2]) &&
      !smag_uses_tiles((int)p->shape[0], (int)p->shape[1], (int)p->shape[2])) {
    if (!p->sfield) CFD_CUDA_OK(cudaMalloc((void**)&p->sfield, 6 * fbytes));
    *sfield = p->sfield;
  }
  return 0;
}

int step_3d(cfd_plan* p, cudaStream_t st, const float* const* v_in, float* const* v_out, float* q_out,
            const StepConsts& c) {
  const int N0 = (int)p->shape[0], N1 = (int)p->shape[1], N2 = (int)p->shape[2];
  const float ih[3] = {(float)(1.0 / p->step[0]), (float)(1.0 / p->step[1]), (float)(1.0 / p->step[2])};
  float* nut = nullptr;
  float* sfield = nullptr;
  if (int e = smag_buffers(p, c, &nut, &sfield)) return e;
  prof_mark(p, st, "begin");
  if (nut) {
    if (int e = launch_smag_nut_3d(st, v_in[0], v_in[1], v_in[2], nut, sfield, p->batch, N0, N1, N2, c)) return e;
    prof_mark(p, st, "smag_nut");
  }
  if (int e = launch_explicit_3d(st, v_in[0], v_in[1], v_in[2], nut, sfield, p->us[0], p->us[1], p->us[2],
                                 p->batch, N0, N1, N2, c, 0))
    return e;
  prof_mark(p, st, "explicit_3d");
  if (int e = launch_divergence_3d(st, p->us[0], p->us[1], p->us[2], p->rhs, p->batch, N0, N1, N2,
                                   ih[0], ih[1], ih[2]))
    return e;
  prof_mark(p, st, "divergence_3d");
  float* q = q_out ? q_out : p->qbuf;
  if (int e = solve_3d(p, st, q)) return e;
  if (int e = launch_correct_3d(st, p->us[0], p->us[1], p->us[2], q, v_out[0], v_out[1], v_out[2],
                                p->batch, N0, N1, N2, ih[0], ih[1], ih[2]))
    return e;
  prof_mark(p, st, "correct_3d");
  return 0;
}

// the warp-marching stencil kernel (explicit_2d.cu) and the vectorised elementwise kernels take
// rows that are a multiple of 4 columns long and at least 3 rows
bool tuned_2d(const cfd_plan* p) { return p->shape[1] % 4 == 0 && p->shape[0] >= 3; }
// chained steps with lazy projection: the tuned 2-D kernels with the line-FFT solve
bool lazy_capable(const cfd_plan* p) { return p->ndim == 2 && p->impl == 0 && tuned_2d(p); }

bool has_smag(const StepConsts& c) {
  for (int t = 0; t < c.n_terms; ++t)
    if (c.term_kind[t] == CFD_FORCE_SMAGORINSKY) return true;
  return false;
}

// 2-D explicit terms: u* (or dv/dt) and, when `rhs` is given, div(u*).  With the Smagorinsky closure
// (subgrid_models.py:188-213: the last forcing term, evaluated from the projected input) the
// stencil kernel leaves the divergence to a separate sweep after the closure has been added.
int explicit_2d_full(cfd_plan* p, cudaStream_t st, const float* u, const float* v, float* us, float* vs,
                     float* rhs, const StepConsts& c, int dvdt_mode) {
  const int Nx = (int)p->shape[0], Ny = (int)p->shape[1];
  const bool tuned = tuned_2d(p);  // else: one-thread-per-cell kernels (generic.cu)
  if (tuned && !has_smag(c)) {
    if (int e = launch_explicit_2d(st, u, v, nullptr, us, vs, rhs, p->batch, Nx, Ny, c, dvdt_mode)) return e;
    prof_mark(p, st, "explicit_2d");
    return 0;
  }
  if (tuned) {
    if (int e = launch_explicit_2d(st, u, v, nullptr, us, vs, nullptr, p->batch, Nx, Ny, c, dvdt_mode)) return e;
  } else if (int e = launch_explicit_2d_generic(st, u, v, us, vs, p->batch, Nx, Ny, c, dvdt_mode)) {
    return e;
  }
  prof_mark(p, st, "explicit_2d");
  if (has_smag(c)) {
    const size_t fbytes = (size_t)p->batch * p->cells * sizeof(float);
    if (!p->nut) CFD_CUDA_OK(cudaMalloc((void**)&p->nut, fbytes));
    if (int e = launch_smag_nut_2d(st, u, v, p->nut, p->batch, Nx, Ny, c)) return e;
    prof_mark(p, st, "smag_nut_2d");
    if (int e = launch_smag_add_2d(st, u, v, p->nut, us, vs, p->batch, Nx, Ny, c, dvdt_mode)) return e;
    prof_mark(p, st, "smag_add_2d");
  }
  if (rhs) {
    if (int e = launch_divergence_2d(st, us, vs, rhs, p->batch, Nx, Ny, c.inv_h[0], c.inv_h[1])) return e;
    prof_mark(p, st, "divergence_2d");
  }
  return 0;
}

int ensure_lazy_buffers(cfd_plan* p) {
  const size_t fbytes = (size_t)p->batch * p->cells * sizeof(float);
  for (int a = 0; a < p->ndim; ++a)
    if (!p->us2[a]) CFD_CUDA_OK(cudaMalloc((void**)&p->us2[a], fbytes));
  if (!p->qbuf2) CFD_CUDA_OK(cudaMalloc((void**)&p->qbuf2, fbytes));
  return 0;
}

}  // namespace

extern "C" {

const char* cfd_last_error(void) { return g_err.c_str(); }
const char* cfd_version(void) { return "cfd_b200 0.1 (sm_100a)"; }
int cfd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
uint64_t cfd_launch_count(void) { return g_launches.load(); }

// Can the radix-2 line-FFT kernels take this grid?  (every axis a power of two within the lengths
// the kernels are built for)
static const char* fft_shape_problem(int ndim, const int64_t* shape) {
  for (int j = 0; j < ndim; ++j)
    if (!is_pow2(shape[j]) || shape[j] < 16) return "every grid axis must be a power of two >= 16";
  if (shape[ndim - 1] < 32) return "last grid axis must be >= 32";
  if (shape[0] > (ndim == 2 ? (1 << 15) : (1 << 14)))
    return "axis 0 longer than 32768 (2-D) / 16384 (3-D) is not supported";
  if (shape[ndim - 1] > (1 << 15)) return "last axis longer than 32768 is not supported yet";
  for (int j = 1; j + 1 < ndim; ++j)
    if (shape[j] > (1 << 14)) return "middle axis longer than 16384 is not supported yet";
  return nullptr;
}

// tables of the line-FFT implementation
static int create_fft_tables(cfd_plan* p) {
  const int ndim = p->ndim, batch = p->batch;
  const int64_t* shape = p->shape;
  const double* step = p->step;
  const int Nx = (int)shape[0], Ny = (int)shape[ndim - 1];
  p->lm_row = ilog2(Ny / 2);
  p->lm_x = ilog2(Nx);
  p->t_paired = choose_t_paired(ndim, p->lm_row, p->lm_x, p->world);
  int err = 0;
  err |= upload(&p->tw_row, build_twiddles(p->lm_row));
  err |= upload(&p->tw_x, ndim == 2 ? build_xline_twiddles(p->lm_x) : build_twiddles(p->lm_x));
  if (p->lm_x == 15 && ndim == 2) err |= big_line_tables(p, (size_t)batch * (Ny / 2));
  if (ndim == 3) {
    p->lm_y = ilog2(shape[1]);
    err |= upload(&p->tw_y, build_twiddles(p->lm_y));
  }
  {
    const int M = Ny / 2;
    std::vector<float2> rtw(M / 2 + 1);
    for (int k = 0; k <= M / 2; ++k) {  // -i * exp(-2 pi i k / N)
      const double ang = -2.0 * M_PI * (double)k / (double)Ny - 0.5 * M_PI;
      rtw[k] = make_float2((float)cos(ang), (float)sin(ang));
    }
    err |= upload(&p->rtw, rtw);
  }
  for (int j = 0; j < ndim; ++j) {
    // eigenvalues of the periodic second-difference operator (array_utils.py:168-173 column
    // [-2, 1, 0, ..., 0, 1] / h^2 transformed by np.fft.fft, fast_diagonalization.py:211-212)
    const int n = (int)shape[j];
    const int len = (j == ndim - 1) ? n / 2 + 1 : n;
    std::vector<double> lam(len);
    for (int k = 0; k < len; ++k)
      lam[k] = (2.0 * cos(2.0 * M_PI * (double)k / (double)n) - 2.0) / (step[j] * step[j]);
    lam[0] = 0.0;
    err |= upload(&p->lam[j], lam);
    std::vector<float> lamf(lam.begin(), lam.end());
    err |= upload(&p->lamf[j], lamf);
  }
  {
    // Every eigenvalue sum except the mean mode has |L| >= min_j |lam_j[1]| (all lam <= 0).  If
    // that exceeds the cutoff with a safety margin, only L(0,...,0) = 0 is discarded and the
    // kernels may form the sum in float32 without testing it.
    double minabs = 1e300;
    for (int j = 0; j < ndim; ++j) {
      const double l1 = fabs((2.0 * cos(2.0 * M_PI / (double)shape[j]) - 2.0) / (step[j] * step[j]));
      if (l1 < minabs) minabs = l1;
    }
    p->fastd = (minabs > 4.0 * p->cutoff) ? 1 : 0;
  }
  p->norm = (float)(1.0 / (2.0 * (double)p->cells));
  const size_t fbytes = (size_t)batch * p->cells * sizeof(float);
  // spectrum: packed Ny/2 columns in 2-D, unpacked N2/2 + 1 planes (two buffers) in 3-D
  const size_t tbytes = ndim == 2 ? fbytes
                                  : (size_t)batch * (shape[2] / 2 + 1) * shape[0] * shape[1] * sizeof(float2);
  if (!err && cudaMalloc((void**)&p->T, tbytes) != cudaSuccess) err = set_error_msg("workspace allocation failed");
  if (!err && ndim == 3 && cudaMalloc((void**)&p->T2, tbytes) != cudaSuccess) err = set_error_msg("workspace allocation failed");
  p->workspace_bytes += tbytes * (ndim == 3 ? 2 : 1);
  return err;
}

// tables of the matmul implementation: analytic real eigenbasis of the periodic Laplacian along
// each axis (what np.linalg.eigh returns up to the choice of basis inside the two-dimensional
// eigenspaces, which the product X f(L) X^T does not depend on) and the pseudo-inverse diagonal,
// formed in float64 and narrowed to float32 like the reference's (fast_diagonalization.py:143-144)
static int create_matmul_tables(cfd_plan* p) {
  const int ndim = p->ndim;
  std::vector<double> lam[CFD_MAX_DIM];
  int err = 0;
  for (int j = 0; j < ndim; ++j) {
    const int n = (int)p->shape[j];
    std::vector<double> V, Vt((size_t)n * n);
    periodic_laplacian_eigenbasis(n, p->step[j], &V, &lam[j]);
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < n; ++c) Vt[(size_t)c * n + i] = V[(size_t)i * n + c];
    err |= upload(&p->mm_V[j], V);
    err |= upload(&p->mm_Vt[j], Vt);
    p->workspace_bytes += 2 * V.size() * sizeof(double);
  }
  std::vector<double> diag(p->cells);
  const int N1 = (int)p->shape[1], N2 = ndim == 3 ? (int)p->shape[2] : 1;
  for (size_t idx = 0; idx < p->cells; ++idx) {
    const int k2 = (int)(idx % N2), k1 = (int)((idx / N2) % N1), k0 = (int)(idx / ((size_t)N1 * N2));
    const double L = lam[0][k0] + lam[1][k1] + (ndim == 3 ? lam[2][k2] : 0.0);
    diag[idx] = fabs(L) > p->cutoff ? (double)(float)(1.0 / L) : 0.0;  // fast_diagonalization.py:259-262
  }
  err |= upload(&p->mm_diag, diag);
  const size_t wbytes = (size_t)p->batch * p->cells * sizeof(double);
  if (!err && cudaMalloc((void**)&p->mm_w1, wbytes) != cudaSuccess) err = set_error_msg("workspace allocation failed");
  if (!err && cudaMalloc((void**)&p->mm_w2, wbytes) != cudaSuccess) err = set_error_msg("workspace allocation failed");
  p->workspace_bytes += 2 * wbytes + p->cells * sizeof(double);
  return err;
}

int cfd_plan_create_impl(cfd_plan** out, int ndim, const int64_t* shape, const double* step, int batch,
                         int device, int implementation) {
  if (out == nullptr || shape == nullptr || step == nullptr) return set_error_msg("null argument");
  *out = nullptr;
  if (ndim != 2 && ndim != 3) return set_error_msg("cfd_plan_create: ndim must be 2 or 3");
  if (batch < 1) return set_error_msg("batch must be >= 1");
  for (int j = 0; j < ndim; ++j) {
    if (shape[j] < 1) return set_error_msg("every grid axis needs at least one cell");
    if (!(step[j] > 0)) return set_error_msg("grid step must be positive");
  }
  if (implementation < CFD_IMPL_AUTO || implementation > CFD_IMPL_MATMUL)
    return set_error_msg("unknown fast-diagonalisation implementation");
  const char* fft_problem = fft_shape_problem(ndim, shape);
  if (implementation == CFD_IMPL_RFFT && fft_problem) return set_error_msg(fft_problem);
  const int impl = (implementation == CFD_IMPL_MATMUL || (implementation == CFD_IMPL_AUTO && fft_problem)) ? 1 : 0;
  if (impl == 1) {
    for (int j = 0; j < ndim; ++j)
      if (shape[j] > 4096)
        return set_error_msg("the matmul implementation is meant for small grids: axes up to 4096 cells");
  }
  if (device < 0 || cfd_device_count() <= device) return set_error_msg("no such CUDA device (no CPU fallback)");
  DeviceGuard guard_;
  CFD_CUDA_OK(cudaSetDevice(device));
  cfd_plan* p = new cfd_plan();
  p->ndim = ndim;
  p->batch = batch;
  p->device = device;
  p->impl = impl;
  p->cells = 1;
  for (int j = 0; j < ndim; ++j) {
    p->shape[j] = shape[j];
    p->step[j] = step[j];
    p->cells *= (size_t)shape[j];
  }
  p->cutoff = 10.0 * 1.1920928955078125e-07;  // 10 * finfo(float32).eps, fast_diagonalization.py:257-258
  const size_t fbytes = (size_t)batch * p->cells * sizeof(float);
  p->workspace_bytes = (size_t)(ndim + 2) * fbytes;
  int err = impl == 1 ? create_matmul_tables(p) : create_fft_tables(p);
  for (int a = 0; a < ndim && !err; ++a) {
    if (cudaMalloc((void**)&p->us[a], fbytes) != cudaSuccess) err = set_error_msg("workspace allocation failed");
  }
  if (!err && cudaMalloc((void**)&p->rhs, fbytes) != cudaSuccess) err = set_error_msg("workspace allocation failed");
  if (!err && cudaMalloc((void**)&p->qbuf, fbytes) != cudaSuccess) err = set_error_msg("workspace allocation failed");
  if (!err && cudaMalloc((void**)&p->diag_dev, 8 * sizeof(double)) != cudaSuccess) err = set_error_msg("workspace allocation failed");
  if (err) {
    cudaGetLastError();
    cfd_plan_destroy(p);
    return 1;
  }
  *out = p;
  return 0;
}

int cfd_plan_create(cfd_plan** out, int ndim, const int64_t* shape, const double* step, int batch,
                    int device) {
  return cfd_plan_create_impl(out, ndim, shape, step, batch, device, CFD_IMPL_AUTO);
}

int cfd_plan_implementation(const cfd_plan* p) { return p ? (p->impl == 1 ? CFD_IMPL_MATMUL : CFD_IMPL_RFFT) : 0; }

void cfd_plan_destroy(cfd_plan* p) {
  if (p == nullptr) return;
  DeviceGuard guard_;
  cudaSetDevice(p->device);
  cudaFree(p->tw_row);
  if (p->shared) {
    for (int r = 0; r < p->world; ++r)
      if (r != p->rank && p->peer_shared[r] && p->peer_shared[r] != p->shared)
        cudaIpcCloseMemHandle(p->peer_shared[r]);
    cudaFree(p->shared);
  }
  cudaFree(p->tw_x);
  cudaFree(p->wbig);
  cudaFree(p->xscratch);
  for (int i = 0; i < p->side.n; ++i) {
    cudaStreamDestroy(p->side.s[i]);
    cudaEventDestroy(p->side.done[i]);
  }
  if (p->side.start) cudaEventDestroy(p->side.start);
  if (p->st_comm) cudaStreamDestroy(p->st_comm);
  for (int i = 0; i < 8; ++i) {
    if (p->ev_blk[i]) cudaEventDestroy(p->ev_blk[i]);
    if (p->ev_chk[i]) cudaEventDestroy(p->ev_chk[i]);
  }
  if (p->ev_comm_done) cudaEventDestroy(p->ev_comm_done);
  if (p->pair_graph) cudaGraphExecDestroy(p->pair_graph);
  cudaFree(p->xstage);
  if (p->ev_ready) cudaEventDestroy(p->ev_ready);
  for (int r = 0; r < CFD_MAX_PEERS; ++r) {
    if (p->st_in[r]) cudaStreamDestroy(p->st_in[r]);
    if (p->st_out[r]) cudaStreamDestroy(p->st_out[r]);
    if (p->ev_out[r]) cudaEventDestroy(p->ev_out[r]);
    for (int c = 0; c < 8; ++c)
      if (p->ev_in[c][r]) cudaEventDestroy(p->ev_in[c][r]);
  }
  for (int c = 0; c < 8; ++c)
    if (p->ev_comp[c]) cudaEventDestroy(p->ev_comp[c]);
  for (int j = 0; j < CFD_MAX_DIM; ++j) {
    cudaFree(p->mm_V[j]);
    cudaFree(p->mm_Vt[j]);
  }
  cudaFree(p->mm_diag);
  cudaFree(p->mm_w1);
  cudaFree(p->mm_w2);
  cudaFree(p->tw_y);
  cudaFree(p->T2);
  cudaFree(p->nut);
  cudaFree(p->sfield);
  cudaFree(p->rtw);
  for (int j = 0; j < CFD_MAX_DIM; ++j) {
    cudaFree(p->lam[j]);
    cudaFree(p->lamf[j]);
    cudaFree(p->us[j]);
    cudaFree(p->us2[j]);
    cudaFree(p->dev_a[j]);
    cudaFree(p->dev_b[j]);
  }
  cudaFree(p->rhs);
  cudaFree(p->qbuf);
  cudaFree(p->qbuf2);
  cudaFree(p->T);
  cudaFree(p->dev_q);
  cudaFree(p->diag_dev);
  if (p->host_stream) cudaStreamDestroy(p->host_stream);
  for (auto ev : p->prof_events) cudaEventDestroy(ev);
  delete p;
}

size_t cfd_plan_workspace_bytes(const cfd_plan* p) { return p ? p->workspace_bytes : 0; }

int cfd_step(cfd_plan* p, cfd_stream stream, const float* const* v_in, float* const* v_out,
             float* q_out, const cfd_params* params) {
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (!v_in || !v_out || !params) return set_error_msg("null argument");
  for (int a = 0; a < p->ndim; ++a) {
    if (!v_in[a] || !v_out[a]) return set_error_msg("null velocity component pointer");
    if (v_in[a] == v_out[a]) return set_error_msg("cfd_step: v_out must not alias v_in");
  }
  cudaStream_t st = (cudaStream_t)stream;
  StepConsts c;
  if (int e = make_consts(p, params, &c)) return e;
  if (p->ndim == 3) return step_3d(p, st, v_in, v_out, q_out, c);
  float* q = q_out ? q_out : p->qbuf;
  prof_mark(p, st, "begin");
  if (int e = explicit_2d_full(p, st, v_in[0], v_in[1], p->us[0], p->us[1], p->rhs, c, 0)) return e;
  if (int e = solve_2d(p, st, q)) return e;
  return correct_2d(p, st, p->us[0], p->us[1], q, v_out[0], v_out[1]);
}

// nsteps chained steps: the projected state is materialised only at the end; in between, step
// n+1 reads (u*, v*, q) of step n and projects while loading (LAZY explicit kernel).
// Small grids are launch-bound (256^2: four kernels of 4-8 us each per step).  Two chained lazy
// steps bring the ping-pong buffers back to where they started, so that pair is captured once per
// plan (and set of constants) as a CUDA graph and replayed.  CFD_GRAPH=0 disables it.
static bool use_pair_graph(const cfd_plan* p) {
  static const int enabled = [] {
    const char* e = getenv("CFD_GRAPH");
    return e ? atoi(e) : 1;
  }();
  static const size_t max_cells = [] {
    const char* e = getenv("CFD_GRAPH_CELLS");
    return e ? (size_t)atoll(e) : ((size_t)1 << 22);  // up to 2048^2 (measured there: 89.6 -> 82.0 us per step)
  }();
  return enabled && !p->profiling && p->ndim == 2 && p->lm_x != 15 && (size_t)p->batch * p->cells <= max_cells;
}

static int lazy_step(cfd_plan* p, cudaStream_t st, const StepConsts& c, float* const* us_cur,
                     const float* q_cur, float* const* us_nxt, float* q_nxt) {
  const int Nx = (int)p->shape[0], Ny = (int)p->shape[1];
  if (int e = launch_explicit_2d(st, us_cur[0], us_cur[1], q_cur, us_nxt[0], us_nxt[1], p->rhs, p->batch,
                                 Nx, Ny, c, 0))
    return e;
  prof_mark(p, st, "explicit_2d_lazy");
  return solve_2d(p, st, q_nxt);
}

// Captures (us, qbuf) -> (us2, qbuf2) -> (us, qbuf) on `st`; on any failure the plan simply keeps
// launching eagerly.
static bool capture_pair_graph(cfd_plan* p, cudaStream_t st, const StepConsts& c) {
  if (p->pair_graph) {
    cudaGraphExecDestroy(p->pair_graph);
    p->pair_graph = nullptr;
  }
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  const uint64_t before = cfd_launch_count();
  int e = lazy_step(p, st, c, p->us, p->qbuf, p->us2, p->qbuf2);
  if (!e) e = lazy_step(p, st, c, p->us2, p->qbuf2, p->us, p->qbuf);
  count_launch(-(int)(cfd_launch_count() - before));  // captured, not launched
  cudaGraph_t g = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &g);
  if (e || ce != cudaSuccess || !g) {
    cudaGetLastError();
    if (g) cudaGraphDestroy(g);
    return false;
  }
  const cudaError_t ie = cudaGraphInstantiate(&p->pair_graph, g, 0);
  cudaGraphDestroy(g);
  if (ie != cudaSuccess) {
    cudaGetLastError();
    p->pair_graph = nullptr;
    return false;
  }
  p->pair_consts = c;
  return true;
}

static int repeated_lazy(cfd_plan* p, cudaStream_t st, const float* const* v_in, float* const* v_out,
                         int nsteps, const cfd_params* params) {
  StepConsts c;
  if (int e = make_consts(p, params, &c)) return e;
  if (int e = ensure_lazy_buffers(p)) return e;
  const int Nx = (int)p->shape[0], Ny = (int)p->shape[1];
  float* us_cur[2] = {p->us[0], p->us[1]};
  float* us_nxt[2] = {p->us2[0], p->us2[1]};
  float* q_cur = p->qbuf;
  float* q_nxt = p->qbuf2;
  if (use_pair_graph(p) && nsteps >= 6) {
    // step 0 and one eager lazy pair first (every kernel variant has then run once outside a
    // capture), then pairs by graph, then at most one eager step
    if (int e = launch_explicit_2d(st, v_in[0], v_in[1], nullptr, p->us[0], p->us[1], p->rhs, p->batch, Nx,
                                   Ny, c, 0))
      return e;
    if (int e = solve_2d(p, st, p->qbuf)) return e;
    int left = nsteps - 1;
    const bool have = p->pair_graph && memcmp(&p->pair_consts, &c, sizeof c) == 0;
    if (!have) {
      if (int e = lazy_step(p, st, c, p->us, p->qbuf, p->us2, p->qbuf2)) return e;
      if (int e = lazy_step(p, st, c, p->us2, p->qbuf2, p->us, p->qbuf)) return e;
      left -= 2;
      capture_pair_graph(p, st, c);
    }
    if (p->pair_graph) {
      for (; left >= 2; left -= 2) {
        CFD_CUDA_OK(cudaGraphLaunch(p->pair_graph, st));
        count_launch(8);
      }
    }
    bool in_pong = false;
    for (; left > 0; --left) {
      if (int e = in_pong ? lazy_step(p, st, c, p->us2, p->qbuf2, p->us, p->qbuf)
                          : lazy_step(p, st, c, p->us, p->qbuf, p->us2, p->qbuf2))
        return e;
      in_pong = !in_pong;
    }
    return in_pong ? correct_2d(p, st, p->us2[0], p->us2[1], p->qbuf2, v_out[0], v_out[1])
                   : correct_2d(p, st, p->us[0], p->us[1], p->qbuf, v_out[0], v_out[1]);
  }
  for (int n = 0; n < nsteps; ++n) {
    prof_mark(p, st, "begin");
    if (n == 0) {
      if (int e = launch_explicit_2d(st, v_in[0], v_in[1], nullptr, us_cur[0], us_cur[1], p->rhs,
                                     p->batch, Nx, Ny, c, 0))
        return e;
      prof_mark(p, st, "explicit_2d");
      if (int e = solve_2d(p, st, q_cur)) return e;
    } else {
      if (int e = launch_explicit_2d(st, us_cur[0], us_cur[1], q_cur, us_nxt[0], us_nxt[1], p->rhs,
                                     p->batch, Nx, Ny, c, 0))
        return e;
      prof_mark(p, st, "explicit_2d_lazy");
      if (int e = solve_2d(p, st, q_nxt)) return e;
      for (int a = 0; a < 2; ++a) {
        float* t = us_cur[a];
        us_cur[a] = us_nxt[a];
        us_nxt[a] = t;
      }
      float* t = q_cur;
      q_cur = q_nxt;
      q_nxt = t;
    }
  }
  return correct_2d(p, st, us_cur[0], us_cur[1], q_cur, v_out[0], v_out[1]);
}

int cfd_repeated(cfd_plan* p, cfd_stream stream, float* const* v_a, float* const* v_b, int nsteps,
                 const cfd_params* params, int* result_in_b) {
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (!v_a || !v_b || !params) return set_error_msg("null argument");
  if (nsteps < 0) return set_error_msg("nsteps must be >= 0");
  for (int a = 0; a < p->ndim; ++a)
    if (!v_a[a] || !v_b[a] || v_a[a] == v_b[a]) return set_error_msg("cfd_repeated: bad buffers");
  // same convention as stepping one by one: the result lands in v_b iff nsteps is odd
  if (result_in_b) *result_in_b = nsteps & 1;
  if (nsteps == 0) return 0;
  float* const* dst = (nsteps & 1) ? v_b : v_a;
  if (nsteps == 1) return cfd_step(p, stream, v_a, dst, nullptr, params);
  bool plain = !lazy_capable(p);  // 3-D, matmul plans, odd shapes: plain ping-pong
  if (!plain) {
    // the Smagorinsky closure is evaluated from the projected state: no lazy chain either
    StepConsts c;
    if (int e = make_consts(p, params, &c)) return e;
    plain = has_smag(c);
  }
  if (plain) {
    for (int n = 0; n < nsteps; ++n) {
      float* const* src = (n & 1) ? v_b : v_a;
      float* const* out = (n & 1) ? v_a : v_b;
      if (int e = cfd_step(p, stream, src, out, nullptr, params)) return e;
    }
    return 0;
  }
  // the chain reads v_a only in its first kernel, so writing the result back into v_a is safe
  return repeated_lazy(p, (cudaStream_t)stream, v_a, dst, nsteps, params);
}

// nsteps steps from v_in into v_out; v_in is never written (what an XLA FFI handler needs: operands
// are immutable).  Lazy-capable plans chain inside the workspace; the others ping-pong between
// v_out and one plan-owned scratch state, arranged so that the last step lands in v_out.
int cfd_advance(cfd_plan* p, cfd_stream stream, const float* const* v_in, float* const* v_out, int nsteps,
                const cfd_params* params) {
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (!v_in || !v_out || !params) return set_error_msg("null argument");
  if (nsteps < 1) return set_error_msg("nsteps must be >= 1");
  for (int a = 0; a < p->ndim; ++a)
    if (!v_in[a] || !v_out[a] || v_in[a] == v_out[a]) return set_error_msg("cfd_advance: bad buffers");
  if (nsteps == 1) return cfd_step(p, stream, v_in, v_out, nullptr, params);
  bool plain = !lazy_capable(p);
  if (!plain) {
    StepConsts c;
    if (int e = make_consts(p, params, &c)) return e;
    plain = has_smag(c);
  }
  if (!plain) return repeated_lazy(p, (cudaStream_t)stream, v_in, v_out, nsteps, params);
  const size_t bytes = (size_t)p->batch * p->cells * sizeof(float);
  for (int a = 0; a < p->ndim; ++a)
    if (!p->dev_b[a]) CFD_CUDA_OK(cudaMalloc((void**)&p->dev_b[a], bytes));
  const float* src[CFD_MAX_DIM] = {v_in[0], v_in[1], p->ndim == 3 ? v_in[2] : nullptr};
  for (int k = 1; k <= nsteps; ++k) {
    float* const* dst = ((nsteps - k) % 2 == 0) ? v_out : p->dev_b;
    if (int e = cfd_step(p, stream, src, dst, nullptr, params)) return e;
    for (int a = 0; a < p->ndim; ++a) src[a] = dst[a];
  }
  return 0;
}

int cfd_explicit_terms(cfd_plan* p, cfd_stream stream, const float* const* v_in,
                       float* const* dvdt_out, const cfd_params* params) {
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (!v_in || !dvdt_out || !params) return set_error_msg("null argument");
  StepConsts c;
  if (int e = make_consts(p, params, &c)) return e;
  for (int a = 0; a < p->ndim; ++a)
    if (v_in[a] == dvdt_out[a]) return set_error_msg("cfd_explicit_terms: output must not alias input");
  if (p->ndim == 3) {
    const int N0 = (int)p->shape[0], N1 = (int)p->shape[1], N2 = (int)p->shape[2];
    float* nut = nullptr;
    float* sfield = nullptr;
    if (int e = smag_buffers(p, c, &nut, &sfield)) return e;
    if (nut)
      if (int e = launch_smag_nut_3d((cudaStream_t)stream, v_in[0], v_in[1], v_in[2], nut, sfield, p->batch, N0, N1, N2, c))
        return e;
    return launch_explicit_3d((cudaStream_t)stream, v_in[0], v_in[1], v_in[2], nut, sfield, dvdt_out[0],
                              dvdt_out[1], dvdt_out[2], p->batch, N0, N1, N2, c, 1);
  }
  return explicit_2d_full(p, (cudaStream_t)stream, v_in[0], v_in[1], dvdt_out[0], dvdt_out[1], nullptr, c, 1);
}

int cfd_project(cfd_plan* p, cfd_stream stream, const float* const* v_in, float* const* v_out,
                float* q_out) {
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (!v_in || !v_out) return set_error_msg("null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->ndim == 3) {
    const int N0 = (int)p->shape[0], N1 = (int)p->shape[1], N2 = (int)p->shape[2];
    const float ih[3] = {(float)(1.0 / p->step[0]), (float)(1.0 / p->step[1]), (float)(1.0 / p->step[2])};
    if (int e = launch_divergence_3d(st, v_in[0], v_in[1], v_in[2], p->rhs, p->batch, N0, N1, N2, ih[0], ih[1], ih[2]))
      return e;
    float* q3 = q_out ? q_out : p->qbuf;
    if (int e = solve_3d(p, st, q3)) return e;
    return launch_correct_3d(st, v_in[0], v_in[1], v_in[2], q3, v_out[0], v_out[1], v_out[2], p->batch,
                             N0, N1, N2, ih[0], ih[1], ih[2]);
  }
  const int Nx = (int)p->shape[0], Ny = (int)p->shape[1];
  if (int e = launch_divergence_2d(st, v_in[0], v_in[1], p->rhs, p->batch, Nx, Ny,
                                   (float)(1.0 / p->step[0]), (float)(1.0 / p->step[1])))
    return e;
  float* q = q_out ? q_out : p->qbuf;
  if (int e = solve_2d(p, st, q)) return e;
  return correct_2d(p, st, v_in[0], v_in[1], q, v_out[0], v_out[1]);
}

int cfd_transform_rfft(cfd_plan* p, cfd_stream stream, const float* in, float* out, const float* diag_lines) {
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (!in || !out || !diag_lines) return set_error_msg("null argument");
  if (p->impl != 0) return set_error_msg("cfd_transform_rfft needs a plan of the rfft implementation");
  // the row kernels read `in` completely before `out` is written: aliasing is fine
  return p->ndim == 3 ? transform_3d(p, (cudaStream_t)stream, in, out, diag_lines)
                      : transform_2d(p, (cudaStream_t)stream, in, out, diag_lines);
}

int cfd_transform_matmul(cfd_plan* p, cfd_stream stream, const float* in, float* out,
                         const double* const* eigvecs, const double* const* eigvecs_t, const double* diag) {
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (!in || !out || !eigvecs || !eigvecs_t || !diag) return set_error_msg("null argument");
  if (p->impl != 1) return set_error_msg("cfd_transform_matmul needs a plan of the matmul implementation");
  return matmul_transform((cudaStream_t)stream, p->ndim, p->shape, p->batch, in, out, eigvecs, eigvecs_t, diag,
                          p->mm_w1, p->mm_w2);
}

int cfd_downsample_component(cfd_stream stream, const float* in, float* out, int ndim, const int64_t* shape,
                             int batch, int direction, int factor) {
  if (!in || !out || !shape) return set_error_msg("null argument");
  if (ndim != 2 && ndim != 3) return set_error_msg("cfd_downsample_component: ndim must be 2 or 3");
  if (direction < 0 || direction >= ndim) return set_error_msg("cfd_downsample_component: bad direction");
  if (factor < 1 || batch < 1) return set_error_msg("cfd_downsample_component: factor and batch must be >= 1");
  for (int j = 0; j < ndim; ++j)  // along `direction` the strided slice drops a remainder (resize.py:72)
    if (shape[j] < 1 || (j != direction && shape[j] % factor))
      return set_error_msg("`block_size` must divide `array.shape`");  // array_utils.py:159-160
  DeviceGuard guard_;
  int dev = 0;
  if (int e = cfd_pointer_device(in, &dev)) return e;  // no CUDA device / host pointer: no CPU fallback
  CFD_CUDA_OK(cudaSetDevice(dev));
  return launch_downsample_component((cudaStream_t)stream, in, out, batch, (int)shape[0], (int)shape[1],
                                     ndim == 3 ? (int)shape[2] : 1, factor, direction);
}

int cfd_vorticity_2d(cfd_stream stream, const float* u, const float* v, float* out, const int64_t* shape,
                     int batch, double dx, double dy) {
  if (!u || !v || !out || !shape) return set_error_msg("null argument");
  if (shape[0] < 1 || shape[1] < 1 || batch < 1) return set_error_msg("cfd_vorticity_2d: bad shape");
  DeviceGuard guard_;
  int dev = 0;
  if (int e = cfd_pointer_device(u, &dev)) return e;
  CFD_CUDA_OK(cudaSetDevice(dev));
  return launch_vorticity_2d((cudaStream_t)stream, u, v, out, batch, (int)shape[0], (int)shape[1], (float)dx,
                             (float)dy);
}

int cfd_scale(cfd_plan* p, cfd_stream stream, const float* const* x, double numer, double denom,
              float* const* out) {
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (!x || !out) return set_error_msg("null argument");
  const size_t n = (size_t)p->batch * p->cells;
  for (int a = 0; a < p->ndim; ++a)
    if (int e = launch_scale((cudaStream_t)stream, x[a], (float)numer, (float)denom, out[a], n)) return e;
  return 0;
}

int cfd_axpy(cfd_plan* p, cfd_stream stream, const float* const* x, int nterms,
             const float* const* const* y, const double* coef, float* const* out) {
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (nterms < 0 || nterms > 4) return set_error_msg("cfd_axpy: 0 <= nterms <= 4");
  const size_t n = (size_t)p->batch * p->cells;
  for (int a = 0; a < p->ndim; ++a) {
    const float* ys[4] = {nullptr, nullptr, nullptr, nullptr};
    float cf[4] = {0, 0, 0, 0};
    for (int k = 0; k < nterms; ++k) {
      ys[k] = y[k][a];
      cf[k] = (float)coef[k];
    }
    if (int e = launch_axpy((cudaStream_t)stream, x[a], nterms, ys, cf, out[a], n)) return e;
  }
  return 0;
}

int cfd_diagnostics(cfd_plan* p, cfd_stream stream, const float* const* v, cfd_diag* out) {
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (!v || !out) return set_error_msg("null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->ndim == 3) {
    if (int e = launch_diag_3d(st, v[0], v[1], v[2], p->batch, (int)p->shape[0], (int)p->shape[1],
                               (int)p->shape[2], (float)(1.0 / p->step[0]), (float)(1.0 / p->step[1]),
                               (float)(1.0 / p->step[2]), p->diag_dev))
      return e;
  } else if (int e = launch_diag_2d(st, v[0], v[1], p->batch, (int)p->shape[0], (int)p->shape[1],
                                    (float)(1.0 / p->step[0]), (float)(1.0 / p->step[1]), p->diag_dev))
    return e;
  double h[4];
  CFD_CUDA_OK(cudaMemcpyAsync(h, p->diag_dev, sizeof h, cudaMemcpyDeviceToHost, st));
  CFD_CUDA_OK(cudaStreamSynchronize(st));
  const double n = (double)p->batch * (double)p->cells;
  out->kinetic_energy = h[0] / n;
  out->enstrophy = h[1] / n;
  out->max_abs_div = h[2];
  out->max_speed_sq = h[3];
  return 0;
}

int cfd_step_host(cfd_plan* p, const float* const* v_in_host, float* const* v_out_host,
                  float* q_out_host, int nsteps, const cfd_params* params) {
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (!v_in_host || !v_out_host || !params) return set_error_msg("null argument");
  if (nsteps < 1) return set_error_msg("nsteps must be >= 1");
  const size_t bytes = (size_t)p->batch * p->cells * sizeof(float);
  if (!p->host_stream) CFD_CUDA_OK(cudaStreamCreateWithFlags(&p->host_stream, cudaStreamNonBlocking));
  for (int a = 0; a < p->ndim; ++a) {
    if (!p->dev_a[a]) CFD_CUDA_OK(cudaMalloc((void**)&p->dev_a[a], bytes));
    if (!p->dev_b[a]) CFD_CUDA_OK(cudaMalloc((void**)&p->dev_b[a], bytes));
  }
  if (q_out_host && !p->dev_q) CFD_CUDA_OK(cudaMalloc((void**)&p->dev_q, bytes));
  cudaStream_t st = p->host_stream;
  for (int a = 0; a < p->ndim; ++a)
    CFD_CUDA_OK(cudaMemcpyAsync(p->dev_a[a], v_in_host[a], bytes, cudaMemcpyHostToDevice, st));
  float** cur = p->dev_a;
  float** nxt = p->dev_b;
  {
    const int chain = q_out_host ? nsteps - 1 : nsteps;  // the last step is separate when q is wanted
    int in_b = 0;
    if (int e = cfd_repeated(p, st, cur, nxt, chain, params, &in_b)) return e;
    if (in_b) {
      float** tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
    if (q_out_host) {
      if (int e = cfd_step(p, st, cur, nxt, p->dev_q, params)) return e;
      float** tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
  }
  for (int a = 0; a < p->ndim; ++a)
    CFD_CUDA_OK(cudaMemcpyAsync(v_out_host[a], cur[a], bytes, cudaMemcpyDeviceToHost, st));
  if (q_out_host) CFD_CUDA_OK(cudaMemcpyAsync(q_out_host, p->dev_q, bytes, cudaMemcpyDeviceToHost, st));
  CFD_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int cfd_step_profile(cfd_plan* p, cfd_stream stream, const float* const* v_in, float* const* v_out,
                     const cfd_params* params, int reps, int max_kernels, float* ms,
                     const char** names, int* n_kernels) {
  // Times every kernel of a 3-step chain (first step, lazy steps, final materialisation) with
  // CUDA events on the launching stream; reports the mean per launch, aggregated by kernel name
  // in first-seen order.
  DeviceGuard guard_;
  if (int e = check_plan(p)) return e;
  if (reps < 1) reps = 1;
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<const char*> uniq;
  std::vector<double> tot;
  std::vector<int> cnt;
  for (int r = 0; r < reps; ++r) {
    for (auto ev : p->prof_events) cudaEventDestroy(ev);
    p->prof_events.clear();
    p->prof_names.clear();
    p->profiling = true;
    bool single = !lazy_capable(p);
    if (!single) {
      StepConsts c;
      if (int e = make_consts(p, params, &c)) return e;
      single = has_smag(c);
    }
    int e = single ? cfd_step(p, stream, v_in, v_out, nullptr, params)
                   : repeated_lazy(p, st, v_in, v_out, 3, params);
    p->profiling = false;
    if (e) return e;
    CFD_CUDA_OK(cudaStreamSynchronize(st));
    for (size_t i = 1; i < p->prof_events.size(); ++i) {
      const char* nm = p->prof_names[i];
      if (strcmp(nm, "begin") == 0) continue;
      float t = 0.f;
      CFD_CUDA_OK(cudaEventElapsedTime(&t, p->prof_events[i - 1], p->prof_events[i]));
      size_t k = 0;
      for (; k < uniq.size(); ++k)
        if (strcmp(uniq[k], nm) == 0) break;
      if (k == uniq.size()) {
        uniq.push_back(nm);
        tot.push_back(0.0);
        cnt.push_back(0);
      }
      tot[k] += t;
      cnt[k] += 1;
    }
  }
  const int n = (int)uniq.size();
  if (n_kernels) *n_kernels = n < max_kernels ? n : max_kernels;
  for (int i = 0; i < n && i < max_kernels; ++i) {
    ms[i] = (float)(tot[i] / cnt[i]);
    names[i] = uniq[i];
  }
  return 0;
}

// ---- device helpers ---------------------------------------------------------------------------
int cfd_malloc(void** dptr, size_t bytes) {
  CFD_CUDA_OK(cudaMalloc(dptr, bytes));
  return 0;
}
int cfd_free(void* dptr) {
  CFD_CUDA_OK(cudaFree(dptr));
  return 0;
}
int cfd_malloc_host(void** hptr, size_t bytes) {
  CFD_CUDA_OK(cudaMallocHost(hptr, bytes));
  return 0;
}
int cfd_free_host(void* hptr) {
  CFD_CUDA_OK(cudaFreeHost(hptr));
  return 0;
}
int cfd_memcpy_h2d(void* dst, const void* src, size_t bytes, cfd_stream s) {
  CFD_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)s));
  return 0;
}
int cfd_memcpy_d2h(void* dst, const void* src, size_t bytes, cfd_stream s) {
  CFD_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)s));
  return 0;
}
int cfd_memcpy_d2d(void* dst, const void* src, size_t bytes, cfd_stream s) {
  CFD_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)s));
  return 0;
}
int cfd_memset(void* dst, int value, size_t bytes, cfd_stream s) {
  CFD_CUDA_OK(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)s));
  return 0;
}
int cfd_stream_create(cfd_stream* out) {
  cudaStream_t s;
  CFD_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *out = (cfd_stream)s;
  return 0;
}
int cfd_stream_destroy(cfd_stream s) {
  CFD_CUDA_OK(cudaStreamDestroy((cudaStream_t)s));
  return 0;
}
int cfd_stream_sync(cfd_stream s) {
  CFD_CUDA_OK(cudaStreamSynchronize((cudaStream_t)s));
  return 0;
}
int cfd_device_sync(void) {
  CFD_CUDA_OK(cudaDeviceSynchronize());
  return 0;
}
int cfd_set_device(int device) {
  CFD_CUDA_OK(cudaSetDevice(device));
  return 0;
}
int cfd_get_device(int* device) {
  CFD_CUDA_OK(cudaGetDevice(device));
  return 0;
}
int cfd_pointer_device(const void* ptr, int* device) {
  cudaPointerAttributes a;
  CFD_CUDA_OK(cudaPointerGetAttributes(&a, ptr));
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged)
    return set_error_msg("not a device pointer");
  *device = a.device;
  return 0;
}
int cfd_event_create(void** ev) {
  cudaEvent_t e;
  CFD_CUDA_OK(cudaEventCreate(&e));
  *ev = (void*)e;
  return 0;
}
int cfd_event_destroy(void* ev) {
  CFD_CUDA_OK(cudaEventDestroy((cudaEvent_t)ev));
  return 0;
}
int cfd_event_record(void* ev, cfd_stream s) {
  CFD_CUDA_OK(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)s));
  return 0;
}
int cfd_event_elapsed_ms(void* start, void* stop, float* ms) {
  CFD_CUDA_OK(cudaEventSynchronize((cudaEvent_t)stop));
  CFD_CUDA_OK(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return 0;
}

}  // extern "C"
