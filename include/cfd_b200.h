/* cfd_b200.h -- C ABI of the B200-native staggered-grid FVM time step.
 *
 * Drop-in boundary for the `step_fn` returned by
 *   jax_cfd.base.equations.semi_implicit_navier_stokes      (equations.py:120-151)
 * and the pieces that callers either side of it use:
 *   jax_cfd.base.pressure.projection / solve_fast_diag       (pressure.py:181-198, 115-157)
 *   jax_cfd.base.equations.dynamic_time_step's max-speed     (equations.py:62-70)
 *   jax_cfd.data.xarray_utils kinetic_energy / enstrophy_2d  (data/xarray_utils.py:155-188)
 *
 * The reference has no FFI of its own (it is pure Python on XLA).  These entry points are what
 * an XLA-FFI handler registered through `jax.ffi.register_ffi_target` forwards to (see
 * INTEGRATION.md and jax-cfd_b200/csrc/xla_ffi_shim.cc); in this JAX-less image the same
 * symbols are driven through ctypes by the Python mirror of the reference interface
 * (jax-cfd_b200/equations.py).
 *
 * Conventions: plain pointers and sizes, no framework types.  Unless a function says "host",
 * every data pointer is a DEVICE pointer to float32, row-major, axis 0 slowest, with an optional
 * leading batch axis: (batch, N0, N1[, N2]).  Velocity component `a` lives at offset
 * grid.cell_faces[a] (grids.py:567-572); pressure at the cell centre.  All boundaries periodic.
 * Every call is enqueue-only on the given stream unless it says it synchronises, runs on the
 * plan's device and restores the caller's current device before returning.  A plan owns ONE
 * workspace: calls on the same plan must be ordered on one stream (use one plan per stream).
 * Return value: 0 on success, non-zero on error (message via cfd_last_error()).  There is no
 * CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef CFD_B200_H_
#define CFD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFD_MAX_DIM 3
#define CFD_MAX_FORCING_TERMS 4

typedef struct cfd_plan cfd_plan;
typedef void* cfd_stream; /* cudaStream_t */

/* Forcing terms, summed left to right starting from 0 exactly like
 * forcings.sum_forcings (forcings.py:125-129) / equations.sum_fields (equations.py:41-42). */
enum cfd_forcing_kind {
  CFD_FORCE_NONE = 0,
  CFD_FORCE_SEPARABLE = 1,  /* f_a = (prod_j prof[a][j][i_j]) * scale[a]   (kolmogorov_forcing
                               forcings.py:63-104, taylor_green_forcing forcings.py:35-60) */
  CFD_FORCE_FIELD = 2,      /* f_a = field[a][...]  (any constant ForcingFn evaluated by the host) */
  CFD_FORCE_LINEAR = 3,     /* f_a = coef * v_a      (linear_forcing forcings.py:107-113) */
  CFD_FORCE_SMAGORINSKY = 4 /* evm_model(smagorinsky_viscosity) subgrid_models.py:40-134 */
};

typedef struct cfd_params {
  double dt;            /* time step (time_stepping.py:101) */
  double density;       /* equations.py:121 */
  double viscosity;     /* equations.py:122; ignored when has_viscosity == 0 (viscosity=None) */
  int32_t has_viscosity;
  int32_t n_terms;                           /* number of forcing terms, <= CFD_MAX_FORCING_TERMS */
  int32_t term_kind[CFD_MAX_FORCING_TERMS];  /* enum cfd_forcing_kind, in summation order */
  double linear_coef;                        /* CFD_FORCE_LINEAR */
  double smagorinsky_cs;                     /* CFD_FORCE_SMAGORINSKY */
  /* CFD_FORCE_SEPARABLE: device float tables prof[a][j] of length shape[j] (NULL = all ones),
   * scale[a]; has_sep[a]==0 means component a of the term is identically zero. */
  const float* sep_prof[CFD_MAX_DIM][CFD_MAX_DIM];
  float sep_scale[CFD_MAX_DIM];
  int32_t has_sep[CFD_MAX_DIM];
  /* CFD_FORCE_FIELD: device arrays of the grid shape (no batch axis), NULL = zero. */
  const float* field[CFD_MAX_DIM];
  /* dt inside the Lax-Wendroff Courant number of the convection term (equations.py:127-128 binds
   * `convect` to the equation builder's dt; the time stepper may be given another one,
   * time_stepping.py:59-106).  0 = same as dt. */
  double convect_dt;
} cfd_params;

typedef struct cfd_diag {
  double kinetic_energy; /* mean over cells of 0.5 * sum_a v_a^2 (raw staggered samples) */
  double enstrophy;      /* 2-D: mean of 0.5 * (D+_x v - D+_y u)^2 ; 3-D: mean 0.5*|curl|^2 */
  double max_abs_div;    /* max |finite_differences.divergence(v)| */
  double max_speed_sq;   /* max over cells of sum_a v_a^2  (equations.py:68) */
} cfd_diag;

/* ---- library / device ------------------------------------------------------------------ */
const char* cfd_last_error(void);
const char* cfd_version(void);
int cfd_device_count(void);           /* 0 when no CUDA device is usable */

/* ---- plan: grid metadata + eigenvalue / twiddle tables + workspace ------------------------
 * Replaces the trace-time work of pressure.solve_fast_diag (pressure.py:115-157),
 * array_utils.laplacian_matrix (array_utils.py:168-173; built analytically here) and
 * fast_diagonalization.pseudoinverse (fast_diagonalization.py:228-266).  `step[j]` = grid.step[j].
 *
 * Implementation of the fast-diagonalisation transform (fast_diagonalization.py:90-125):
 *   CFD_IMPL_RFFT    shared-memory line FFTs ('rfft'; 'fft' computes the same transform for real
 *                    input and maps here too); needs every axis a power of two >= 16 (>= 32 on the
 *                    last axis)
 *   CFD_IMPL_MATMUL  eigenvector products along each axis on the FP64 tensor cores ('matmul',
 *                    fast_diagonalization.py:129-165); any shape with axes of up to 4096 cells
 *   CFD_IMPL_AUTO    rfft when the shape allows it, else matmul -- the reference's own fallback
 *                    for odd last axes (fast_diagonalization.py:107-108), extended to every shape
 *                    the radix-2 line kernels do not take
 * cfd_plan_create == cfd_plan_create_impl(..., CFD_IMPL_AUTO). */
enum cfd_implementation { CFD_IMPL_AUTO = 0, CFD_IMPL_RFFT = 1, CFD_IMPL_MATMUL = 2 };
int cfd_plan_create(cfd_plan** out, int ndim, const int64_t* shape, const double* step,
                    int batch, int device);
int cfd_plan_create_impl(cfd_plan** out, int ndim, const int64_t* shape, const double* step,
                         int batch, int device, int implementation);
int cfd_plan_implementation(const cfd_plan* plan); /* CFD_IMPL_RFFT or CFD_IMPL_MATMUL */
void cfd_plan_destroy(cfd_plan* plan);
size_t cfd_plan_workspace_bytes(const cfd_plan* plan);

/* ---- the hot path ------------------------------------------------------------------------
 * One forward-Euler projection step  v' = P(v + dt F(v))   (time_stepping.py:88-104,109-118).
 * v_in / v_out: ndim device pointers each; v_out must not alias v_in.  q_out (nullable)
 * receives solve_fast_diag's `q` (pressure.py:151-157). */
int cfd_step(cfd_plan* plan, cfd_stream stream, const float* const* v_in, float* const* v_out,
             float* q_out, const cfd_params* params);

/* `nsteps` steps, ping-ponging between v_a and v_b (funcutils.repeated, funcutils.py:82-88);
 * the result is in v_a if nsteps is even, v_b if odd.  Returns which through *result_in_b. */
int cfd_repeated(cfd_plan* plan, cfd_stream stream, float* const* v_a, float* const* v_b,
                 int nsteps, const cfd_params* params, int* result_in_b);

/* `nsteps` >= 1 steps from v_in into v_out without ever writing v_in (operands of an XLA custom
 * call are immutable): what csrc/xla_ffi_shim.cc forwards to.  v_out must not alias v_in. */
int cfd_advance(cfd_plan* plan, cfd_stream stream, const float* const* v_in, float* const* v_out,
                int nsteps, const cfd_params* params);

/* dv/dt = conv + (nu/rho) lap + forcing/rho  (equations.navier_stokes_explicit_terms,
 * equations.py:77-116) -- the F of navier_stokes_rk (time_stepping.py:59-106). */
int cfd_explicit_terms(cfd_plan* plan, cfd_stream stream, const float* const* v_in,
                       float* const* dvdt_out, const cfd_params* params);

/* pressure.projection (pressure.py:181-198): v_out = v_in - grad(q), q = pinv(div v_in).
 * v_out may alias v_in. */
int cfd_project(cfd_plan* plan, cfd_stream stream, const float* const* v_in,
                float* const* v_out, float* q_out);

/* fast_diagonalization.transform (fast_diagonalization.py:28-126) on one real field of the plan's
 * grid (batched like everything else):  out = X diag X^-1 in.
 *   rfft plans:   circulant symmetric operators.  diag_lines = func(eigenvalue sums) in rfftn layout
 *                 (N0, [N1,] N_last/2 + 1), TRANSPOSED to line layout (N_last/2 + 1, [N1,] N0), device
 *                 float32 (the reference narrows `diagonals` to complex64, fast_diagonalization.py:214).
 *                 Used by diffusion.solve_fast_diag (diffusion.py:166-212) and by the spectral filter of
 *                 initial_conditions.filtered_velocity_field (filter_utils.py:32-42).  in may alias out.
 *   matmul plans: any hermitian operators.  eigvecs[j] / eigvecs_t[j] = device float64 N_j x N_j
 *                 row-major eigenvector matrix of axis j (columns = eigenvectors, np.linalg.eigh) and its
 *                 transpose; diag = device float64 table of the grid shape in that eigenbasis. */
int cfd_transform_rfft(cfd_plan* plan, cfd_stream stream, const float* in, float* out,
                       const float* diag_lines);
int cfd_transform_matmul(cfd_plan* plan, cfd_stream stream, const float* in, float* out,
                         const double* const* eigvecs, const double* const* eigvecs_t,
                         const double* diag);

/* out = numer * x / denom per component, in float32 in that order (the normalisation step of
 * initial_conditions.filtered_velocity_field, initial_conditions.py:118-121).  out may alias x. */
int cfd_scale(cfd_plan* plan, cfd_stream stream, const float* const* x, double numer, double denom,
              float* const* out);

/* Post-processing either side of the path (`trajectory(post_process=...)`, funcutils.py:118-121), plan-free:
 * resize.downsample_staggered_velocity_component (resize.py:38-74) -- of the fine faces normal to
 * `direction` keep those on a coarse face and average the factor^(ndim-1) that tile it (a
 * divergence-free field stays divergence free); in: (batch, *shape), out: (batch, *shape / factor).
 * `factor` must divide every axis other than `direction` (array_utils.block_reduce,
 * array_utils.py:155-160); along `direction` a remainder is dropped like the reference's slice. */
int cfd_downsample_component(cfd_stream stream, const float* in, float* out, int ndim,
                             const int64_t* shape, int batch, int direction, int factor);
/* vorticity_2d at offset (1, 1): (S(v,+1,x) - v) / dx - (S(u,+1,y) - u) / dy, periodic
 * (data/xarray_utils.py:155-163). */
int cfd_vorticity_2d(cfd_stream stream, const float* u, const float* v, float* out,
                     const int64_t* shape, int batch, double dx, double dy);

/* out = x + sum_k coef[k] * y[k]  per component (the stage combinations of navier_stokes_rk,
 * time_stepping.py:96-101).  y is an array of nterms pointers-to-component-arrays. */
int cfd_axpy(cfd_plan* plan, cfd_stream stream, const float* const* x, int nterms,
             const float* const* const* y, const double* coef, float* const* out);

/* Diagnostics: enqueues the reduction and synchronises the stream before returning. */
int cfd_diagnostics(cfd_plan* plan, cfd_stream stream, const float* const* v, cfd_diag* out);

/* ---- host-buffer entry points (the call a reference user makes: numpy in, numpy out) ------
 * Copies the state host->device (pinned staging inside the plan), runs nsteps, copies the result
 * back, synchronises.  q_out_host nullable. */
int cfd_step_host(cfd_plan* plan, const float* const* v_in_host, float* const* v_out_host,
                  float* q_out_host, int nsteps, const cfd_params* params);

/* ---- slab-decomposed multi-GPU step (one process per GPU) -----------
 * The reference has no distributed solver; this is the new path for
 * BASELINE configs #4 (2-D, 32768^2) and #5 (3-D, 512^3 with the Smagorinsky closure:
 * cfd_dist_plan_create_nd with ndim = 3).  The grid is split along axis 0; halo rows / planes and
 * the FFT transposes move over NVLink through CUDA-IPC peer mappings inside the kernels.
 * Protocol per rank:
 *   cfd_dist_plan_create -> cfd_dist_export (64-byte blob) -> [host all-gathers the blobs]
 *   -> cfd_dist_connect(all blobs) -> cfd_dist_load(local rows) -> cfd_dist_advance(n) ...
 *   -> cfd_dist_store(local rows).  Every rank must make the same calls (device-side barriers). */
int cfd_dist_plan_create(cfd_plan** out, const int64_t* global_shape, const double* step, int rank,
                         int world, int device); /* 2-D */
int cfd_dist_plan_create_nd(cfd_plan** out, int ndim, const int64_t* global_shape, const double* step,
                            int rank, int world, int device); /* ndim = 2 or 3 */
size_t cfd_dist_handle_bytes(void);
int cfd_dist_export(cfd_plan* plan, void* blob);
int cfd_dist_connect(cfd_plan* plan, const void* all_blobs);
int cfd_dist_load(cfd_plan* plan, cfd_stream stream, const float* const* v_local);
int cfd_dist_advance(cfd_plan* plan, cfd_stream stream, int nsteps, const cfd_params* params);
int cfd_dist_store(cfd_plan* plan, cfd_stream stream, float* const* v_local_out, float* q_local_out);
int cfd_dist_check(cfd_plan* plan); /* non-zero if a device-side barrier ever timed out */
int cfd_dist_profile(cfd_plan* plan, cfd_stream stream, int nsteps, const cfd_params* params,
                     int max_kernels, float* ms, const char** names, int* n_kernels);

/* ---- thin device-memory helpers so hosts without a CUDA array library can drive the ABI --- */
int cfd_malloc(void** dptr, size_t bytes);
int cfd_free(void* dptr);
int cfd_malloc_host(void** hptr, size_t bytes); /* pinned */
int cfd_free_host(void* hptr);
int cfd_memcpy_h2d(void* dst, const void* src, size_t bytes, cfd_stream stream);
int cfd_memcpy_d2h(void* dst, const void* src, size_t bytes, cfd_stream stream);
int cfd_memcpy_d2d(void* dst, const void* src, size_t bytes, cfd_stream stream);
int cfd_memset(void* dst, int value, size_t bytes, cfd_stream stream);
int cfd_stream_create(cfd_stream* out);
int cfd_stream_destroy(cfd_stream s);
int cfd_stream_sync(cfd_stream s);
int cfd_device_sync(void);
int cfd_set_device(int device);
int cfd_get_device(int* device);
int cfd_pointer_device(const void* ptr, int* device); /* device that owns a device pointer */
/* CUDA-event timing on the launching stream (bench.py). */
int cfd_event_create(void** ev);
int cfd_event_destroy(void* ev);
int cfd_event_record(void* ev, cfd_stream s);
int cfd_event_elapsed_ms(void* start, void* stop, float* ms); /* synchronises on `stop` */
/* Per-kernel CUDA-event timing of one step (bench.py roofline): fills ms[0..n) for the kernels
 * of cfd_step in launch order and returns their names (static strings). */
int cfd_step_profile(cfd_plan* plan, cfd_stream stream, const float* const* v_in,
                     float* const* v_out, const cfd_params* params, int reps, int max_kernels,
                     float* ms, const char** names, int* n_kernels);
/* Number of kernel launches this library has made since load (bench.py "gpu_launches"). */
uint64_t cfd_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CFD_B200_H_ */
