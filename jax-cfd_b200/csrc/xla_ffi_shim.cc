// XLA FFI (typed FFI, API v1) handlers forwarding to the C ABI of include/cfd_b200.h.
//
// NOT part of the default build: the XLA FFI headers (xla/ffi/api/ffi.h, located by
// jax.ffi.include_dir()) and JAX itself are absent from this image.  Build where JAX is installed:
//   g++ -O2 -std=c++17 -fPIC -shared -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I/usr/local/cuda/include -I../../include xla_ffi_shim.cc -L../lib -lcfd_b200 \
//       -o ../lib/libcfd_b200_xla.so
// In this repository the handler bodies are compiled against a small stand-in for that header and
// run by tests/test_xla_ffi_shim.py (argument validation on the CPU box; the full call against
// cfd_step / cfd_repeated on a GPU), so the logic below is exercised even though XLA is not.
//
// Handlers (registered from jax-cfd_b200/jax_ffi.py with jax.ffi.register_ffi_target):
//   B200CfdStep2D / B200CfdStep3D        step_fn of equations.semi_implicit_navier_stokes
//                                        (equations.py:120-151), `nsteps` steps per call
//                                        (funcutils.repeated, funcutils.py:82-88)
//   B200CfdProject2D / B200CfdProject3D  pressure.projection + solve_fast_diag's q
//                                        (pressure.py:181-198, 115-157)
// Everything that defines the step travels as TYPED SCALAR ATTRIBUTES (serialisable: nothing here
// is a host pointer, so the persistent compilation cache and multi-process XLA work); the tables of
// a separable / constant forcing travel as operands, evaluated by the reference's own jnp
// expressions on the Python side (bit-exact forcing).  The plan (tables + workspace) is created
// lazily per (device, grid, batch, implementation) and cached behind a mutex; XLA owns every
// operand / result buffer and the handlers never write an operand (cfd_advance).
// Not command-buffer compatible: the first call of a plan allocates its workspace.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "cfd_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

struct PlanKey {
  int device, ndim, batch, implementation;
  int64_t shape[CFD_MAX_DIM];
  double step[CFD_MAX_DIM];
  bool operator<(const PlanKey& o) const { return std::memcmp(this, &o, sizeof *this) < 0; }
};

std::mutex g_mu;
std::map<PlanKey, cfd_plan*> g_plans;

ffi::Error invalid(const std::string& msg) { return ffi::Error(ffi::ErrorCode::kInvalidArgument, msg); }
ffi::Error internal() { return ffi::Error(ffi::ErrorCode::kInternal, cfd_last_error()); }

// grid shape = the trailing `ndim` dimensions, batch = the product of the leading ones
// (grids.py:59-63: leading batch / time dimensions are tolerated; jax.vmap with
// vmap_method="broadcast_all" adds them)
ffi::Error grid_of(ffi::Span<const int64_t> dims, int ndim, int64_t* shape, int* batch) {
  if ((int)dims.size() < ndim) return invalid("b200cfd: operand rank is smaller than the grid rank");
  int64_t b = 1;
  for (size_t i = 0; i + ndim < dims.size(); ++i) b *= dims[i];
  for (int j = 0; j < ndim; ++j) shape[j] = dims[dims.size() - ndim + j];
  if (b < 1 || b > (1 << 30)) return invalid("b200cfd: bad batch size");
  *batch = (int)b;
  return ffi::Error::Success();
}

bool same_dims(ffi::Span<const int64_t> a, ffi::Span<const int64_t> b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i)
    if (a[i] != b[i]) return false;
  return true;
}

ffi::Error plan_for(const float* any_operand, int ndim, const int64_t* shape, ffi::Span<const double> step,
                    int batch, int implementation, cfd_plan** out) {
  if ((int)step.size() != ndim) return invalid("b200cfd: `step` needs one entry per grid axis");
  PlanKey k;
  std::memset(&k, 0, sizeof k);
  if (cfd_pointer_device(any_operand, &k.device) != 0) return internal();
  k.ndim = ndim;
  k.batch = batch;
  k.implementation = implementation;
  for (int j = 0; j < ndim; ++j) {
    k.shape[j] = shape[j];
    k.step[j] = step[j];
  }
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_plans.find(k);
  if (it == g_plans.end()) {
    cfd_plan* p = nullptr;
    if (cfd_plan_create_impl(&p, ndim, k.shape, k.step, batch, k.device, implementation) != 0) return internal();
    it = g_plans.emplace(k, p).first;
  }
  *out = it->second;
  return ffi::Error::Success();
}

// Scalar description of the equation (every field an FFI attribute)
struct StepAttrs {
  double dt, convect_dt, density, viscosity;  // viscosity < 0: None (equations.py:106)
  double linear_coef;                         // used when `terms` lists CFD_FORCE_LINEAR
  double smagorinsky_cs;                      // ... CFD_FORCE_SMAGORINSKY
  ffi::Span<const int32_t> terms;             // enum cfd_forcing_kind, in summation order
  ffi::Span<const double> sep_scale;          // per component, CFD_FORCE_SEPARABLE
  int64_t sep_mask;                           // bit a * ndim + j: profile (a, j) present in `sep_prof`
  int64_t sep_has;                            // bit a: component a forced by the separable term
  int64_t field_mask;                         // bit a: component a present in `field`
};

ffi::Error fill_params(int ndim, const int64_t* shape, const StepAttrs& a, ffi::Buffer<ffi::F32>& sep_prof,
                       ffi::Buffer<ffi::F32>& field, cfd_params* prm) {
  std::memset(prm, 0, sizeof *prm);
  prm->dt = a.dt;
  prm->convect_dt = a.convect_dt;
  prm->density = a.density;
  prm->has_viscosity = a.viscosity >= 0 ? 1 : 0;
  prm->viscosity = a.viscosity >= 0 ? a.viscosity : 0.0;
  prm->linear_coef = a.linear_coef;
  prm->smagorinsky_cs = a.smagorinsky_cs;
  if (a.terms.size() > CFD_MAX_FORCING_TERMS) return invalid("b200cfd: too many forcing terms");
  prm->n_terms = (int32_t)a.terms.size();
  for (size_t t = 0; t < a.terms.size(); ++t) prm->term_kind[t] = a.terms[t];
  // separable profiles: rows of `sep_prof` (n_present, max N), in (a, j) order of the set bits
  int64_t nmax = 0;
  for (int j = 0; j < ndim; ++j) nmax = shape[j] > nmax ? shape[j] : nmax;
  int row = 0;
  for (int c = 0; c < ndim; ++c) {
    prm->has_sep[c] = (a.sep_has >> c) & 1;
    prm->sep_scale[c] = c < (int)a.sep_scale.size() ? (float)a.sep_scale[c] : 0.f;
    for (int j = 0; j < ndim; ++j)
      if ((a.sep_mask >> (c * ndim + j)) & 1) prm->sep_prof[c][j] = sep_prof.typed_data() + (row++) * nmax;
  }
  if (row > 0) {
    auto d = sep_prof.dimensions();
    if (d.size() != 2 || d[0] != row || d[1] != nmax)
      return invalid("b200cfd: `sep_prof` must have shape (number of profiles, max grid extent)");
  }
  // constant fields: `field` has shape (n_present, *grid), components in the order of the set bits
  int64_t cells = 1;
  for (int j = 0; j < ndim; ++j) cells *= shape[j];
  int nf = 0;
  for (int c = 0; c < ndim; ++c)
    if ((a.field_mask >> c) & 1) prm->field[c] = field.typed_data() + (nf++) * cells;
  if (nf > 0 && (int64_t)field.element_count() != nf * cells)
    return invalid("b200cfd: `field` must have shape (number of forced components, *grid)");
  return ffi::Error::Success();
}

template <int NDIM>
ffi::Error StepImpl(cudaStream_t stream, const ffi::Buffer<ffi::F32>* v, ffi::Buffer<ffi::F32>& sep_prof,
                    ffi::Buffer<ffi::F32>& field, ffi::Result<ffi::Buffer<ffi::F32>>* out,
                    ffi::Span<const double> step, int32_t nsteps, int32_t implementation, const StepAttrs& a) {
  int64_t shape[CFD_MAX_DIM];
  int batch = 1;
  if (auto e = grid_of(v[0].dimensions(), NDIM, shape, &batch); e.failure()) return e;
  for (int c = 0; c < NDIM; ++c)
    if (!same_dims(v[c].dimensions(), v[0].dimensions()) || !same_dims(out[c]->dimensions(), v[0].dimensions()))
      return invalid("b200cfd_step: operands and results must all have the same shape");
  if (nsteps < 1) return invalid("b200cfd_step: nsteps must be >= 1");
  cfd_plan* plan = nullptr;
  if (auto e = plan_for(v[0].typed_data(), NDIM, shape, step, batch, implementation, &plan); e.failure()) return e;
  cfd_params prm;
  if (auto e = fill_params(NDIM, shape, a, sep_prof, field, &prm); e.failure()) return e;
  const float* in[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};
  float* res[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};
  for (int c = 0; c < NDIM; ++c) {
    in[c] = v[c].typed_data();
    res[c] = out[c]->typed_data();
  }
  // one plan = one workspace: calls that share it are serialised on the host (XLA may invoke the
  // handler from several host threads; on one stream they are ordered on the device anyway)
  std::lock_guard<std::mutex> lock(g_mu);
  if (cfd_advance(plan, stream, in, res, nsteps, &prm) != 0) return internal();
  return ffi::Error::Success();
}

template <int NDIM>
ffi::Error ProjectImpl(cudaStream_t stream, const ffi::Buffer<ffi::F32>* v, ffi::Result<ffi::Buffer<ffi::F32>>* out,
                       ffi::Result<ffi::Buffer<ffi::F32>>& q, ffi::Span<const double> step,
                       int32_t implementation) {
  int64_t shape[CFD_MAX_DIM];
  int batch = 1;
  if (auto e = grid_of(v[0].dimensions(), NDIM, shape, &batch); e.failure()) return e;
  for (int c = 0; c < NDIM; ++c)
    if (!same_dims(v[c].dimensions(), v[0].dimensions()) || !same_dims(out[c]->dimensions(), v[0].dimensions()))
      return invalid("b200cfd_project: operands and results must all have the same shape");
  if (!same_dims(q->dimensions(), v[0].dimensions())) return invalid("b200cfd_project: q must have the operand shape");
  cfd_plan* plan = nullptr;
  if (auto e = plan_for(v[0].typed_data(), NDIM, shape, step, batch, implementation, &plan); e.failure()) return e;
  const float* in[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};
  float* res[CFD_MAX_DIM] = {nullptr, nullptr, nullptr};
  for (int c = 0; c < NDIM; ++c) {
    in[c] = v[c].typed_data();
    res[c] = out[c]->typed_data();
  }
  std::lock_guard<std::mutex> lock(g_mu);
  if (cfd_project(plan, stream, in, res, q->typed_data()) != 0) return internal();
  return ffi::Error::Success();
}

}  // namespace

// ---- 2-D --------------------------------------------------------------------------------------------
ffi::Error B200CfdStep2DImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> u, ffi::Buffer<ffi::F32> v,
                             ffi::Buffer<ffi::F32> sep_prof, ffi::Buffer<ffi::F32> field,
                             ffi::Result<ffi::Buffer<ffi::F32>> uo, ffi::Result<ffi::Buffer<ffi::F32>> vo,
                             ffi::Span<const double> step, int32_t nsteps, int32_t implementation, double dt,
                             double convect_dt, double density, double viscosity, double linear_coef,
                             double smagorinsky_cs, ffi::Span<const int32_t> terms, ffi::Span<const double> sep_scale,
                             int64_t sep_mask, int64_t sep_has, int64_t field_mask) {
  const ffi::Buffer<ffi::F32> in[2] = {u, v};
  ffi::Result<ffi::Buffer<ffi::F32>> out[2] = {uo, vo};
  const StepAttrs a = {dt, convect_dt, density, viscosity, linear_coef, smagorinsky_cs, terms, sep_scale,
                       sep_mask, sep_has, field_mask};
  return StepImpl<2>(stream, in, sep_prof, field, out, step, nsteps, implementation, a);
}

ffi::Error B200CfdProject2DImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> u, ffi::Buffer<ffi::F32> v,
                                ffi::Result<ffi::Buffer<ffi::F32>> uo, ffi::Result<ffi::Buffer<ffi::F32>> vo,
                                ffi::Result<ffi::Buffer<ffi::F32>> q, ffi::Span<const double> step,
                                int32_t implementation) {
  const ffi::Buffer<ffi::F32> in[2] = {u, v};
  ffi::Result<ffi::Buffer<ffi::F32>> out[2] = {uo, vo};
  return ProjectImpl<2>(stream, in, out, q, step, implementation);
}

// ---- 3-D --------------------------------------------------------------------------------------------
ffi::Error B200CfdStep3DImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> u, ffi::Buffer<ffi::F32> v,
                             ffi::Buffer<ffi::F32> w, ffi::Buffer<ffi::F32> sep_prof, ffi::Buffer<ffi::F32> field,
                             ffi::Result<ffi::Buffer<ffi::F32>> uo, ffi::Result<ffi::Buffer<ffi::F32>> vo,
                             ffi::Result<ffi::Buffer<ffi::F32>> wo, ffi::Span<const double> step, int32_t nsteps,
                             int32_t implementation, double dt, double convect_dt, double density,
                             double viscosity, double linear_coef, double smagorinsky_cs,
                             ffi::Span<const int32_t> terms, ffi::Span<const double> sep_scale, int64_t sep_mask,
                             int64_t sep_has, int64_t field_mask) {
  const ffi::Buffer<ffi::F32> in[3] = {u, v, w};
  ffi::Result<ffi::Buffer<ffi::F32>> out[3] = {uo, vo, wo};
  const StepAttrs a = {dt, convect_dt, density, viscosity, linear_coef, smagorinsky_cs, terms, sep_scale,
                       sep_mask, sep_has, field_mask};
  return StepImpl<3>(stream, in, sep_prof, field, out, step, nsteps, implementation, a);
}

ffi::Error B200CfdProject3DImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> u, ffi::Buffer<ffi::F32> v,
                                ffi::Buffer<ffi::F32> w, ffi::Result<ffi::Buffer<ffi::F32>> uo,
                                ffi::Result<ffi::Buffer<ffi::F32>> vo, ffi::Result<ffi::Buffer<ffi::F32>> wo,
                                ffi::Result<ffi::Buffer<ffi::F32>> q, ffi::Span<const double> step,
                                int32_t implementation) {
  const ffi::Buffer<ffi::F32> in[3] = {u, v, w};
  ffi::Result<ffi::Buffer<ffi::F32>> out[3] = {uo, vo, wo};
  return ProjectImpl<3>(stream, in, out, q, step, implementation);
}

#define CFD_STEP_ATTRS()                         \
  .Attr<ffi::Span<const double>>("step")         \
      .Attr<int32_t>("nsteps")                   \
      .Attr<int32_t>("implementation")           \
      .Attr<double>("dt")                        \
      .Attr<double>("convect_dt")                \
      .Attr<double>("density")                   \
      .Attr<double>("viscosity")                 \
      .Attr<double>("linear_coef")               \
      .Attr<double>("smagorinsky_cs")            \
      .Attr<ffi::Span<const int32_t>>("terms")   \
      .Attr<ffi::Span<const double>>("sep_scale") \
      .Attr<int64_t>("sep_mask")                 \
      .Attr<int64_t>("sep_has")                  \
      .Attr<int64_t>("field_mask")

XLA_FFI_DEFINE_HANDLER_SYMBOL(B200CfdStep2D, B200CfdStep2DImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()  // u
                                  .Arg<ffi::Buffer<ffi::F32>>()  // v
                                  .Arg<ffi::Buffer<ffi::F32>>()  // sep_prof (may be empty)
                                  .Arg<ffi::Buffer<ffi::F32>>()  // field (may be empty)
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>() CFD_STEP_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(B200CfdStep3D, B200CfdStep3DImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()  // sep_prof
                                  .Arg<ffi::Buffer<ffi::F32>>()  // field
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>() CFD_STEP_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(B200CfdProject2D, B200CfdProject2DImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()  // q
                                  .Attr<ffi::Span<const double>>("step")
                                  .Attr<int32_t>("implementation"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(B200CfdProject3D, B200CfdProject3DImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()  // q
                                  .Attr<ffi::Span<const double>>("step")
                                  .Attr<int32_t>("implementation"));
