// Drives the handler bodies of csrc/xla_ffi_shim.cc (compiled against tests/host/xla_stub) the way
// XLA would: operands / results as ffi::Buffer, every parameter as a typed attribute.
//   harness cpu   argument validation + error mapping (no GPU needed)
//   harness gpu   B200CfdStep2D / 3D with nsteps = 1, 2, 3 against cfd_step / cfd_repeated through the
//                 plain C ABI (bit-identical), Project2D against cfd_project, operands untouched
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "cfd_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using Buf = ffi::Buffer<ffi::F32>;
using Res = ffi::Result<Buf>;

ffi::Error B200CfdStep2DImpl(cudaStream_t, Buf, Buf, Buf, Buf, Res, Res, ffi::Span<const double>, int32_t, int32_t,
                             double, double, double, double, double, double, ffi::Span<const int32_t>,
                             ffi::Span<const double>, int64_t, int64_t, int64_t);
ffi::Error B200CfdStep3DImpl(cudaStream_t, Buf, Buf, Buf, Buf, Buf, Res, Res, Res, ffi::Span<const double>, int32_t,
                             int32_t, double, double, double, double, double, double, ffi::Span<const int32_t>,
                             ffi::Span<const double>, int64_t, int64_t, int64_t);
ffi::Error B200CfdProject2DImpl(cudaStream_t, Buf, Buf, Res, Res, Res, ffi::Span<const double>, int32_t);

static int fails = 0;
#define EXPECT(cond, what)                                \
  do {                                                    \
    if (!(cond)) {                                        \
      printf("FAIL: %s (%s:%d)\n", what, __FILE__, __LINE__); \
      ++fails;                                            \
    }                                                     \
  } while (0)

static int run_cpu() {
  std::vector<double> step = {0.1, 0.1};
  std::vector<int32_t> terms;
  std::vector<double> scale = {0, 0};
  float dummy[4];
  Buf u(dummy, {64, 32}), v(dummy, {64, 32}), bad(dummy, {32, 32}), empty(nullptr, {0});
  // mismatching operand shapes -> kInvalidArgument, before anything touches a device
  auto e = B200CfdStep2DImpl(nullptr, u, bad, empty, empty, Res(u), Res(v), step, 1, 0, 0.01, 0, 1.0, 1e-3, 0, 0, terms,
                             scale, 0, 0, 0);
  EXPECT(e.failure() && e.errc() == ffi::ErrorCode::kInvalidArgument, "shape mismatch must be kInvalidArgument");
  e = B200CfdStep2DImpl(nullptr, u, v, empty, empty, Res(u), Res(v), step, 0, 0, 0.01, 0, 1.0, 1e-3, 0, 0, terms, scale, 0,
                        0, 0);
  EXPECT(e.failure() && e.errc() == ffi::ErrorCode::kInvalidArgument, "nsteps = 0 must be kInvalidArgument");
  Buf r1(dummy, {32});
  e = B200CfdStep2DImpl(nullptr, r1, r1, empty, empty, Res(r1), Res(r1), step, 1, 0, 0.01, 0, 1.0, 1e-3, 0, 0, terms,
                        scale, 0, 0, 0);
  EXPECT(e.failure() && e.errc() == ffi::ErrorCode::kInvalidArgument, "rank-1 operand must be kInvalidArgument");
  std::vector<double> step3 = {0.1, 0.1, 0.1};
  e = B200CfdStep2DImpl(nullptr, u, v, empty, empty, Res(u), Res(v), step3, 1, 0, 0.01, 0, 1.0, 1e-3, 0, 0, terms, scale,
                        0, 0, 0);
  EXPECT(e.failure(), "wrong `step` length must fail");
  if (cfd_device_count() == 0) {
    // host pointers / no device: the C ABI refuses, the handler maps it to kInternal with the message
    e = B200CfdStep2DImpl(nullptr, u, v, empty, empty, Res(u), Res(v), step, 2, 0, 0.01, 0, 1.0, 1e-3, 0, 0, terms, scale,
                          0, 0, 0);
    EXPECT(e.failure() && e.errc() == ffi::ErrorCode::kInternal && !e.message().empty(),
           "no CUDA device must surface as kInternal with cfd_last_error()");
  }
  printf(fails ? "XLA SHIM CPU FAIL\n" : "XLA SHIM CPU PASS\n");
  return fails;
}

static float* dev(const std::vector<float>& h) {
  void* d = nullptr;
  cfd_malloc(&d, h.size() * sizeof(float));
  cfd_memcpy_h2d(d, h.data(), h.size() * sizeof(float), nullptr);
  cfd_stream_sync(nullptr);
  return (float*)d;
}
static std::vector<float> host(const float* d, size_t n) {
  std::vector<float> h(n);
  cfd_memcpy_d2h(h.data(), d, n * sizeof(float), nullptr);
  cfd_stream_sync(nullptr);
  return h;
}
static std::vector<float> field(size_t n, int seed) {
  std::vector<float> h(n);
  unsigned s = 12345u + 977u * seed;
  for (auto& x : h) {
    s = s * 1664525u + 1013904223u;
    x = ((s >> 8) & 0xffff) / 65536.0f - 0.5f;
  }
  return h;
}

template <int ND>
static void run_step_case(const std::vector<int64_t>& dims, int batch_dims) {
  const int nd_all = (int)dims.size();
  size_t n = 1;
  for (auto d : dims) n *= (size_t)d;
  int64_t shape[3];
  double step[3];
  int batch = 1;
  for (int i = 0; i < batch_dims; ++i) batch *= (int)dims[i];
  for (int j = 0; j < ND; ++j) {
    shape[j] = dims[nd_all - ND + j];
    step[j] = 6.283185307179586 / (double)shape[j];
  }
  std::vector<double> stepv(step, step + ND);
  // Kolmogorov (separable, component 0 varies along axis 1) + linear forcing, as attributes + operand
  int64_t nmax = 0;
  for (int j = 0; j < ND; ++j) nmax = shape[j] > nmax ? shape[j] : nmax;
  std::vector<float> prof(nmax, 0.f);
  for (int64_t j = 0; j < shape[1]; ++j) prof[j] = sinf(4.f * (float)((j + 0.5) * step[1]));
  float* dprof = dev(prof);
  std::vector<int32_t> terms = {CFD_FORCE_SEPARABLE, CFD_FORCE_LINEAR};
  std::vector<double> scale(ND, 0.0);
  scale[0] = 1.0;
  const int64_t sep_mask = 1ll << (0 * ND + 1), sep_has = 1;
  const double dt = 0.2 * step[0], nu = 1e-3;
  float* in[3] = {nullptr, nullptr, nullptr};
  std::vector<std::vector<float>> h_in;
  for (int c = 0; c < ND; ++c) {
    h_in.push_back(field(n, c));
    in[c] = dev(h_in.back());
  }
  cfd_plan* plan = nullptr;
  if (cfd_plan_create(&plan, ND, shape, step, batch, 0) != 0) {
    printf("plan: %s\n", cfd_last_error());
    ++fails;
    return;
  }
  cfd_params prm;
  memset(&prm, 0, sizeof prm);
  prm.dt = dt;
  prm.density = 1.0;
  prm.viscosity = nu;
  prm.has_viscosity = 1;
  prm.n_terms = 2;
  prm.term_kind[0] = CFD_FORCE_SEPARABLE;
  prm.term_kind[1] = CFD_FORCE_LINEAR;
  prm.linear_coef = -0.1;
  prm.sep_prof[0][1] = dprof;
  prm.sep_scale[0] = 1.f;
  prm.has_sep[0] = 1;
  for (int nsteps = 1; nsteps <= 3; ++nsteps) {
    // reference result through the plain C ABI
    float *a[3], *b[3], *o[3];
    for (int c = 0; c < ND; ++c) {
      a[c] = dev(h_in[c]);
      b[c] = dev(h_in[c]);
      o[c] = dev(h_in[c]);
    }
    int in_b = 0;
    if (cfd_repeated(plan, nullptr, a, b, nsteps, &prm, &in_b) != 0) {
      printf("cfd_repeated: %s\n", cfd_last_error());
      ++fails;
    }
    std::vector<Buf> bi;
    std::vector<Res> bo;
    for (int c = 0; c < ND; ++c) {
      bi.emplace_back(in[c], dims);
      bo.emplace_back(Buf(o[c], dims));
    }
    Buf sp(dprof, {1, nmax}), empty(nullptr, {0});
    ffi::Error e;
    if (ND == 2)
      e = B200CfdStep2DImpl(nullptr, bi[0], bi[1], sp, empty, bo[0], bo[1], stepv, nsteps, 0, dt, 0.0, 1.0, nu, -0.1, 0.0,
                            terms, scale, sep_mask, sep_has, 0);
    else
      e = B200CfdStep3DImpl(nullptr, bi[0], bi[1], bi[2], sp, empty, bo[0], bo[1], bo[2], stepv, nsteps, 0, dt, 0.0, 1.0,
                            nu, -0.1, 0.0, terms, scale, sep_mask, sep_has, 0);
    if (e.failure()) {
      printf("handler failed: %s\n", e.message().c_str());
      ++fails;
    }
    cfd_device_sync();
    for (int c = 0; c < ND; ++c) {
      auto want = host(in_b ? b[c] : a[c], n), got = host(o[c], n), op = host(in[c], n);
      EXPECT(memcmp(want.data(), got.data(), n * sizeof(float)) == 0, "handler result differs from cfd_repeated");
      EXPECT(memcmp(op.data(), h_in[c].data(), n * sizeof(float)) == 0, "handler wrote an operand");
      cfd_free(a[c]);
      cfd_free(b[c]);
      cfd_free(o[c]);
    }
  }
  printf("step %dD dims[0]=%lld batch=%d: %s\n", ND, (long long)dims[0], batch, fails ? "FAIL" : "ok");
  cfd_plan_destroy(plan);
}

static int run_gpu() {
  run_step_case<2>({64, 32}, 0);
  run_step_case<2>({3, 64, 32}, 1);   // leading batch dimension (jax.vmap, vmap_method="broadcast_all")
  run_step_case<2>({48, 36}, 0);      // matmul plan, plain ping-pong through the scratch state
  run_step_case<3>({16, 16, 32}, 0);
  // projection
  {
    std::vector<int64_t> dims = {64, 32};
    const size_t n = 64 * 32;
    int64_t shape[2] = {64, 32};
    double step[2] = {0.1, 0.2};
    std::vector<double> stepv = {0.1, 0.2};
    auto hu = field(n, 7), hv = field(n, 8);
    float *u = dev(hu), *v = dev(hv), *uo = dev(hu), *vo = dev(hv), *q = dev(hu);
    float *ru = dev(hu), *rv = dev(hv), *rq = dev(hu);
    cfd_plan* plan = nullptr;
    cfd_plan_create(&plan, 2, shape, step, 1, 0);
    const float* in[2] = {u, v};
    float* out[2] = {ru, rv};
    cfd_project(plan, nullptr, in, out, rq);
    auto e = B200CfdProject2DImpl(nullptr, Buf(u, dims), Buf(v, dims), Res(Buf(uo, dims)), Res(Buf(vo, dims)),
                                  Res(Buf(q, dims)), stepv, 0);
    EXPECT(e.success(), "projection handler failed");
    cfd_device_sync();
    EXPECT(host(uo, n) == host(ru, n) && host(vo, n) == host(rv, n) && host(q, n) == host(rq, n),
           "projection handler differs from cfd_project");
    cfd_plan_destroy(plan);
  }
  printf(fails ? "XLA SHIM GPU FAIL\n" : "XLA SHIM GPU PASS\n");
  return fails;
}

int main(int argc, char** argv) {
  const bool gpu = argc > 1 && strcmp(argv[1], "gpu") == 0;
  int rc = run_cpu();
  if (gpu) rc += run_gpu();
  return rc ? 1 : 0;
}
