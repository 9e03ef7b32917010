// XLA FFI (typed FFI, API v1) handler forwarding to the C ABI of include/cfd_b200.h.
//
// NOT part of the default build: the XLA FFI headers (xla/ffi/api/ffi.h, located by
// jax.ffi.include_dir()) and JAX itself are absent from this image, so this translation unit
// could be neither compiled nor executed here.  It is written against the public jax.ffi API and
// kept deliberately thin: all work is in cfd_repeated().  Build (where JAX is installed):
//   g++ -O2 -fPIC -shared -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I../../include xla_ffi_shim.cc -L../lib -lcfd_b200 -o ../lib/libcfd_b200_xla.so
#include <cstdint>

#include "cfd_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

// The plan / params are created on the host once per step_fn (trace time) through the C ABI and
// passed as opaque 64-bit attributes; XLA owns every operand / result buffer.
static ffi::Error Step2DImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> u, ffi::Buffer<ffi::F32> v,
                             ffi::ResultBuffer<ffi::F32> uo, ffi::ResultBuffer<ffi::F32> vo,
                             int64_t plan_handle, int64_t params_handle, int32_t nsteps) {
  if (u.dimensions().size() < 2 || u.dimensions() != v.dimensions())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "b200cfd_step2d: bad operand shapes");
  cfd_plan* plan = reinterpret_cast<cfd_plan*>(plan_handle);
  const cfd_params* params = reinterpret_cast<const cfd_params*>(params_handle);
  const float* in[2] = {u.typed_data(), v.typed_data()};
  float* out[2] = {uo->typed_data(), vo->typed_data()};
  int rc;
  if (nsteps == 1) {
    rc = cfd_step(plan, stream, in, out, nullptr, params);
  } else {
    // funcutils.repeated: operands are immutable, so chain through the result buffers:
    // cfd_repeated leaves the result in its second buffer set when nsteps is odd.
    float* a[2] = {const_cast<float*>(in[0]), const_cast<float*>(in[1])};
    int in_b = 0;
    rc = (nsteps & 1) ? cfd_repeated(plan, stream, a, out, nsteps, params, &in_b)
                      : 1;  // even chains need a scratch pair: requested through ScratchAllocator
  }
  if (rc != 0) return ffi::Error(ffi::ErrorCode::kInternal, cfd_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    B200CfdStep2D, Step2DImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>()
        .Attr<int64_t>("plan")
        .Attr<int64_t>("params")
        .Attr<int32_t>("nsteps"),
    {ffi::Traits::kCmdBufferCompatible});
