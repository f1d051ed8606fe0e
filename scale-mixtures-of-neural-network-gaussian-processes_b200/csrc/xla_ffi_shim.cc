// XLA FFI custom-call shim over the C-ABI of include/smnngp.h: what the reference's
// JAX code binds with jax.ffi.register_ffi_target(name, capsule, platform="CUDA") so that SPR.loss / SPR.test_nll stay
// inside objax.Jit / jax.jit tracing (spax/models.py:93-120, experiments/regression/train.py:61-67).
//
// NOT COMPILED IN THIS IMAGE: jaxlib (xla/ffi/api/ffi.h) is absent, so this translation unit is excluded from the
// library build (_lib.py SOURCES) and has never been built or run here.  It is kept in the tree as the concrete
// binding a maintainer adds once jaxlib headers are available:
//     g++ -shared -fPIC -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -I../../include \
//         xla_ffi_shim.cc -L../lib -lsmnngp -o libsmnngp_xla.so
// Design rules it follows: the six trainable scalars are an f64[6] OPERAND (they are traced
// values, spax/kernels.py:19-21), only num_hiddens / act / arch / kind are static attributes; every buffer is owned
// by XLA, the scratch is an extra result so the XLA allocator accounts for it; handlers only enqueue on the stream
// XLA hands over; a non-PD matrix is not an error (NaN outputs + info, like lax.linalg.cholesky).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define SMNNGP_HAVE_XLA_FFI 1
#endif
#endif

#ifdef SMNNGP_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include "smnngp.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Status(int rc) {
  return rc == SMNNGP_OK ? ffi::Error::Success() : ffi::Error::Internal(smnngp_last_error());
}

// out f64[4] = {log p, loss, sum log L_ii, quad}; info s32[1]; workspace u8[smnngp_lml_workspace_bytes(...)]
static ffi::Error LmlImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> y,
                          ffi::Buffer<ffi::F64> hp, int32_t num_hiddens, int32_t act, int32_t arch, int32_t kind,
                          ffi::ResultBuffer<ffi::F64> out, ffi::ResultBuffer<ffi::S32> info,
                          ffi::ResultBuffer<ffi::U8> workspace) {
  auto d = x.dimensions();
  return Status(smnngp_lml_f64(stream, x.typed_data(), y.typed_data(), d[0], d[1], num_hiddens, act, arch,
                               hp.typed_data(), kind, workspace->typed_data(), workspace->size_bytes(),
                               out->typed_data(), info->typed_data()));
}

// value + gradient: the fwd rule of a jax.custom_vjp around the loss; grad f64[6] = d loss / d hp (the bwd rule is
// cotangent * grad, and JAX differentiates through softplus itself because hp is built from safe_value in Python)
static ffi::Error LmlGradImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> y,
                              ffi::Buffer<ffi::F64> hp, int32_t num_hiddens, int32_t act, int32_t arch, int32_t kind,
                              ffi::ResultBuffer<ffi::F64> out, ffi::ResultBuffer<ffi::F64> grad,
                              ffi::ResultBuffer<ffi::S32> info, ffi::ResultBuffer<ffi::U8> workspace) {
  auto d = x.dimensions();
  return Status(smnngp_lml_grad_f64(stream, x.typed_data(), y.typed_data(), d[0], d[1], num_hiddens, act, arch,
                                    hp.typed_data(), kind, workspace->typed_data(), workspace->size_bytes(),
                                    out->typed_data(), grad->typed_data(), info->typed_data()));
}

// NNGPKernel.predict (spax/kernels.py:29-32): mean f64[T, C], var f64[T] (= diag cov), relative regulariser
static ffi::Error PredictImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> y,
                              ffi::Buffer<ffi::F64> x_test, ffi::Buffer<ffi::F64> hp, int32_t num_hiddens, int32_t act,
                              int32_t arch, ffi::ResultBuffer<ffi::F64> mean, ffi::ResultBuffer<ffi::F64> var,
                              ffi::ResultBuffer<ffi::S32> info, ffi::ResultBuffer<ffi::U8> workspace) {
  auto d = x.dimensions();
  auto dy = y.dimensions();
  auto dt = x_test.dimensions();
  const int64_t C = dy.size() > 1 ? dy[1] : 1;
  return Status(smnngp_predict_f64(stream, x.typed_data(), y.typed_data(), x_test.typed_data(), d[0], dt[0], C, d[1],
                                   num_hiddens, act, arch, hp.typed_data(), SMNNGP_SHIFT_EPS_REL,
                                   workspace->typed_data(), workspace->size_bytes(), mean->typed_data(),
                                   var->typed_data(), info->typed_data()));
}

// SPR.test_nll (spax/models.py:100-120): nll f64[1]
static ffi::Error TestNllImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> y,
                              ffi::Buffer<ffi::F64> x_test, ffi::Buffer<ffi::F64> y_test, ffi::Buffer<ffi::F64> hp,
                              int32_t num_hiddens, int32_t act, int32_t arch, int32_t kind, double y_mean, double y_std,
                              ffi::ResultBuffer<ffi::F64> nll, ffi::ResultBuffer<ffi::S32> info,
                              ffi::ResultBuffer<ffi::U8> workspace) {
  auto d = x.dimensions();
  auto dt = x_test.dimensions();
  return Status(smnngp_test_nll_f64(stream, x.typed_data(), y.typed_data(), x_test.typed_data(), y_test.typed_data(),
                                    d[0], dt[0], d[1], num_hiddens, act, arch, hp.typed_data(), kind, y_mean, y_std,
                                    workspace->typed_data(), workspace->size_bytes(), nll->typed_data(), nullptr,
                                    nullptr, nullptr, info->typed_data()));
}

#define SMNNGP_STACK_ATTRS() .Attr<int32_t>("num_hiddens").Attr<int32_t>("act").Attr<int32_t>("arch")

XLA_FFI_DEFINE_HANDLER_SYMBOL(SmnngpLml, LmlImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>() SMNNGP_STACK_ATTRS()
                                  .Attr<int32_t>("kind")
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(SmnngpLmlGrad, LmlGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>() SMNNGP_STACK_ATTRS()
                                  .Attr<int32_t>("kind")
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(SmnngpPredict, PredictImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>() SMNNGP_STACK_ATTRS()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(SmnngpTestNll, TestNllImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>() SMNNGP_STACK_ATTRS()
                                  .Attr<int32_t>("kind")
                                  .Attr<double>("y_mean")
                                  .Attr<double>("y_std")
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
#endif  // SMNNGP_HAVE_XLA_FFI
