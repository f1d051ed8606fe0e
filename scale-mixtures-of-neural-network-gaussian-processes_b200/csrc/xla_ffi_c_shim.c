/* XLA FFI custom-call handlers over the C-ABI of include/smnngp.h, written against XLA's plain-C call-frame ABI
 * (SURVEY section 8b / 8f row N4).  What the reference's JAX code registers with
 *     jax.ffi.register_ffi_target("smnngp_lml", jax.ffi.pycapsule(lib.SmnngpLml), platform="CUDA")
 * so that SPR.loss / SPR.test_nll stay inside objax.Jit / jax.jit tracing (spax/models.py:93-120,
 * experiments/regression/train.py:61-67).
 *
 * Two builds of the same source:
 *   default                      : against xla_ffi_min/c_api_subset.h (a local re-declaration of the subset of
 *                                  xla/ffi/api/c_api.h the handlers touch; jaxlib is absent from this image) - this is
 *                                  what libsmnngp_xla.so in this repo is, exercised with a mock call frame by
 *                                  tests/test_xla_ffi_mock.py.  NOT verified against a real XLA runtime here.
 *   -DSMNNGP_USE_REAL_XLA_FFI    : against the real header, once jaxlib is installed.
 *
 * Conventions: operands x f64[N, D], y f64[N] (...), hp f64[6] - the six trainable scalars are a traced
 * OPERAND (spax/kernels.py:19-21), only num_hiddens / act / arch / kind (s32) and y_mean / y_std (f64) are attributes;
 * every buffer belongs to XLA, the scratch is an extra result (u8[workspace_bytes]) so XLA's allocator accounts for it;
 * handlers only enqueue on the stream XLA hands over; a non-PD matrix is not an error (NaN outputs + info).
 */
#ifdef SMNNGP_USE_REAL_XLA_FFI
#include "xla/ffi/api/c_api.h"
#else
#include "xla_ffi_min/c_api_subset.h"
#endif

#include <string.h>

#include "../../include/smnngp.h"

static XLA_FFI_Error* make_error(const XLA_FFI_Api* api, XLA_FFI_Error_Code code, const char* msg) {
  XLA_FFI_Error_Create_Args a;
  memset(&a, 0, sizeof a);
  a.struct_size = sizeof a;
  a.message = msg;
  a.errc = code;
  return api->XLA_FFI_Error_Create(&a);
}

/* registration handshake: XLA calls the handler with a metadata extension and expects the API version back */
static int answer_metadata(XLA_FFI_CallFrame* f) {
  XLA_FFI_Extension_Base* e = f->extension_start;
  for (; e != NULL; e = e->next) {
    if (e->type == XLA_FFI_Extension_Metadata) {
      XLA_FFI_Metadata_Extension* m = (XLA_FFI_Metadata_Extension*)e;
      m->metadata->api_version.major_version = XLA_FFI_API_MAJOR;
      m->metadata->api_version.minor_version = XLA_FFI_API_MINOR;
      return 1;
    }
  }
  return 0;
}

static const XLA_FFI_Buffer* arg_buf(const XLA_FFI_CallFrame* f, int64_t i, XLA_FFI_DataType dt) {
  if (i >= f->args.size || f->args.types[i] != XLA_FFI_ArgType_BUFFER) return NULL;
  const XLA_FFI_Buffer* b = (const XLA_FFI_Buffer*)f->args.args[i];
  return b->dtype == dt ? b : NULL;
}
static const XLA_FFI_Buffer* ret_buf(const XLA_FFI_CallFrame* f, int64_t i, XLA_FFI_DataType dt) {
  if (i >= f->rets.size || f->rets.types[i] != XLA_FFI_RetType_BUFFER) return NULL;
  const XLA_FFI_Buffer* b = (const XLA_FFI_Buffer*)f->rets.rets[i];
  return b->dtype == dt ? b : NULL;
}
static const XLA_FFI_Scalar* attr_scalar(const XLA_FFI_CallFrame* f, const char* name, XLA_FFI_DataType dt) {
  const size_t len = strlen(name);
  for (int64_t i = 0; i < f->attrs.size; i++) {
    const XLA_FFI_ByteSpan* n = f->attrs.names[i];
    if (n->len == len && memcmp(n->ptr, name, len) == 0) {
      if (f->attrs.types[i] != XLA_FFI_AttrType_SCALAR) return NULL;
      const XLA_FFI_Scalar* s = (const XLA_FFI_Scalar*)f->attrs.attrs[i];
      return s->dtype == dt ? s : NULL;
    }
  }
  return NULL;
}
static int attr_i32(const XLA_FFI_CallFrame* f, const char* name, int* out) {
  const XLA_FFI_Scalar* s = attr_scalar(f, name, XLA_FFI_DataType_S32);
  if (!s) return 0;
  *out = *(const int32_t*)s->value;
  return 1;
}
static int attr_f64(const XLA_FFI_CallFrame* f, const char* name, double* out) {
  const XLA_FFI_Scalar* s = attr_scalar(f, name, XLA_FFI_DataType_F64);
  if (!s) return 0;
  *out = *(const double*)s->value;
  return 1;
}
static size_t buf_bytes(const XLA_FFI_Buffer* b, size_t elem) {
  size_t n = elem;
  for (int64_t i = 0; i < b->rank; i++) n *= (size_t)b->dims[i];
  return n;
}
static XLA_FFI_Error* get_stream(XLA_FFI_CallFrame* f, void** stream) {
  XLA_FFI_Stream_Get_Args a;
  memset(&a, 0, sizeof a);
  a.struct_size = sizeof a;
  a.ctx = f->ctx;
  XLA_FFI_Error* e = f->api->XLA_FFI_Stream_Get(&a);
  *stream = a.stream;
  return e;
}

#define PROLOGUE()                                                           \
  if (answer_metadata(f)) return NULL;                                       \
  if (f->stage != XLA_FFI_ExecutionStage_EXECUTE) return NULL;               \
  void* stream = NULL;                                                       \
  {                                                                          \
    XLA_FFI_Error* e__ = get_stream(f, &stream);                             \
    if (e__) return e__;                                                     \
  }                                                                          \
  int num_hiddens, act, arch;                                                \
  if (!attr_i32(f, "num_hiddens", &num_hiddens) || !attr_i32(f, "act", &act) || !attr_i32(f, "arch", &arch)) \
    return make_error(f->api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "smnngp: attributes num_hiddens / act / arch (s32) required")

#define BAD_OPERANDS(what) make_error(f->api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "smnngp " what ": operand / result types or ranks")
#define STATUS(rc) ((rc) == SMNNGP_OK ? NULL : make_error(f->api, XLA_FFI_Error_Code_INTERNAL, smnngp_last_error()))

/* (x f64[N,D], y f64[N], hp f64[6]) -> (out f64[4] = {log p, loss, sum log L_ii, quad}, info s32[1], workspace u8[*]) */
XLA_FFI_Error* SmnngpLml(XLA_FFI_CallFrame* f) {
  PROLOGUE();
  int kind;
  if (!attr_i32(f, "kind", &kind)) return make_error(f->api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "smnngp_lml: attribute kind");
  const XLA_FFI_Buffer *x = arg_buf(f, 0, XLA_FFI_DataType_F64), *y = arg_buf(f, 1, XLA_FFI_DataType_F64),
                       *hp = arg_buf(f, 2, XLA_FFI_DataType_F64);
  const XLA_FFI_Buffer *out = ret_buf(f, 0, XLA_FFI_DataType_F64), *info = ret_buf(f, 1, XLA_FFI_DataType_S32),
                       *ws = ret_buf(f, 2, XLA_FFI_DataType_U8);
  if (!x || !y || !hp || !out || !info || !ws || x->rank != 2) return BAD_OPERANDS("lml");
  return STATUS(smnngp_lml_f64(stream, (const double*)x->data, (const double*)y->data, x->dims[0], x->dims[1],
                               num_hiddens, act, arch, (const double*)hp->data, kind, ws->data, buf_bytes(ws, 1),
                               (double*)out->data, (int*)info->data));
}

/* value + gradient (fwd rule of a jax.custom_vjp): adds grad f64[6] = d loss / d hp before info */
XLA_FFI_Error* SmnngpLmlGrad(XLA_FFI_CallFrame* f) {
  PROLOGUE();
  int kind;
  if (!attr_i32(f, "kind", &kind)) return make_error(f->api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "smnngp_lml_grad: attribute kind");
  const XLA_FFI_Buffer *x = arg_buf(f, 0, XLA_FFI_DataType_F64), *y = arg_buf(f, 1, XLA_FFI_DataType_F64),
                       *hp = arg_buf(f, 2, XLA_FFI_DataType_F64);
  const XLA_FFI_Buffer *out = ret_buf(f, 0, XLA_FFI_DataType_F64), *grad = ret_buf(f, 1, XLA_FFI_DataType_F64),
                       *info = ret_buf(f, 2, XLA_FFI_DataType_S32), *ws = ret_buf(f, 3, XLA_FFI_DataType_U8);
  if (!x || !y || !hp || !out || !grad || !info || !ws || x->rank != 2) return BAD_OPERANDS("lml_grad");
  return STATUS(smnngp_lml_grad_f64(stream, (const double*)x->data, (const double*)y->data, x->dims[0], x->dims[1],
                                    num_hiddens, act, arch, (const double*)hp->data, kind, ws->data, buf_bytes(ws, 1),
                                    (double*)out->data, (double*)grad->data, (int*)info->data));
}

/* NNGPKernel.predict: (x, y f64[N] or f64[N,C], x_test f64[T,D], hp) -> (mean f64[T,C], var f64[T], info, workspace) */
XLA_FFI_Error* SmnngpPredict(XLA_FFI_CallFrame* f) {
  PROLOGUE();
  const XLA_FFI_Buffer *x = arg_buf(f, 0, XLA_FFI_DataType_F64), *y = arg_buf(f, 1, XLA_FFI_DataType_F64),
                       *xt = arg_buf(f, 2, XLA_FFI_DataType_F64), *hp = arg_buf(f, 3, XLA_FFI_DataType_F64);
  const XLA_FFI_Buffer *mean = ret_buf(f, 0, XLA_FFI_DataType_F64), *var = ret_buf(f, 1, XLA_FFI_DataType_F64),
                       *info = ret_buf(f, 2, XLA_FFI_DataType_S32), *ws = ret_buf(f, 3, XLA_FFI_DataType_U8);
  if (!x || !y || !xt || !hp || !mean || !var || !info || !ws || x->rank != 2 || xt->rank != 2) return BAD_OPERANDS("predict");
  const int64_t C = y->rank > 1 ? y->dims[1] : 1;
  return STATUS(smnngp_predict_f64(stream, (const double*)x->data, (const double*)y->data, (const double*)xt->data,
                                   x->dims[0], xt->dims[0], C, x->dims[1], num_hiddens, act, arch,
                                   (const double*)hp->data, SMNNGP_SHIFT_EPS_REL, ws->data, buf_bytes(ws, 1),
                                   (double*)mean->data, (double*)var->data, (int*)info->data));
}

/* SPR.test_nll: (x, y, x_test, y_test, hp) + attrs kind, y_mean, y_std -> (nll f64[1], info, workspace) */
XLA_FFI_Error* SmnngpTestNll(XLA_FFI_CallFrame* f) {
  PROLOGUE();
  int kind;
  double y_mean, y_std;
  if (!attr_i32(f, "kind", &kind) || !attr_f64(f, "y_mean", &y_mean) || !attr_f64(f, "y_std", &y_std))
    return make_error(f->api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "smnngp_test_nll: attributes kind (s32), y_mean, y_std (f64)");
  const XLA_FFI_Buffer *x = arg_buf(f, 0, XLA_FFI_DataType_F64), *y = arg_buf(f, 1, XLA_FFI_DataType_F64),
                       *xt = arg_buf(f, 2, XLA_FFI_DataType_F64), *yt = arg_buf(f, 3, XLA_FFI_DataType_F64),
                       *hp = arg_buf(f, 4, XLA_FFI_DataType_F64);
  const XLA_FFI_Buffer *nll = ret_buf(f, 0, XLA_FFI_DataType_F64), *info = ret_buf(f, 1, XLA_FFI_DataType_S32),
                       *ws = ret_buf(f, 2, XLA_FFI_DataType_U8);
  if (!x || !y || !xt || !yt || !hp || !nll || !info || !ws || x->rank != 2 || xt->rank != 2) return BAD_OPERANDS("test_nll");
  return STATUS(smnngp_test_nll_f64(stream, (const double*)x->data, (const double*)y->data, (const double*)xt->data,
                                    (const double*)yt->data, x->dims[0], xt->dims[0], x->dims[1], num_hiddens, act, arch,
                                    (const double*)hp->data, kind, y_mean, y_std, ws->data, buf_bytes(ws, 1),
                                    (double*)nll->data, NULL, NULL, NULL, (int*)info->data));
}
