/* Locally re-declared SUBSET of XLA's FFI C API (xla/ffi/api/c_api.h, API version 0.x) - only what a handler needs to
 * decode a call frame: buffers, scalar attributes, the stream getter and error creation.
 *
 * Why it exists: jaxlib (which ships the real header) is absent from this image and there is no network, so the
 * C++ binding-layer shim (../xla_ffi_shim.cc, needs xla/ffi/api/ffi.h) cannot be compiled here.  The C-level handlers
 * in ../xla_ffi_c_shim.c only touch the plain-C call-frame structs, which are a versioned, append-only ABI; declaring
 * that subset here lets the handlers be compiled and exercised against a mock call frame (tests/test_xla_ffi_mock.py).
 *
 * STATUS: written from the public header's layout, NOT verified against a real jaxlib in this environment.  When
 * jaxlib is available, build with -DSMNNGP_USE_REAL_XLA_FFI -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())")
 * and this file is not used at all (the shim then includes "xla/ffi/api/c_api.h").
 */
#ifndef SMNNGP_XLA_FFI_C_API_SUBSET_H_
#define SMNNGP_XLA_FFI_C_API_SUBSET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct XLA_FFI_Api XLA_FFI_Api;
typedef struct XLA_FFI_InternalApi XLA_FFI_InternalApi;
typedef struct XLA_FFI_Error XLA_FFI_Error;
typedef struct XLA_FFI_Future XLA_FFI_Future;
typedef struct XLA_FFI_ExecutionContext XLA_FFI_ExecutionContext;

typedef enum { XLA_FFI_Extension_Metadata = 1 } XLA_FFI_Extension_Type;
typedef struct XLA_FFI_Extension_Base {
  size_t struct_size;
  XLA_FFI_Extension_Type type;
  struct XLA_FFI_Extension_Base* next;
} XLA_FFI_Extension_Base;

typedef struct XLA_FFI_Api_Version {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int major_version; /* out */
  int minor_version; /* out */
} XLA_FFI_Api_Version;

typedef enum {
  XLA_FFI_Error_Code_OK = 0,
  XLA_FFI_Error_Code_CANCELLED = 1,
  XLA_FFI_Error_Code_UNKNOWN = 2,
  XLA_FFI_Error_Code_INVALID_ARGUMENT = 3,
  XLA_FFI_Error_Code_INTERNAL = 13
} XLA_FFI_Error_Code;

typedef struct XLA_FFI_Error_Create_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  const char* message;
  XLA_FFI_Error_Code errc;
} XLA_FFI_Error_Create_Args;
typedef XLA_FFI_Error* XLA_FFI_Error_Create(XLA_FFI_Error_Create_Args* args);

typedef enum {
  XLA_FFI_DataType_INVALID = 0,
  XLA_FFI_DataType_PRED = 1,
  XLA_FFI_DataType_S8 = 2,
  XLA_FFI_DataType_S16 = 3,
  XLA_FFI_DataType_S32 = 4,
  XLA_FFI_DataType_S64 = 5,
  XLA_FFI_DataType_U8 = 6,
  XLA_FFI_DataType_U16 = 7,
  XLA_FFI_DataType_U32 = 8,
  XLA_FFI_DataType_U64 = 9,
  XLA_FFI_DataType_F16 = 10,
  XLA_FFI_DataType_F32 = 11,
  XLA_FFI_DataType_F64 = 12
} XLA_FFI_DataType;

typedef struct XLA_FFI_Buffer {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_DataType dtype;
  void* data;
  int64_t* dims;
  int64_t rank;
} XLA_FFI_Buffer;

typedef enum { XLA_FFI_ArgType_BUFFER = 1 } XLA_FFI_ArgType;
typedef enum { XLA_FFI_RetType_BUFFER = 1 } XLA_FFI_RetType;
typedef enum {
  XLA_FFI_AttrType_ARRAY = 1,
  XLA_FFI_AttrType_DICTIONARY = 2,
  XLA_FFI_AttrType_SCALAR = 3,
  XLA_FFI_AttrType_STRING = 4
} XLA_FFI_AttrType;

typedef struct XLA_FFI_ByteSpan {
  const char* ptr;
  size_t len;
} XLA_FFI_ByteSpan;

typedef struct XLA_FFI_Scalar {
  XLA_FFI_DataType dtype;
  void* value;
} XLA_FFI_Scalar;

typedef enum {
  XLA_FFI_ExecutionStage_INSTANTIATE = 0,
  XLA_FFI_ExecutionStage_PREPARE = 1,
  XLA_FFI_ExecutionStage_INITIALIZE = 2,
  XLA_FFI_ExecutionStage_EXECUTE = 3
} XLA_FFI_ExecutionStage;

typedef struct XLA_FFI_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_ArgType* types; /* length == size */
  void** args;            /* length == size: XLA_FFI_Buffer* for BUFFER */
} XLA_FFI_Args;

typedef struct XLA_FFI_Rets {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_RetType* types;
  void** rets;
} XLA_FFI_Rets;

typedef struct XLA_FFI_Attrs {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_AttrType* types;  /* length == size */
  XLA_FFI_ByteSpan** names; /* length == size, sorted by name */
  void** attrs;             /* length == size: XLA_FFI_Scalar* for SCALAR */
} XLA_FFI_Attrs;

typedef struct XLA_FFI_CallFrame {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  const XLA_FFI_Api* api;
  XLA_FFI_ExecutionContext* ctx;
  XLA_FFI_ExecutionStage stage;
  XLA_FFI_Args args;
  XLA_FFI_Rets rets;
  XLA_FFI_Attrs attrs;
  XLA_FFI_Future* future; /* out: optional, for asynchronous handlers */
} XLA_FFI_CallFrame;

typedef struct XLA_FFI_Metadata {
  size_t struct_size;
  XLA_FFI_Api_Version api_version;
  uint32_t traits;
} XLA_FFI_Metadata;
typedef struct XLA_FFI_Metadata_Extension {
  XLA_FFI_Extension_Base extension_base;
  XLA_FFI_Metadata* metadata;
} XLA_FFI_Metadata_Extension;

typedef struct XLA_FFI_Stream_Get_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_ExecutionContext* ctx;
  void* stream; /* out */
} XLA_FFI_Stream_Get_Args;
typedef XLA_FFI_Error* XLA_FFI_Stream_Get(XLA_FFI_Stream_Get_Args* args);

/* The leading members of XLA_FFI_Api (the struct is append-only; the handlers use nothing past Stream_Get). */
struct XLA_FFI_Api {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_Api_Version api_version;
  XLA_FFI_InternalApi* internal_api;
  XLA_FFI_Error_Create* XLA_FFI_Error_Create;
  void* XLA_FFI_Error_GetMessage;
  void* XLA_FFI_Error_Destroy;
  void* XLA_FFI_Handler_Register;
  XLA_FFI_Stream_Get* XLA_FFI_Stream_Get;
};

#define XLA_FFI_API_MAJOR 0
#define XLA_FFI_API_MINOR 1

typedef XLA_FFI_Error* XLA_FFI_Handler(XLA_FFI_CallFrame* call_frame);

#ifdef __cplusplus
}
#endif
#endif /* SMNNGP_XLA_FFI_C_API_SUBSET_H_ */
