"""The XLA FFI handlers (csrc/xla_ffi_c_shim.c -> lib/libsmnngp_xla.so) driven by a MOCK XLA runtime: the call-frame
structs are rebuilt with ctypes, the two API callbacks the handlers use (error creation, stream getter) are Python
functions.  jaxlib is absent from this image, so this is the only way to execute the handler code here; what it proves
is the decode-and-dispatch logic, not ABI compatibility with a real XLA (see the header of csrc/xla_ffi_min/)."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

F64, S32, U8 = 12, 4, 6            # XLA_FFI_DataType


class ExtBase(C.Structure):
    pass


ExtBase._fields_ = [("struct_size", C.c_size_t), ("type", C.c_int), ("next", C.POINTER(ExtBase))]


class ApiVersion(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("major_version", C.c_int),
                ("minor_version", C.c_int)]


class Metadata(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("api_version", ApiVersion), ("traits", C.c_uint32)]


class MetadataExt(C.Structure):
    _fields_ = [("extension_base", ExtBase), ("metadata", C.POINTER(Metadata))]


class ErrorCreateArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("message", C.c_char_p), ("errc", C.c_int)]


class StreamGetArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("ctx", C.c_void_p), ("stream", C.c_void_p)]


ERR_CREATE = C.CFUNCTYPE(C.c_void_p, C.POINTER(ErrorCreateArgs))
STREAM_GET = C.CFUNCTYPE(C.c_void_p, C.POINTER(StreamGetArgs))


class Api(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("api_version", ApiVersion),
                ("internal_api", C.c_void_p), ("XLA_FFI_Error_Create", ERR_CREATE), ("XLA_FFI_Error_GetMessage", C.c_void_p),
                ("XLA_FFI_Error_Destroy", C.c_void_p), ("XLA_FFI_Handler_Register", C.c_void_p),
                ("XLA_FFI_Stream_Get", STREAM_GET)]


class Buffer(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("dtype", C.c_int), ("data", C.c_void_p),
                ("dims", C.POINTER(C.c_int64)), ("rank", C.c_int64)]


class ByteSpan(C.Structure):
    _fields_ = [("ptr", C.c_char_p), ("len", C.c_size_t)]


class Scalar(C.Structure):
    _fields_ = [("dtype", C.c_int), ("value", C.c_void_p)]


class Args(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("size", C.c_int64),
                ("types", C.POINTER(C.c_int)), ("args", C.POINTER(C.c_void_p))]


class Attrs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("size", C.c_int64),
                ("types", C.POINTER(C.c_int)), ("names", C.POINTER(C.POINTER(ByteSpan))), ("attrs", C.POINTER(C.c_void_p))]


class CallFrame(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.POINTER(ExtBase)), ("api", C.POINTER(Api)),
                ("ctx", C.c_void_p), ("stage", C.c_int), ("args", Args), ("rets", Args), ("attrs", Attrs),
                ("future", C.c_void_p)]


class MockXla:
    """owns every ctypes object a call frame points to"""

    def __init__(self, stream=0):
        self.errors, self.keep, self.stream = [], [], stream

        def err_create(a):
            self.errors.append((a.contents.errc, a.contents.message.decode()))
            return 0xdead0000 + len(self.errors)

        def stream_get(a):
            a.contents.stream = self.stream
            return None

        self._ec, self._sg = ERR_CREATE(err_create), STREAM_GET(stream_get)
        self.api = Api(struct_size=C.sizeof(Api), XLA_FFI_Error_Create=self._ec, XLA_FFI_Stream_Get=self._sg)

    def buffers(self, specs):
        """specs: [(dtype, device_ptr, dims)] -> Args"""
        n = len(specs)
        types, ptrs = (C.c_int * n)(*[1] * n), (C.c_void_p * n)()
        for i, (dt, ptr, dims) in enumerate(specs):
            d = (C.c_int64 * max(len(dims), 1))(*dims)
            b = Buffer(struct_size=C.sizeof(Buffer), dtype=dt, data=ptr, dims=d, rank=len(dims))
            self.keep += [d, b]
            ptrs[i] = C.addressof(b)
        self.keep += [types, ptrs]
        return Args(struct_size=C.sizeof(Args), size=n, types=types, args=ptrs)

    def attrs(self, kv):
        items = sorted(kv.items())
        n = len(items)
        types, names, vals = (C.c_int * n)(*[3] * n), (C.POINTER(ByteSpan) * n)(), (C.c_void_p * n)()
        for i, (k, v) in enumerate(items):
            kb = k.encode()
            span = ByteSpan(ptr=kb, len=len(kb))
            val = C.c_double(v) if isinstance(v, float) else C.c_int32(v)
            sc = Scalar(dtype=F64 if isinstance(v, float) else S32, value=C.addressof(val))
            self.keep += [kb, span, val, sc]
            names[i] = C.pointer(span)
            vals[i] = C.addressof(sc)
        self.keep += [types, names, vals]
        return Attrs(struct_size=C.sizeof(Attrs), size=n, types=types, names=names, attrs=vals)

    def frame(self, args, rets, attrs, stage=3, ext=None):
        return CallFrame(struct_size=C.sizeof(CallFrame), extension_start=ext, api=C.pointer(self.api), ctx=None,
                         stage=stage, args=args, rets=rets, attrs=attrs, future=None)


def _lib():
    import smnngp_b200 as sm
    path = sm._lib.build_xla_shim()
    lib = C.CDLL(path)
    for name in ("SmnngpLml", "SmnngpLmlGrad", "SmnngpPredict", "SmnngpTestNll"):
        fn = getattr(lib, name)
        fn.restype = C.c_void_p
        fn.argtypes = [C.POINTER(CallFrame)]
    return lib


def test_handlers_build_answer_metadata_and_reject_bad_frames():
    lib = _lib()
    xla = MockXla()
    # (1) registration handshake: a metadata extension gets the API version back, nothing else happens
    md = Metadata(struct_size=C.sizeof(Metadata))
    ext = MetadataExt(extension_base=ExtBase(struct_size=C.sizeof(MetadataExt), type=1, next=None), metadata=C.pointer(md))
    empty = xla.buffers([])
    f = xla.frame(empty, empty, xla.attrs({}), ext=C.cast(C.pointer(ext), C.POINTER(ExtBase)))
    for name in ("SmnngpLml", "SmnngpLmlGrad", "SmnngpPredict", "SmnngpTestNll"):
        md.api_version.major_version = md.api_version.minor_version = -1
        assert getattr(lib, name)(C.byref(f)) is None
        assert (md.api_version.major_version, md.api_version.minor_version) == (0, 1)
    # (2) non-EXECUTE stages are no-ops
    assert lib.SmnngpLml(C.byref(xla.frame(empty, empty, xla.attrs({}), stage=1))) is None
    # (3) missing attributes / wrong operand types -> XLA_FFI_Error through the API callback, no launch
    assert lib.SmnngpLml(C.byref(xla.frame(empty, empty, xla.attrs({"act": 0})))) is not None
    assert xla.errors[-1][0] == 3 and "num_hiddens" in xla.errors[-1][1]
    f = xla.frame(xla.buffers([(S32, 0, [4, 2])]), empty, xla.attrs({"num_hiddens": 3, "act": 0, "arch": 0, "kind": 1}))
    assert lib.SmnngpLml(C.byref(f)) is not None and xla.errors[-1][0] == 3
    f = xla.frame(empty, empty, xla.attrs({"num_hiddens": 3, "act": 0, "arch": 0, "kind": 1}))
    assert lib.SmnngpTestNll(C.byref(f)) is not None and "y_mean" in xla.errors[-1][1]


@pytest.mark.gpu
def test_handlers_compute_through_a_mock_call_frame():
    import torch
    import smnngp_b200 as sm
    from oracle import nngp_oracle as orc
    from tests.synth import regression_data, DEFAULT_HP as hp
    lib = _lib()
    n, d, t = 1500, 8, 200
    x, y, xt, yt, ym, ys = regression_data(n, d, t=t)
    xd, yd, xtd, ytd = (torch.from_numpy(v).cuda() for v in (x, y, xt, yt))
    hpd = sm.make_hp(**hp)
    kw = dict(num_hiddens=3, act="relu", arch="mlp", w_std=hp["w_std"], b_std=hp["b_std"], last_w_std=hp["last_w_std"])
    s = torch.cuda.Stream()
    xla = MockXla(stream=s.cuda_stream)
    base = lib_ws = sm._lib.load()
    attrs = {"num_hiddens": 3, "act": 0, "arch": 0, "kind": 1}
    out, grad = torch.zeros(4, dtype=torch.float64, device="cuda"), torch.zeros(6, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = torch.empty(base.smnngp_lml_grad_workspace_bytes(n, d, 3, 0), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ins = xla.buffers([(F64, xd.data_ptr(), [n, d]), (F64, yd.data_ptr(), [n]), (F64, hpd.data_ptr(), [6])])
    rets = xla.buffers([(F64, out.data_ptr(), [4]), (S32, info.data_ptr(), [1]), (U8, ws.data_ptr(), [ws.numel()])])
    assert lib.SmnngpLml(C.byref(xla.frame(ins, rets, xla.attrs(attrs)))) is None, xla.errors
    s.synchronize()
    ref = orc.spr_loss(x, y, eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"], **kw)
    assert int(info.item()) == 0 and abs(out[1].item() - ref) <= 1e-8 * abs(ref)
    rets = xla.buffers([(F64, out.data_ptr(), [4]), (F64, grad.data_ptr(), [6]), (S32, info.data_ptr(), [1]),
                        (U8, ws.data_ptr(), [ws.numel()])])
    assert lib.SmnngpLmlGrad(C.byref(xla.frame(ins, rets, xla.attrs(attrs)))) is None, xla.errors
    s.synchronize()
    _, gref = orc.spr_loss_grad(x, y, eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"], **kw)
    g = grad.cpu().numpy()
    assert np.all(np.abs(g - gref) <= 1e-6 * np.abs(gref) + 1e-11 * np.abs(gref).max())
    nll = torch.zeros(1, dtype=torch.float64, device="cuda")
    ws2 = torch.empty(base.smnngp_predict_workspace_bytes(n, t, 1, d, 3, 0), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ins = xla.buffers([(F64, xd.data_ptr(), [n, d]), (F64, yd.data_ptr(), [n]), (F64, xtd.data_ptr(), [t, d]),
                       (F64, ytd.data_ptr(), [t]), (F64, hpd.data_ptr(), [6])])
    rets = xla.buffers([(F64, nll.data_ptr(), [1]), (S32, info.data_ptr(), [1]), (U8, ws2.data_ptr(), [ws2.numel()])])
    a2 = dict(attrs, y_mean=float(ym), y_std=float(ys))
    assert lib.SmnngpTestNll(C.byref(xla.frame(ins, rets, xla.attrs(a2)))) is None, xla.errors
    s.synchronize()
    ref_nll = orc.spr_test_nll(x, y, xt, yt, ym, ys, eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"], **kw)
    assert abs(float(nll.item()) - ref_nll) <= 1e-8 * abs(ref_nll)
    mean, var = torch.zeros((t, 1), dtype=torch.float64, device="cuda"), torch.zeros(t, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ins = xla.buffers([(F64, xd.data_ptr(), [n, d]), (F64, yd.data_ptr(), [n]), (F64, xtd.data_ptr(), [t, d]),
                       (F64, hpd.data_ptr(), [6])])
    rets = xla.buffers([(F64, mean.data_ptr(), [t, 1]), (F64, var.data_ptr(), [t]), (S32, info.data_ptr(), [1]),
                        (U8, ws2.data_ptr(), [ws2.numel()])])
    assert lib.SmnngpPredict(C.byref(xla.frame(ins, rets, xla.attrs({"num_hiddens": 3, "act": 0, "arch": 0})))) is None
    s.synchronize()
    mean_ref, cov_ref = orc.nt_predict(x, y[:, None], xt, hp["eps"], kernel_kwargs=kw)
    assert np.abs(mean.cpu().numpy() - mean_ref).max() <= 1e-8 * np.abs(mean_ref).max()
