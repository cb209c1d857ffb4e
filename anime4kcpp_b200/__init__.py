"""anime4kcpp_b200 -- B200 (sm_100a) backend for Anime4KCPP v3's CNN upscaling hot path.

The product is the native library ``lib/libac_b200.so`` (CUDA kernels, the thin C-ABI ``acb200_*``, the
``ac::core`` host classes and the libac_c-compatible ``ac_*`` binding) plus the pybind11 module ``pyac``.
This package is only the Python-side loader: ctypes prototypes for the C-ABI and small helpers that hand
numpy arrays (host path) or torch CUDA tensors (device-resident path) to it.

There is no CPU fallback: every compute call needs the native library and a CUDA device and fails loudly
otherwise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACB200_LIB") or os.path.join(_HERE, "lib", "libac_b200.so")

UINT8, UINT16, FLOAT16, FLOAT32 = 0x001, 0x002, 0x202, 0x204
FAMILY_ACNET_LEGACY, FAMILY_ACNET, FAMILY_ARNET = 0, 1, 2
_NP_TYPES = {np.dtype(np.uint8): UINT8, np.dtype(np.uint16): UINT16, np.dtype(np.float16): FLOAT16, np.dtype(np.float32): FLOAT32}

_vp, _i, _d, _cp = C.c_void_p, C.c_int, C.c_double, C.c_char_p
_fp = C.POINTER(C.c_float)
_lib = None


class Plane(C.Structure):
    """acb200_plane == one entry of ac::video::Frame::plane (video/include/AC/Video/Pipeline.hpp:18-24)."""
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channel", C.c_int), ("stride", C.c_int), ("data", C.c_void_p)]


def _planes_of(arrays):
    """numpy planes (H,W) / (H,W,2) -> ctypes array of acb200_plane (the arrays must stay alive during the call)."""
    out = (Plane * len(arrays))()
    for i, a in enumerate(arrays):
        if a.strides[-1] != a.itemsize or (a.ndim == 3 and a.strides[1] != a.itemsize * a.shape[2]):
            raise ValueError("planes must be contiguous along a row")
        out[i] = Plane(a.shape[1], a.shape[0], 1 if a.ndim == 2 else a.shape[2], a.strides[0], a.ctypes.data)
    return out


def frame_result_planes(planes, factor):
    """Destination planes of one video frame: every plane `factor` x its source (what Pipeline::request allocates)."""
    return [np.empty((int(p.shape[0] * factor), int(p.shape[1] * factor)) + p.shape[2:], p.dtype) for p in planes]


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """The native library (loaded once).  Raises NativeLibraryMissing if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(LIB_PATH + " not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                       "(or `make -C anime4kcpp_b200/csrc`); there is no Python/CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.acb200_device_count.restype = _i
        L.acb200_device_info.argtypes = [_i, _cp, _i, C.POINTER(C.c_size_t), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]
        L.acb200_model_create.argtypes = [_i, _i, _fp, _i, _fp, _i, _fp, _i, C.POINTER(_vp)]
        L.acb200_model_create_wide.argtypes = [_i, _i, _i, _fp, _i, _fp, _i, _fp, _i, C.POINTER(_vp)]
        L.acb200_model_destroy.argtypes = [_vp]
        L.acb200_model_destroy.restype = None
        L.acb200_session_create.argtypes = [_i, C.POINTER(_vp)]
        L.acb200_session_destroy.argtypes = [_vp]
        L.acb200_session_destroy.restype = None
        L.acb200_session_error.argtypes = [_vp]
        L.acb200_session_error.restype = _cp
        L.acb200_session_sync.argtypes = [_vp]
        L.acb200_session_set_engine.argtypes = [_vp, _i]
        L.acb200_session_set_tensor_impl.argtypes = [_vp, _i]
        L.acb200_session_set_fusion.argtypes = [_vp, _i]
        L.acb200_session_last_kernel_ms.argtypes = [_vp]
        L.acb200_session_last_kernel_ms.restype = C.c_float
        L.acb200_process_host.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _i]
        L.acb200_process_device.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _i, _vp]
        L.acb200_rgb2yuv_host.argtypes = [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _i]
        L.acb200_yuv2rgb_host.argtypes = [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _i]
        L.acb200_rgb2yuv_packed_host.argtypes = [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i]
        L.acb200_yuv2rgb_packed_host.argtypes = [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i]
        L.acb200_resize_catmull_rom_host.argtypes = [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i]
        L.acb200_process_frame_host.argtypes = [_vp, _vp, C.POINTER(Plane), C.POINTER(Plane), _i, _i, _i, _d]
        L.acb200_process_frame_device.argtypes = [_vp, _vp, C.POINTER(Plane), C.POINTER(Plane), _i, _i, _i, _d, _vp]
        L.acb200_stream_submit_frame.argtypes = [_vp, C.POINTER(Plane), C.POINTER(Plane), _i, _i, _i, _d, C.POINTER(C.c_longlong)]
        L.acb200_model_halo.argtypes = [_vp]
        L.acb200_band_plan.argtypes = [_i, _d, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]
        L.acb200_process_host_band.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _i, _i, _vp, _i]
        L.acb200_frame_owner.argtypes = [C.c_longlong, _i]
        L.acb200_stream_create.argtypes = [_vp, C.POINTER(_i), _i, _i, _i, C.POINTER(_vp)]
        L.acb200_stream_submit.argtypes = [_vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _i, C.POINTER(C.c_longlong)]
        L.acb200_stream_next.argtypes = [_vp, C.POINTER(C.c_longlong), C.POINTER(_i)]
        L.acb200_stream_destroy.argtypes = [_vp]
        L.acb200_stream_destroy.restype = None
        L.acb200_launch_count.restype = C.c_ulonglong
        L.acb200_error_string.argtypes = [_i]
        L.acb200_error_string.restype = _cp
        L.acb200_version.restype = _cp
        L.ac_b200_resolve_model.argtypes = [_cp]
        L.ac_b200_resolve_model.restype = _cp
        L.ac_b200_model_arrays.argtypes = [_cp, C.POINTER(_i), C.POINTER(_fp), C.POINTER(_i), C.POINTER(_fp), C.POINTER(_i), C.POINTER(_fp), C.POINTER(_i)]
        L.ac_b200_model_arrays.restype = _i
        L.ac_b200_model_features.argtypes = [_cp]
        _lib = L
    return _lib


class Acb200Error(RuntimeError):
    pass


def _check(rc, session=None):
    if rc != 0:
        text = lib().acb200_error_string(rc).decode()
        if session is not None:
            text += ": " + lib().acb200_session_error(session).decode()
        raise Acb200Error(text)


def resolve_model(name):
    """Canonical model name a model string selects (Processor.cpp:26-187 rules); '' if out of scope."""
    return lib().ac_b200_resolve_model(name.encode() if name is not None else None).decode()


def model_arrays(name):
    """(family, blocks, kernels, biases, alphas) exactly as the library hands them to the CUDA layer."""
    blocks, nk, nb, na = _i(), _i(), _i(), _i()
    k, b, a = _fp(), _fp(), _fp()
    fam = lib().ac_b200_model_arrays(name.encode(), blocks, k, nk, b, nb, a, na)
    if fam < 0:
        raise Acb200Error("model family out of scope: %r" % name)
    ka = np.ctypeslib.as_array(k, (nk.value,)).copy()
    ba = np.ctypeslib.as_array(b, (nb.value,)).copy()
    aa = np.ctypeslib.as_array(a, (na.value,)).copy() if na.value else np.zeros(0, np.float32)
    return fam, blocks.value, ka, ba, aa


class Model:
    """acb200_model: flat fp32 arrays of one network variant."""

    def __init__(self, name=None, family=None, blocks=None, kernels=None, biases=None, alphas=None, features=8):
        if name is not None:
            family, blocks, kernels, biases, alphas = model_arrays(name)
            features = lib().ac_b200_model_features(name.encode())
            self.name = resolve_model(name)
        else:
            self.name = "custom"
        self.family, self.blocks, self.features = family, blocks, features
        kernels = np.ascontiguousarray(kernels, np.float32)
        biases = np.ascontiguousarray(biases, np.float32)
        alphas = np.ascontiguousarray(alphas if alphas is not None else np.zeros(0), np.float32)
        h = _vp()
        rc = lib().acb200_model_create_wide(family, features, blocks, kernels.ctypes.data_as(_fp), kernels.size, biases.ctypes.data_as(_fp), biases.size,
                                            alphas.ctypes.data_as(_fp) if alphas.size else C.cast(None, _fp), alphas.size, h)
        _check(rc)
        self.handle = h

    def halo(self):
        return lib().acb200_model_halo(self.handle)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:
            _lib.acb200_model_destroy(h)


class Session:
    """acb200_session: one CUDA stream + scratch on one device; use from one thread at a time."""

    def __init__(self, device=0):
        h = _vp()
        _check(lib().acb200_session_create(device, h))
        self.handle = h
        self.device = device

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:
            _lib.acb200_session_destroy(h)

    def set_engine(self, engine):
        _check(lib().acb200_session_set_engine(self.handle, engine), self.handle)

    def set_tensor_impl(self, impl):
        _check(lib().acb200_session_set_tensor_impl(self.handle, impl), self.handle)

    def set_fusion(self, on):
        """Colour split / chroma resize / merge inside the TMEM engine's segment kernels (8-bit RGB, 2x); bit-identical either way."""
        _check(lib().acb200_session_set_fusion(self.handle, int(on)), self.handle)        # False / True, or 2: RGBA as well

    def sync(self):
        _check(lib().acb200_session_sync(self.handle), self.handle)

    def last_kernel_ms(self):
        return float(lib().acb200_session_last_kernel_ms(self.handle))

    def process_host(self, model, img, factor=2.0, out=None):
        """Processor::process on a host numpy image (H,W) / (H,W,C); rows may be strided."""
        if img.ndim not in (2, 3) or img.strides[-1] != img.itemsize or (img.ndim == 3 and img.strides[1] != img.itemsize * img.shape[2]):
            img = np.ascontiguousarray(img)
        h, w = img.shape[:2]
        c = 1 if img.ndim == 2 else img.shape[2]
        oshape = (int(h * factor), int(w * factor)) + (() if img.ndim == 2 else (c,))
        if out is None:
            out = np.empty(oshape, img.dtype)
        rc = lib().acb200_process_host(self.handle, model.handle, img.ctypes.data, w, h, c, img.strides[0], _NP_TYPES[img.dtype], float(factor),
                                       out.ctypes.data, out.strides[0])
        _check(rc, self.handle)
        return out

    def process_device(self, model, src, factor=2.0, out=None, stream=None):
        """Same on a torch CUDA tensor already in HBM (uint8/int16-as-uint16/float16/float32, (H,W) or (H,W,C), contiguous
        rows).  Asynchronous on `stream` (a raw cudaStream_t int; default: the session's own stream)."""
        import torch
        assert src.is_cuda and src.stride(-1) == 1
        h, w = src.shape[:2]
        c = 1 if src.dim() == 2 else src.shape[2]
        tcode = {torch.uint8: UINT8, torch.int16: UINT16, torch.float16: FLOAT16, torch.float32: FLOAT32}[src.dtype]
        if out is None:
            out = torch.empty((int(h * factor), int(w * factor)) + (() if src.dim() == 2 else (c,)), dtype=src.dtype, device=src.device)
        if stream == 0:
            stream = 1      # cudaStreamLegacy: NULL means "the session's own stream" in the C-ABI
        rc = lib().acb200_process_device(self.handle, model.handle, src.data_ptr(), w, h, c, src.stride(0) * src.element_size(), tcode,
                                         float(factor), out.data_ptr(), out.stride(0) * out.element_size(), stream)
        _check(rc, self.handle)
        return out

    def process_frame_device(self, model, planes, out, factor=2.0, shift=0, stream=None):
        """acb200_process_frame_device on torch CUDA tensors (planes and `out`: lists of (H,W) / (H,W,2) tensors in HBM)."""
        import torch
        tcode = {torch.uint8: UINT8, torch.int16: UINT16, torch.float16: FLOAT16, torch.float32: FLOAT32}[planes[0].dtype]

        def pack(ts):
            arr = (Plane * len(ts))()
            for i, t in enumerate(ts):
                assert t.is_cuda and t.stride(-1) == 1
                arr[i] = Plane(t.shape[1], t.shape[0], 1 if t.dim() == 2 else t.shape[2], t.stride(0) * t.element_size(), t.data_ptr())
            return arr
        if stream == 0:
            stream = 1
        _check(lib().acb200_process_frame_device(self.handle, model.handle, pack(planes), pack(out), len(planes), tcode, int(shift), float(factor), stream),
               self.handle)
        return out

    def process_frame(self, model, planes, factor=2.0, shift=0, out=None):
        """One planar / semi-planar YUV frame (acb200_process_frame_host): planes[0] = luma (H,W) through the network,
        the others = chroma (h,w) or (h,w,2) through the Catmull-Rom resize; returns the list of result planes."""
        if out is None:
            out = frame_result_planes(planes, factor)
        src, dst = _planes_of(planes), _planes_of(out)
        _check(lib().acb200_process_frame_host(self.handle, model.handle, src, dst, len(planes), _NP_TYPES[planes[0].dtype], int(shift), float(factor)),
               self.handle)
        return out

    def rgb2yuv(self, img):
        img = np.ascontiguousarray(img)
        h, w, c = img.shape
        y = np.empty((h, w), img.dtype)
        uv = np.empty((h, w, c - 1), img.dtype)
        _check(lib().acb200_rgb2yuv_host(self.handle, img.ctypes.data, w, h, c, img.strides[0], _NP_TYPES[img.dtype], y.ctypes.data, y.strides[0],
                                         uv.ctypes.data, uv.strides[0]), self.handle)
        return y, uv

    def yuv2rgb(self, y, uv):
        y, uv = np.ascontiguousarray(y), np.ascontiguousarray(uv)
        h, w = y.shape
        c = uv.shape[2] + 1
        out = np.empty((h, w, c), y.dtype)
        _check(lib().acb200_yuv2rgb_host(self.handle, y.ctypes.data, y.strides[0], uv.ctypes.data, uv.strides[0], w, h, c, _NP_TYPES[y.dtype],
                                         out.ctypes.data, out.strides[0]), self.handle)
        return out

    def resize_catmull_rom(self, img, ow, oh):
        img = np.ascontiguousarray(img)
        h, w = img.shape[:2]
        c = 1 if img.ndim == 2 else img.shape[2]
        out = np.empty((oh, ow) + (() if img.ndim == 2 else (c,)), img.dtype)
        _check(lib().acb200_resize_catmull_rom_host(self.handle, img.ctypes.data, w, h, c, img.strides[0], _NP_TYPES[img.dtype], out.ctypes.data,
                                                    ow, oh, out.strides[0]), self.handle)
        return out


def band_plan(h, factor, halo, n_bands, band):
    """(src_y0, src_y1, out_y0, out_y1) of one row band (acb200_band_plan)."""
    a, b, c, d = _i(), _i(), _i(), _i()
    _check(lib().acb200_band_plan(h, float(factor), halo, n_bands, band, a, b, c, d))
    return a.value, b.value, c.value, d.value


def frame_owner(seq, n_devices):
    return lib().acb200_frame_owner(seq, n_devices)


def process_band(session, model, img, factor, n_bands, band, out, out_y0=0):
    """Upscale row band `band` of `n_bands` of a host image on `session`'s GPU, writing only that band's rows of `out`.  `out` is the whole
    result image, or -- with `out_y0` = the band's first output row (band_plan) -- a buffer that holds just the band's rows (the shape a
    multi-GPU job uses: every rank keeps its own band of the result)."""
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    _check(lib().acb200_process_host_band(session.handle, model.handle, img.ctypes.data, w, h, c, img.strides[0], _NP_TYPES[img.dtype], float(factor),
                                          n_bands, band, out.ctypes.data - out_y0 * out.strides[0], out.strides[0]), session.handle)


class FrameStream:
    """acb200_stream: frames dealt round-robin over `devices`, results delivered in submission order."""

    def __init__(self, model, devices, workers_per_device=2, queue_depth=2):
        arr = (_i * len(devices))(*devices)
        h = _vp()
        _check(lib().acb200_stream_create(model.handle, arr, len(devices), workers_per_device, queue_depth, h))
        self.handle, self.model = h, model
        self._keep = {}

    def submit(self, img, factor, out):
        h, w = img.shape[:2]
        c = 1 if img.ndim == 2 else img.shape[2]
        seq = C.c_longlong()
        _check(lib().acb200_stream_submit(self.handle, img.ctypes.data, w, h, c, img.strides[0], _NP_TYPES[img.dtype], float(factor),
                                          out.ctypes.data, out.strides[0], seq))
        self._keep[seq.value] = (img, out)
        return seq.value

    def submit_frame(self, planes, factor, out, shift=0):
        """planar-frame form of submit (acb200_stream_submit_frame)."""
        seq = C.c_longlong()
        _check(lib().acb200_stream_submit_frame(self.handle, _planes_of(planes), _planes_of(out), len(planes), _NP_TYPES[planes[0].dtype], int(shift),
                                                float(factor), seq))
        self._keep[seq.value] = (planes, out)
        return seq.value

    def next(self):
        seq, status = C.c_longlong(), _i()
        _check(lib().acb200_stream_next(self.handle, seq, status))
        _check(status.value)
        return seq.value, self._keep.pop(seq.value)[1]

    def close(self):
        h, self.handle = self.handle, None
        if h:
            lib().acb200_stream_destroy(h)

    def __del__(self):
        if getattr(self, "handle", None) and _lib is not None:
            self.close()


def device_count():
    return lib().acb200_device_count()


def launch_count():
    return int(lib().acb200_launch_count())
