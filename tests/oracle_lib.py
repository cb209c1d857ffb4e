"""ctypes loaders for the CPU oracle (oracle/libac_oracle.so) and, when it has been built, the
compiled reference (oracle/_ref/libac_ref.so).  Test infrastructure: imported only by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
WEIGHTS = os.path.join(ROOT, "anime4kcpp_b200", "weights", "acnet.bin")

U8, U16, F32 = 0x001, 0x002, 0x204
NP_TYPES = {np.dtype(np.uint8): U8, np.dtype(np.uint16): U16, np.dtype(np.float32): F32}

FAMILY_LEGACY, FAMILY_ACNET, FAMILY_ARNET, FAMILY_ARTCNN, FAMILY_FSRCNNX = 0, 1, 2, 3, 4

_vp, _fp, _i, _d = C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_double


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libac_oracle.so")
    src = os.path.join(ORACLE_DIR, "ac_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libac_oracle.so"], stdout=subprocess.DEVNULL)
    return so


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(build_oracle())
        lib.orc_luma_pass.argtypes = [_i, _i, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _vp, _i]
        lib.orc_luma_pass.restype = _i
        lib.orc_process.argtypes = [_i, _i, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _d, _vp, _i]
        lib.orc_process.restype = _i
        lib.orc_rgb2yuv.argtypes = [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _i]
        lib.orc_yuv2rgb.argtypes = [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _i]
        lib.orc_resize_catmull_rom.argtypes = [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i]
        lib.orc_resize_catmull_rom.restype = _i
        _oracle = lib
    return _oracle


def ref_path():
    return os.path.join(ORACLE_DIR, "_ref", "libac_ref.so")


def ref():
    """The compiled reference, or None when oracle/_ref has not been built."""
    global _ref
    if _ref is None and os.path.exists(ref_path()):
        lib = C.CDLL(ref_path())
        lib.ref_process.argtypes = [C.c_char_p, _i, _vp, _i, _i, _i, _i, _i, _d, _vp, _i]
        lib.ref_process.restype = _i
        lib.ref_processor_name.argtypes = [C.c_char_p, _i]
        lib.ref_processor_name.restype = C.c_char_p
        lib.ref_rgb2yuv.argtypes = [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _i]
        lib.ref_yuv2rgb.argtypes = [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _i]
        lib.ref_model_arrays.argtypes = [C.c_char_p, C.POINTER(_fp), C.POINTER(_i), C.POINTER(_fp), C.POINTER(_i), C.POINTER(_fp), C.POINTER(_i)]
        lib.ref_model_arrays.restype = _i
        lib.ref_benchmark.argtypes = [C.c_char_p, _i, _i, _i, _i, _i, _i, C.c_uint]
        lib.ref_benchmark.restype = _d
        _ref = lib
    return _ref


# ---------------------------------------------------------------------------------------------
# weights: the committed blob (real ACNet numbers) + the seeded ARNet stand-ins
# ---------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _splitmix64(state):
    state[0] = (state[0] + 0x9E3779B97F4A7C15) & _M64
    z = state[0]
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def synth_arnet(name, blocks):
    """Python restatement of anime4kcpp_b200/csrc/synth_weights.h (exact fp32 arithmetic)."""
    h = 0xCBF29CE484222325
    for ch in name.encode():
        h = ((h ^ ch) * 0x100000001B3) & _M64
    st = [h]
    f32 = np.float32
    inv = f32(1.0) / f32(16777216.0)

    def uni():
        return f32(_splitmix64(st) >> 40) * inv

    def normal():
        u = uni()
        u = f32(u + uni())
        u = f32(u + uni())
        u = f32(u + uni())
        return f32(f32(u - f32(2.0)) * f32(1.7320508))

    nk = 72 + 576 * blocks * 2 + 64 + 288
    nb = 8 + 8 * (blocks * 2 + 1) + 4
    na = 8 * (blocks + 1)
    body_end = 72 + 576 * blocks * 2
    k = np.empty(nk, np.float32)
    for i in range(nk):
        sigma = f32(0.2) if i < 72 else (f32(0.05) if i < body_end else (f32(0.2) if i < body_end + 64 else f32(0.05)))
        k[i] = f32(normal() * sigma)
    b = np.array([f32(normal() * f32(0.02)) for _ in range(nb)], np.float32)
    a = np.array([f32(f32(0.05) + f32(f32(0.25) * uni())) for _ in range(na)], np.float32)
    return k, b, a


_models = None


def models():
    """name -> (family, blocks, kernels, biases, alphas) for the real-weight variants (for ArtCNN / FSRCNNX `blocks`
    carries the feature count in its upper 16 bits: blocks | F << 16, as written by tools/gen_weights.cpp)."""
    global _models
    if _models is None:
        raw = open(WEIGHTS, "rb").read()
        assert raw[:8] == b"ACB2WTS1"
        n, _ = struct.unpack_from("<II", raw, 8)
        base = 16 + n * 72
        out = {}
        for m in range(n):
            name, fam, blocks, nk, nb, na, off = struct.unpack_from("<48sIIIIII", raw, 16 + m * 72)
            name = name.split(b"\0")[0].decode()
            data = np.frombuffer(raw, np.float32, nk + nb + na, base + 4 * off)
            out[name] = (fam, blocks & 0xffff, data[:nk].copy(), data[nk:nk + nb].copy(), data[nk + nb:].copy())
            _features[name] = (blocks >> 16) or 8
        _models = out
    return _models


_features = {}


def features(name):
    """F of the model template: 8, except ArtCNN<16/32> and FSRCNNX<16>."""
    models()
    return _features.get(canonical(name), 8)


_arnet_cache = {}


def model(name):
    name = canonical(name)
    if name.startswith("arnet"):
        if name not in _arnet_cache:
            blocks = int(name.split("b")[1].split("-")[0])
            _arnet_cache[name] = (FAMILY_ARNET, blocks) + synth_arnet(name, blocks)
        return _arnet_cache[name]
    return models()[name]


def canonical(s):
    """Model-string parsing of core/src/processor/Processor.cpp:26-187 (ACNet/ARNet branches)."""
    s = (s or "").lower()
    if "fsrcnnx" in s:
        return "fsrcnnx-f%sb4" % ("16" if "f16" in s else "8") + ("-distort-plus" if ("distort" in s or "dp" in s) else "")
    if "artcnn" in s:
        return "artcnn-c4f%s" % ("32" if "f32" in s else "16") + ("-dn" if "dn" in s else ("-ds" if "ds" in s else ""))
    if "arnet" in s:
        b = "b8"
        for cand in ("b8", "b16", "b32", "b64"):
            if cand in s:
                b = cand
                break
        return "arnet-f8" + b + ("-box" if "box" in s else "") + ("-hdn" if "hdn" in s else "")
    if "acnet" in s and "legacy" not in s and "fsrcnnx" not in s and "artcnn" not in s:
        b = "b8"
        for cand in ("b4", "b8", "b18"):
            if cand in s:
                b = cand
                break
        return "acnet-f8" + b + ("-box" if "box" in s else "") + ("-hdn" if "hdn" in s else "")
    if "acnet" in s and "legacy" in s and "hdn" in s:
        for ch in s:
            if ch in "0123":
                return "acnet-legacy-hdn" + ch
        return "acnet-legacy-hdn0"
    return "acnet-legacy-gan"


def _ptr(a):
    return a.ctypes.data_as(_fp) if a is not None and a.size else C.cast(None, _fp)


def _type_of(img):
    return NP_TYPES[img.dtype]


def oracle_process(name, img, factor=2.0):
    """Processor::process(src, factor) on the CPU oracle.  img: (H,W) or (H,W,C) contiguous."""
    fam, blocks, k, b, a = model(name)
    if fam >= FAMILY_ARTCNN:
        blocks |= features(name) << 16     # the oracle's C entry takes F in the upper half of `blocks` for these families
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty((int(h * factor), int(w * factor)) + (() if img.ndim == 2 else (c,)), img.dtype)
    rc = oracle().orc_process(fam, blocks, _ptr(k), _ptr(b), _ptr(a), img.ctypes.data, w, h, c, img.strides[0], _type_of(img),
                              float(factor), out.ctypes.data, out.strides[0])
    assert rc == 0
    return out


def ref_process(name, img, factor=2.0, arch=1):
    """Processor::process on the compiled reference (arch 1 = Generic, 0 = auto ISA)."""
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty((int(h * factor), int(w * factor)) + (() if img.ndim == 2 else (c,)), img.dtype)
    rc = ref().ref_process(name.encode(), arch, img.ctypes.data, w, h, c, img.strides[0], _type_of(img), float(factor),
                           out.ctypes.data, out.strides[0])
    assert rc == 0
    return out


def oracle_resize(img, ow, oh):
    """ac::core::resize(src, dst, 0, 0) (Catmull-Rom) on the CPU oracle; img (H,W) or (H,W,C)."""
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty((oh, ow) + (() if img.ndim == 2 else (c,)), img.dtype)
    assert oracle().orc_resize_catmull_rom(img.ctypes.data, w, h, c, img.strides[0], _type_of(img), out.ctypes.data, ow, oh, out.strides[0]) == 0
    return out


def oracle_frame(name, planes, factor=2.0, shift=0):
    """The reference's per-frame video callback (cli/src/Main.cpp:183-206) composed from oracle pieces:
    luma: shl(shift) -> Processor::process -> shr(shift) (core/src/ImageProcess.cpp:601-616: integer types only);
    every other plane: ac::core::resize(srcp, dstp, 0.0, 0.0) to factor x its size."""
    y = planes[0]
    integer = y.dtype in (np.uint8, np.uint16)
    if shift and integer:
        y = np.left_shift(y.astype(np.int64), shift).astype(y.dtype)        # `a << n` truncated to the element type
    oy = oracle_process(name, y, factor)
    if shift and integer:
        oy = np.right_shift(oy, shift).astype(oy.dtype)
    return [oy] + [oracle_resize(p, int(p.shape[1] * factor), int(p.shape[0] * factor)) for p in planes[1:]]


# ---------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------
def noise_u8(h, w, c=1, seed=1234):
    """mt19937 bytes in 4-byte words, the reference generators' shape (tools/benchmark/src/Benchmark.cpp:18-37)
    with a fixed seed."""
    rs = np.random.RandomState(seed)
    n = h * w * c
    words = rs.randint(0, 1 << 32, size=(n + 3) // 4, dtype=np.uint64).astype(np.uint32)
    a = words.view(np.uint8)[:n]
    return a.reshape((h, w) if c == 1 else (h, w, c)).copy()


def smooth_u8(h, w, c=1, seed=7):
    """anime-like flats and edges: smooth field + ~2% salt noise."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    chans = []
    rs = np.random.RandomState(seed)
    for ch in range(c):
        f = (np.sin(x / (9.0 + ch)) + np.cos(y / (13.0 - ch)) + 2.0) / 4.0 * 255.0
        f = np.where(((x // 17 + y // 23) % 5) == 0, 255.0 - f, f)
        img = f.astype(np.uint8)
        salt = rs.rand(h, w) < 0.02
        img[salt] = rs.randint(0, 256, size=int(salt.sum()), dtype=np.uint8)
        chans.append(img)
    return chans[0] if c == 1 else np.stack(chans, axis=-1)


def compare_u8(a, b):
    """(max abs diff, fraction of elements bit-exact)."""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return int(d.max()), float((d == 0).mean())


ORDER_GENERIC, ORDER_FMA = 0, 1


def set_order(order):
    """Summation order of the oracle's network layers: ORDER_GENERIC (reference `create("cpu", 1, ...)`) or
    ORDER_FMA (what the reference's auto-ISA backend executes on FMA-capable x86)."""
    oracle().orc_set_order(order)
