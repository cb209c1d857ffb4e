"""CPU suite, part 2: the native library loads, exports every symbol include/*.h declares, and its host logic
(model strings, weight tables, Image semantics through the C binding, loud failure without a device) behaves like
the reference's.  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import anime4kcpp_b200 as A
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for path in ("include/acb200.h", "include/AC/Core/Image.h", "include/AC/Core/Processor.h"):
        text = open(os.path.join(ROOT, path)).read()
        names |= set(re.findall(r"(?:ACB200_API|AC_C_API)\s+[\w\s\*]+?\b(acb200_\w+|ac_\w+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = A.lib()
    syms = _declared_symbols()
    assert len(syms) >= 45
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s


def test_weights_in_library_match_blob_and_synth():
    for name in list(O.models().keys()) + ["arnet-f8b8", "arnet-f8b32-box-hdn"]:
        fam, blocks, k, b, a = A.model_arrays(name)
        ofam, oblocks, ok, ob, oa = O.model(name)
        assert (fam, blocks) == (ofam, oblocks)
        assert np.array_equal(k.view(np.uint32), ok.view(np.uint32))
        assert np.array_equal(b.view(np.uint32), ob.view(np.uint32))
        assert np.array_equal(a.view(np.uint32), oa.view(np.uint32))


@pytest.mark.parametrize("s", ["acnet-hdn", "ACNet-Legacy-HDN2", "acnet-legacy", "acnet-legacy-hdn", "arnet-f8b16-box", "arnet", "ARNET-B64-hdn",
                               "acnet-f8b18-box-hdn", "acnet-b4", "unknown", "", "acnet-f8b8-box",
                               "artcnn", "ArtCNN-C4F32-DS", "artcnn-c4f16-dn", "artcnn-f32", "fsrcnnx", "fsrcnnx-f16b4-dp", "FSRCNNX-F8B4-Distort-Plus",
                               "fsrcnnx-f16"])
def test_model_string_resolution_matches_reference_rules(s):
    assert A.resolve_model(s) == O.canonical(s)


def test_wide_families_resolve_with_their_feature_count():
    """ArtCNN<16/32> / FSRCNNX<8/16> (core/src/processor/Processor.cpp:39-76): family, features and array lengths."""
    for name, fam, feat in (("artcnn-c4f16", 3, 16), ("artcnn-c4f32-dn", 3, 32), ("fsrcnnx-f8b4", 4, 8), ("fsrcnnx-f16b4-distort-plus", 4, 16)):
        f, blocks, k, b, a = A.model_arrays(name)
        assert (f, blocks) == (fam, 4) and A.lib().ac_b200_model_features(name.encode()) == feat == O.features(name)
        m = A.Model(name)          # lengths accepted by acb200_model_create_wide
        assert m.features == feat and m.halo() == 7
        with pytest.raises(A.Acb200Error):
            A.Model(family=f, blocks=blocks, kernels=k, biases=b, alphas=a, features=8 if feat != 8 else 16)


def test_model_create_validates_lengths():
    fam, blocks, k, b, a = A.model_arrays("acnet-f8b4")
    A.Model(family=fam, blocks=blocks, kernels=k, biases=b, alphas=a)
    with pytest.raises(A.Acb200Error):
        A.Model(family=fam, blocks=blocks, kernels=k[:-1], biases=b, alphas=a)
    with pytest.raises(A.Acb200Error):
        A.Model(family=fam, blocks=5, kernels=k, biases=b, alphas=a)


class ACImage(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("stride", C.c_int), ("element_type", C.c_int),
                ("ptr", C.c_void_p), ("hptr", C.c_void_p)]


class ACProcessor(C.Structure):
    _fields_ = [("device", C.c_int), ("type", C.c_char_p), ("model", C.c_char_p), ("hptr", C.c_void_p)]


def _cimage(lib, w, h, c, t, stride=0):
    lib.ac_image_alloc.restype = C.POINTER(ACImage)
    img = lib.ac_image_alloc()
    img.contents.width, img.contents.height, img.contents.channels, img.contents.element_type, img.contents.stride = w, h, c, t, stride
    return img


def test_c_binding_image_semantics():
    # reference tests/core/src/ImageTest.cpp: create / stride alignment / from / to / view / clone / ref
    lib = A.lib()
    img = _cimage(lib, 5, 3, 3, 1)
    assert lib.ac_image_create(img) == 0
    assert img.contents.stride == 16 and img.contents.ptr      # 15-byte line rounded up to 4
    data = np.arange(45, dtype=np.uint8).reshape(3, 15)
    src = _cimage(lib, 5, 3, 3, 1)
    assert lib.ac_image_from(src, data.ctypes.data_as(C.c_void_p)) == 0
    back = np.zeros((3, 15), np.uint8)
    assert lib.ac_image_to(src, back.ctypes.data_as(C.c_void_p), 0) == 0
    assert np.array_equal(back, data)
    view = _cimage(lib, 0, 0, 0, 0)
    assert lib.ac_image_view(src, view, 1, 1, 100, 100) == 0           # clipped to the image
    assert (view.contents.width, view.contents.height, view.contents.stride) == (4, 2, src.contents.stride)
    assert view.contents.ptr == src.contents.ptr + src.contents.stride + 3
    clone = _cimage(lib, 0, 0, 0, 0)
    assert lib.ac_image_clone(view, clone) == 0
    assert (clone.contents.width, clone.contents.height) == (4, 2) and clone.contents.ptr != view.contents.ptr
    mapped = _cimage(lib, 15, 3, 1, 1)
    mapped.contents.ptr = data.ctypes.data
    assert lib.ac_image_map(mapped) == 0 and mapped.contents.ptr == data.ctypes.data and mapped.contents.stride == 15
    assert lib.ac_image_ref(None, img) == -22 and lib.ac_image_to(img, None, 0) == -22        # -AC_EINVAL
    for im in (img, src, view, clone, mapped):
        lib.ac_image_free(C.byref(im))
        assert not im


def test_c_binding_processor_without_compute():
    lib = A.lib()
    lib.ac_processor_alloc.restype = C.POINTER(ACProcessor)
    lib.ac_processor_error.restype = C.c_char_p
    lib.ac_processor_type_name.restype = C.c_char_p
    lib.ac_processor_list_info.restype = C.c_char_p
    lib.ac_processor_info.restype = C.c_char_p
    assert lib.ac_processor_ok(None) == -22
    assert b"CUDA:" in lib.ac_processor_list_info() and lib.ac_processor_info(2).startswith(b"CUDA:")
    p = lib.ac_processor_alloc()
    p.contents.type, p.contents.model, p.contents.device = b"cpu", b"acnet-legacy-hdn0", 0
    assert lib.ac_processor_create(p) == -256        # -AC_EPROCESSOR: no CPU backend, reported not thrown
    assert b"no CPU backend" in lib.ac_processor_error(p)
    lib.ac_processor_free(C.byref(p))
    p = lib.ac_processor_alloc()
    p.contents.type, p.contents.model, p.contents.device = b"cuda", b"acnet-legacy-hdn0", 0
    rc = lib.ac_processor_create(p)
    if A.device_count() == 0:
        assert rc == -256 and lib.ac_processor_error(p) == b"no CUDA device"      # fails loudly, no fallback
    else:
        assert rc == 0 and lib.ac_processor_type_name(p) == b"CUDA" and lib.ac_processor_type(p) == 2
    lib.ac_processor_free(C.byref(p))


def test_pyac_surface():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "anime4kcpp_b200"))
    import pyac
    assert pyac.core.Processor.CUDA == 2 and pyac.core.Processor.CPU == 0 and pyac.core.Processor.OpenCL == 1
    assert int(pyac.core.RESIZE_CATMULL_ROM) == 1 and int(pyac.core.RESIZE_BILINEAR) == 16 and int(pyac.core.IMREAD_RGBA) == 4
    names = [m["name"] for m in pyac.specs.ModelList]
    assert "acnet-legacy-hdn0" in names and "arnet-f8b64" in names and "artcnn-c4f32-ds" in names and "fsrcnnx-f16b4" in names and len(names) == 43
    p = pyac.core.Processor("cpu", 0, "acnet-legacy-hdn0")
    assert not p.ok() and "CPU" in p.error()
    with pytest.raises(RuntimeError):
        p(np.zeros((4, 4), np.uint8))
    if A.device_count() == 0:
        q = pyac.core.Processor()            # defaults: auto, 0, acnet-f8b8-hdn
        assert not q.ok() and q.error() == "no CUDA device"


def test_resize_with_an_unsupported_mode_is_an_error_not_unwritten_memory():
    """The reference defaults pyac.core.resize / ac_resize to RESIZE_BILINEAR; only RESIZE_CATMULL_ROM exists on this path.  The call
    must fail through the error path (exception / -AC_EINVAL) and leave the destination untouched -- before any GPU work."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "anime4kcpp_b200"))
    import pyac
    src = np.arange(48, dtype=np.uint8).reshape(6, 8)
    with pytest.raises(ValueError):
        pyac.core.resize(src, (16, 12))                                     # default mode = RESIZE_BILINEAR
    with pytest.raises(ValueError):
        pyac.core.resize(src, None, 2.0, 2.0, pyac.core.RESIZE_BILINEAR)
    lib = A.lib()
    lib.ac_resize.argtypes = [C.POINTER(ACImage), C.POINTER(ACImage), C.c_double, C.c_double, C.c_int]
    a = _cimage(lib, 8, 6, 1, 1)
    assert lib.ac_image_from(a, src.ctypes.data_as(C.c_void_p)) == 0
    b = _cimage(lib, 0, 0, 0, 0)
    assert lib.ac_resize(a, b, 2.0, 2.0, 16) == -22                         # RESIZE_BILINEAR -> -AC_EINVAL
    assert b.contents.width == 0 and not b.contents.ptr                     # nothing allocated, nothing published
    for im in (a, b):
        lib.ac_image_free(C.byref(im))


def test_reference_benchmark_tool_links_against_the_drop_in():
    """oracle/_ref/ac_benchmark = the reference's tools/benchmark/src/Benchmark.cpp compiled unchanged against the re-created headers and
    libac_b200.so (oracle/Makefile ref_callers; built where /root/reference exists).  Without a GPU it must start, list the backends and
    report the processor error the way the reference tool does -- not crash, not fall back."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "ac_benchmark")
    if not os.path.isfile(exe):
        pytest.skip("oracle/_ref/ac_benchmark not built (needs /root/reference at build time)")
    out = subprocess.run([exe, "acnet-legacy-hdn0", "cuda", "0", "64", "48", "4", "2"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "core version:" in out.stdout and "CUDA:" in out.stdout, out.stdout + out.stderr
    if A.device_count() == 0:
        assert "no CUDA device" in out.stdout and "FPS:" not in out.stdout
    else:
        assert "FPS:" in out.stdout


def _png_bytes(arr, depth=8, palette=None, interlace=0):
    """A PNG file built by hand with zlib (filter 0 on even rows, filter 2 'up' on odd rows), for the decoder tests."""
    import struct
    import zlib

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    h, w = arr.shape[:2]
    c = 1 if arr.ndim == 2 else arr.shape[2]
    ctype = 3 if palette is not None else {1: 0, 2: 4, 3: 2, 4: 6}[c]
    rows = arr.reshape(h, -1).astype(">u2" if depth == 16 else np.uint8)
    raw = b""
    prev = np.zeros(rows.shape[1] * rows.dtype.itemsize, np.uint8)
    for y in range(h):
        cur = np.frombuffer(rows[y].tobytes(), np.uint8)
        if y % 2:
            raw += b"\x02" + ((cur.astype(np.int16) - prev) & 255).astype(np.uint8).tobytes()
        else:
            raw += b"\x00" + cur.tobytes()
        prev = cur
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
    if palette is not None:
        out += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    return out + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")


def test_image_file_io_roundtrips_and_decodes(tmp_path):
    """ac::core::imread / imwrite (reference core/src/ImageIO.cpp:20-87) on the drop-in's own codecs, through pyac and the C binding:
    write -> read round trips for .png / .bmp / .tga, hand-built PNGs (8 / 16 bit, palette, padded rows), PNM, the decoder's channel
    conversions, and the failure paths (JPEG, missing file, interlaced PNG) -- all without a GPU."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "anime4kcpp_b200"))
    import pyac
    rs = np.random.RandomState(7)
    for c in (1, 3, 4):
        img = rs.randint(0, 256, size=(37, 53) if c == 1 else (37, 53, c), dtype=np.uint8)
        for ext in ("png", "bmp", "tga"):
            path = str(tmp_path / ("rt%d.%s" % (c, ext)))
            assert pyac.core.imwrite(path, img)
            back = pyac.core.imread(path)
            want = img if not (ext == "bmp" and c == 1) else np.repeat(img[:, :, None], 3, axis=2)     # BMP has no gray form
            assert back.dtype == np.uint8 and back.shape == want.shape and np.array_equal(back, want), (c, ext)
    # a strided view (padded rows) is written row by row
    big = rs.randint(0, 256, size=(20, 64, 3), dtype=np.uint8)
    view = big[:, 5:40]
    assert pyac.core.imwrite(str(tmp_path / "view.png"), view) and np.array_equal(pyac.core.imread(str(tmp_path / "view.png")), view)
    # hand-built PNGs: 8-bit RGB / gray+alpha with the 'up' filter, 16-bit gray (high byte kept), palette
    rgb = rs.randint(0, 256, size=(9, 11, 3), dtype=np.uint8)
    (tmp_path / "a.png").write_bytes(_png_bytes(rgb))
    assert np.array_equal(pyac.core.imread(str(tmp_path / "a.png")), rgb)
    ga = rs.randint(0, 256, size=(6, 7, 2), dtype=np.uint8)
    (tmp_path / "ga.png").write_bytes(_png_bytes(ga))
    assert np.array_equal(pyac.core.imread(str(tmp_path / "ga.png")), ga)
    g16 = rs.randint(0, 65536, size=(5, 8)).astype(np.uint16)
    (tmp_path / "g16.png").write_bytes(_png_bytes(g16, depth=16))
    assert np.array_equal(pyac.core.imread(str(tmp_path / "g16.png")), (g16 >> 8).astype(np.uint8))
    pal = rs.randint(0, 256, size=(16, 3), dtype=np.uint8)
    idx = rs.randint(0, 16, size=(7, 9)).astype(np.uint8)
    (tmp_path / "p.png").write_bytes(_png_bytes(idx, palette=pal))
    assert np.array_equal(pyac.core.imread(str(tmp_path / "p.png")), pal[idx])
    # PNM
    (tmp_path / "g.pgm").write_bytes(b"P5\n# comment\n11 9\n255\n" + rgb[:, :, 0].tobytes())
    assert np.array_equal(pyac.core.imread(str(tmp_path / "g.pgm")), rgb[:, :, 0])
    (tmp_path / "c.ppm").write_bytes(b"P6 11 9 255\n" + rgb.tobytes())
    assert np.array_equal(pyac.core.imread(str(tmp_path / "c.ppm")), rgb)
    # channel conversions of the decoder: luma = (77 r + 150 g + 29 b) >> 8, alpha filled with 255 / dropped
    gray = pyac.core.imread(str(tmp_path / "a.png"), pyac.core.IMREAD_GRAYSCALE)
    want = ((rgb[:, :, 0].astype(np.int32) * 77 + rgb[:, :, 1].astype(np.int32) * 150 + rgb[:, :, 2].astype(np.int32) * 29) >> 8).astype(np.uint8)
    assert gray.shape == (9, 11) and np.array_equal(gray, want)
    rgba = pyac.core.imread(str(tmp_path / "a.png"), pyac.core.IMREAD_RGBA)
    assert rgba.shape == (9, 11, 4) and np.array_equal(rgba[:, :, :3], rgb) and (rgba[:, :, 3] == 255).all()
    assert np.array_equal(pyac.core.imread(str(tmp_path / "g.pgm"), pyac.core.IMREAD_COLOR), np.repeat(rgb[:, :, :1], 3, axis=2))
    # failure paths: never a crash, never garbage
    assert not pyac.core.imwrite(str(tmp_path / "x.jpg"), rgb) and not pyac.core.imwrite(str(tmp_path / "noext"), rgb)
    with pytest.raises(RuntimeError):
        pyac.core.imread(str(tmp_path / "missing.png"))
    (tmp_path / "i.png").write_bytes(_png_bytes(rgb, interlace=1))
    with pytest.raises(RuntimeError):
        pyac.core.imread(str(tmp_path / "i.png"))
    (tmp_path / "t.png").write_bytes(_png_bytes(rgb)[:60])
    with pytest.raises(RuntimeError):
        pyac.core.imread(str(tmp_path / "t.png"))
    # the C binding: ac_imread / ac_imwrite with the reference's return codes
    lib = A.lib()
    lib.ac_imread.argtypes = [C.c_char_p, C.c_int, C.POINTER(ACImage)]
    lib.ac_imwrite.argtypes = [C.c_char_p, C.POINTER(ACImage)]
    im = _cimage(lib, 0, 0, 0, 0)
    assert lib.ac_imread(str(tmp_path / "a.png").encode(), 0, im) == 0
    assert (im.contents.width, im.contents.height, im.contents.channels, im.contents.element_type) == (11, 9, 3, 1)
    got = np.ctypeslib.as_array(C.cast(im.contents.ptr, C.POINTER(C.c_ubyte)), shape=(9, im.contents.stride))[:, :33].reshape(9, 11, 3)
    assert np.array_equal(got, rgb)
    assert lib.ac_imwrite(str(tmp_path / "c.bmp").encode(), im) == 0 and np.array_equal(pyac.core.imread(str(tmp_path / "c.bmp")), rgb)
    assert lib.ac_imread(str(tmp_path / "missing.png").encode(), 0, im) == -5 and lib.ac_imwrite(str(tmp_path / "c.jpg").encode(), im) == -5       # -AC_EIO
    lib.ac_image_free(C.byref(im))


def test_pinned_host_allocation_entry_points():
    """acb200_host_alloc / acb200_host_free (page-locked staging memory, also what ac::core::Image's pool is built on): NULL without a
    usable device -- callers fall back to their own memory, nothing is emulated -- and a writable block with one."""
    lib = A.lib()
    lib.acb200_host_alloc.restype = C.c_void_p
    lib.acb200_host_alloc.argtypes = [C.c_size_t]
    lib.acb200_host_free.argtypes = [C.c_void_p]
    lib.acb200_host_free.restype = None
    p = lib.acb200_host_alloc(1 << 20)
    if A.device_count() == 0:
        assert not p
    else:
        assert p
        C.memset(p, 0x5a, 1 << 20)
        assert C.string_at(p + (1 << 20) - 1, 1) == b"\x5a"
        lib.acb200_host_free(p)
    lib.acb200_host_free(None)          # harmless
    assert not lib.acb200_host_alloc(0)


def test_session_without_device_fails_loudly():
    if A.device_count() == 0:
        with pytest.raises(A.Acb200Error):
            A.Session(0)


def test_cpp_model_descriptors_like_reference_modeltest(tmp_path):
    """tests/cpp/model_check.cpp: the reference's ModelTest.cpp checks (offsets / lengths of every layer add up) over all 43
    variants of the five families, compiled against the re-created headers and linked with libac_b200.so."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++ on this box")
    exe = str(tmp_path / "model_check")
    libdir = os.path.join(ROOT, "anime4kcpp_b200", "lib")
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "anime4kcpp_b200", "csrc", "host", "include"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "model_check.cpp"), "-o", exe, "-L" + libdir, "-lac_b200", "-Wl,-rpath," + libdir]
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(cmd, check=True, env=env)
    out = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and "model checks ok" in out.stdout, out.stdout + out.stderr
