"""GPU suite: the CUDA path through the C-ABI against the CPU oracle (same seeded inputs) and against the golden
vectors minted from the compiled reference.  Bars (BASELINE.json north_star): 8/16-bit output <= 1 LSB per channel and
>= 99.9 % of samples bit-exact; float32 max abs error <= 1e-3."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import anime4kcpp_b200 as A
import oracle_lib as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
ALL_REAL = sorted(O.models().keys())
ARNETS = ["arnet-f8b8", "arnet-f8b16-hdn", "arnet-f8b32-box", "arnet-f8b64-box-hdn"]

# Bars (BASELINE.json north_star): 8-bit output <= 1 LSB per channel and >= 99.9 % bit-exact; float32 max abs <= 1e-3.
LSB_MAX = 1
EXACT_MIN = 0.999
F32_TOL = 1e-3
# 16-bit output is 256x finer than the 8-bit bar: hold it to 4 LSB16 (1/64 of an 8-bit LSB)
U16_LSB_MAX = 4
# Two 2x passes compound: the reference's OWN backends (FMA vs Generic) differ by up to 4 LSB / ~0.2 % of samples at 4x
# (measured in this repo: tests/test_oracle_cpu.py::test_reference_isa_spread_4x), so against Generic-order vectors a 4x
# result is held to that spread; against FMA-order vectors the exact engine must still be bit-identical.
X4_LSB_MAX, X4_EXACT_MIN = 4, 0.995
ENGINE_EXACT, ENGINE_TENSOR, ENGINE_AUTO = 0, 1, 2
ENGINES = [ENGINE_EXACT]
# tensor engine implementations: 0 = mma.sync, 1 = tcgen05 with the maps in shared memory, 2 = tcgen05 with the maps resident in
# TMEM (library default; ARNet segments fall back to 0 inside the library); the tensor-engine tests run all three
TENSOR_IMPL = 2
TENSOR_IMPLS = [0, 1, 2]


@pytest.fixture(scope="module")
def session():
    assert A.device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    return A.Session(0)


@pytest.fixture(autouse=True)
def _orders(session):
    O.set_order(O.ORDER_GENERIC)
    session.set_engine(ENGINE_EXACT)
    session.set_tensor_impl(TENSOR_IMPL)
    yield
    O.set_order(O.ORDER_GENERIC)


_models = {}


def gpu_model(name):
    if name not in _models:
        _models[name] = A.Model(name)
    return _models[name]


def check_close(out, want, x4=False):
    """The north_star tolerance against a Generic-order (or any differently rounded) reference result."""
    if out.dtype == np.float32:
        assert float(np.abs(out - want).max()) <= F32_TOL
        return
    mx, exact = O.compare_u8(out, want)
    if out.dtype == np.uint16:
        assert mx <= U16_LSB_MAX * (4 if x4 else 1), "max diff %d LSB16" % mx
        return
    lim, frac = (X4_LSB_MAX, X4_EXACT_MIN) if x4 else (LSB_MAX, EXACT_MIN)
    assert mx <= lim and exact >= frac, "max diff %d LSB, %.4f %% exact" % (mx, 100 * exact)


def check_int(out, want):
    check_close(out, want)


def src_for(kind):
    return {"gray_noise_2x": GOLD["in_gray_noise"], "gray_smooth_2x": GOLD["in_gray_smooth"], "rgb_2x": GOLD["in_rgb"], "rgb_4x": GOLD["in_rgb"],
            "rgba_2x": GOLD["in_rgba"], "gray_4x": GOLD["in_gray_noise"][:20, :24],
            "gray_f32_2x": GOLD["in_gray_noise"].astype(np.float32) / np.float32(255),
            "gray_u16_2x": GOLD["in_gray_noise"].astype(np.uint16) * 257}[kind]


GOLD_KEYS = sorted(k for k in GOLD.files if "/" in k and not k.startswith("fma:"))


@pytest.mark.parametrize("key", GOLD_KEYS)
def test_exact_engine_is_bit_identical_to_reference_fma_backend(session, key):
    """Vectors minted from the compiled reference's FMA backend (what its auto-ISA processor executes on x86):
    the exact engine reproduces them bit for bit -- u8, u16 and f32, gray / RGB / RGBA, 2x and 4x, every model."""
    kind, name = key.split("/", 1)
    out = session.process_host(gpu_model(name), src_for(kind), 4.0 if kind.endswith("4x") else 2.0)
    assert np.array_equal(out, GOLD["fma:" + key])


@pytest.mark.parametrize("key", GOLD_KEYS)
def test_golden_generic_backend_within_tolerance(session, key):
    """Vectors minted from the reference's Generic backend (`create("cpu", 1, ...)`, ground truth of its own tests)."""
    kind, name = key.split("/", 1)
    out = session.process_host(gpu_model(name), src_for(kind), 4.0 if kind.endswith("4x") else 2.0)
    check_close(out, GOLD[key], x4=kind.endswith("4x"))


@pytest.mark.parametrize("name", ["acnet-legacy-hdn0", "acnet-legacy-gan", "acnet-f8b4", "acnet-f8b8-hdn", "acnet-f8b18-box-hdn", "arnet-f8b8", "arnet-f8b16"])
@pytest.mark.parametrize("shape", [(3, 3), (1, 7), (9, 1), (17, 31), (40, 40), (41, 39), (64, 64), (97, 131), (255, 257)])
def test_odd_sizes_gray_vs_oracle(session, name, shape):
    img = O.noise_u8(shape[0], shape[1], 1, seed=shape[0] * 1000 + shape[1])
    out = session.process_host(gpu_model(name), img, 2.0)
    check_close(out, O.oracle_process(name, img, 2.0))
    O.set_order(O.ORDER_FMA)
    assert np.array_equal(out, O.oracle_process(name, img, 2.0))


@pytest.mark.parametrize("name", ["acnet-legacy-hdn1", "acnet-f8b8", "arnet-f8b8-box"])
@pytest.mark.parametrize("c", [3, 4])
@pytest.mark.parametrize("factor", [2.0, 4.0])
def test_colour_vs_oracle(session, name, c, factor):
    img = O.noise_u8(45, 61, c, seed=c * 10 + int(factor))
    f = img.astype(np.float32) / np.float32(255)
    u = img.astype(np.uint16) * 257
    outs = [session.process_host(gpu_model(name), x, factor) for x in (img, f, u)]
    for x, out in zip((img, f, u), outs):
        want = O.oracle_process(name, x, factor)
        if c == 4:
            # yuva2rgba divides by alpha (ImageProcess.cpp:289-295): a rounding-order difference in the luma is amplified
            # by 1/alpha, between the reference's own backends too -- hold the tolerance where alpha >= 1/2 only
            opaque = want[..., 3].astype(np.float64) >= 0.5 * (1.0 if x.dtype == np.float32 else float(np.iinfo(x.dtype).max))
            check_close(out[opaque], want[opaque], x4=True)
        else:
            check_close(out, want, x4=factor == 4.0)
    O.set_order(O.ORDER_FMA)
    for x, out in zip((img, f, u), outs):
        assert np.array_equal(out, O.oracle_process(name, x, factor))


# ---- tensor-core engine (split-fp16 MMA): the 8-bit bar against BOTH reference orders -------------------------------------
@pytest.mark.parametrize("key", GOLD_KEYS)
@pytest.mark.parametrize("impl", TENSOR_IMPLS)
def test_tensor_engine_golden(session, key, impl):
    session.set_tensor_impl(impl)
    kind, name = key.split("/", 1)
    x4 = kind.endswith("4x")
    session.set_engine(ENGINE_TENSOR)
    out = session.process_host(gpu_model(name), src_for(kind), 4.0 if x4 else 2.0)
    check_close(out, GOLD[key], x4=x4)
    check_close(out, GOLD["fma:" + key], x4=x4)
    if x4:
        # default engine: exact for the first pass, tensor for the last -> the 2x bar holds at 4x against the FMA-order vectors
        session.set_engine(ENGINE_AUTO)
        check_close(session.process_host(gpu_model(name), src_for(kind), 4.0), GOLD["fma:" + key])


@pytest.mark.parametrize("name", ["acnet-legacy-hdn0", "acnet-legacy-gan", "acnet-f8b4", "acnet-f8b8-hdn", "acnet-f8b18-box-hdn", "arnet-f8b8", "arnet-f8b16"])
@pytest.mark.parametrize("shape", [(3, 3), (1, 7), (9, 1), (17, 31), (40, 40), (41, 39), (64, 64), (97, 131), (255, 257)])
@pytest.mark.parametrize("impl", TENSOR_IMPLS)
def test_tensor_engine_odd_sizes(session, name, shape, impl):
    session.set_tensor_impl(impl)
    img = O.noise_u8(shape[0], shape[1], 1, seed=shape[0] * 1000 + shape[1])
    session.set_engine(ENGINE_TENSOR)
    out = session.process_host(gpu_model(name), img, 2.0)
    O.set_order(O.ORDER_FMA)
    check_close(out, O.oracle_process(name, img, 2.0))


@pytest.mark.parametrize("name", ["acnet-legacy-hdn0", "acnet-f8b8-hdn", "arnet-f8b8"])
@pytest.mark.parametrize("impl", TENSOR_IMPLS)
def test_tensor_engine_1080p_against_exact_engine(session, name, impl):
    session.set_tensor_impl(impl)
    # full BASELINE size: the exact engine (bit-identical to the reference) is the checker for the tensor engine
    img = O.smooth_u8(1080, 1920, 1, seed=3)
    m = gpu_model(name)
    session.set_engine(ENGINE_EXACT)
    want = session.process_host(m, img, 2.0)
    session.set_engine(ENGINE_TENSOR)
    got = session.process_host(m, img, 2.0)
    mx, exact = O.compare_u8(got, want)
    assert mx <= 1 and exact >= 0.9995, (mx, exact)
    noise = O.noise_u8(540, 960, 1, seed=4)
    session.set_engine(ENGINE_EXACT)
    want = session.process_host(m, noise, 2.0)
    session.set_engine(ENGINE_TENSOR)
    mx, exact = O.compare_u8(session.process_host(m, noise, 2.0), want)
    assert mx <= 1 and exact >= 0.999, (mx, exact)


@pytest.mark.parametrize("impl", TENSOR_IMPLS)
def test_tensor_engine_colour_and_types(session, impl):
    session.set_tensor_impl(impl)
    for name in ("acnet-legacy-hdn1", "acnet-f8b8", "arnet-f8b8-box"):
        img = O.noise_u8(45, 61, 3, seed=17)
        O.set_order(O.ORDER_FMA)
        session.set_engine(ENGINE_TENSOR)
        check_close(session.process_host(gpu_model(name), img, 2.0), O.oracle_process(name, img, 2.0))
        f = img.astype(np.float32) / np.float32(255)
        check_close(session.process_host(gpu_model(name), f, 2.0), O.oracle_process(name, f, 2.0))
        u = img.astype(np.uint16) * 257
        check_close(session.process_host(gpu_model(name), u, 2.0), O.oracle_process(name, u, 2.0))
        session.set_engine(ENGINE_AUTO)
        check_close(session.process_host(gpu_model(name), img, 4.0), O.oracle_process(name, img, 4.0))


@pytest.mark.parametrize("value", [0, 128, 255])
def test_constant_and_checkerboard(session, value):
    # saturation and clamp-to-edge checks (SURVEY 8d iii)
    img = np.full((50, 70), value, np.uint8)
    for name in ("acnet-legacy-hdn0", "acnet-f8b8-hdn"):
        check_int(session.process_host(gpu_model(name), img, 2.0), O.oracle_process(name, img, 2.0))
    yy, xx = np.mgrid[0:50, 0:70]
    cb = (((yy + xx) & 1) * 255).astype(np.uint8)
    check_int(session.process_host(gpu_model("acnet-legacy-hdn0"), cb, 2.0), O.oracle_process("acnet-legacy-hdn0", cb, 2.0))


def test_strided_view_input_and_preallocated_dst(session):
    # filter frontends hand mapped, padded planes (filter/vapoursynth/src/Filter.cpp:32-42); a non-empty dst is written in place
    big = O.noise_u8(80, 128, 1, seed=3)
    view = big[7:60, 13:90]
    assert not view.flags["C_CONTIGUOUS"]
    out_big = np.zeros((130, 256), np.uint8)
    out = out_big[5:5 + 106, 11:11 + 154]
    rc = A.lib().acb200_process_host(session.handle, gpu_model("acnet-legacy-hdn0").handle, view.ctypes.data, 77, 53, 1, view.strides[0], A.UINT8, 2.0,
                                     out.ctypes.data, out.strides[0])
    assert rc == 0
    check_int(out, O.oracle_process("acnet-legacy-hdn0", np.ascontiguousarray(view), 2.0))
    assert out_big[:5].sum() == 0 and out_big[:, :11].sum() == 0 and out_big[:, 11 + 154:].sum() == 0     # nothing outside the view was touched


def test_colour_conversion_bit_exact(session):
    rgb = GOLD["in_rgb"]
    y, uv = session.rgb2yuv(rgb)
    assert np.array_equal(y, GOLD["rgb2yuv_y"]) and np.array_equal(uv, GOLD["rgb2yuv_uv"])
    assert np.array_equal(session.yuv2rgb(y, uv), GOLD["yuv2rgb_back"])
    rgba = O.noise_u8(31, 37, 4, seed=8)
    yo = np.empty((31, 37), np.uint8)
    uvo = np.empty((31, 37, 3), np.uint8)
    O.oracle().orc_rgb2yuv(rgba.ctypes.data, 37, 31, 4, rgba.strides[0], O.U8, yo.ctypes.data, yo.strides[0], uvo.ctypes.data, uvo.strides[0])
    y, uva = session.rgb2yuv(rgba)
    assert np.array_equal(y, yo) and np.array_equal(uva, uvo)
    back = np.empty_like(rgba)
    O.oracle().orc_yuv2rgb(yo.ctypes.data, yo.strides[0], uvo.ctypes.data, uvo.strides[0], 37, 31, 4, O.U8, back.ctypes.data, back.strides[0])
    assert np.array_equal(session.yuv2rgb(y, uva), back)


@pytest.mark.parametrize("scale", [2, 4])
@pytest.mark.parametrize("c", [1, 2, 3])
def test_catmull_rom_resize_bit_exact_vs_oracle(session, scale, c):
    img = O.noise_u8(23, 29, c, seed=scale * 7 + c)
    want = np.empty((23 * scale, 29 * scale) + (() if c == 1 else (c,)), np.uint8)
    assert O.oracle().orc_resize_catmull_rom(img.ctypes.data, 29, 23, c, img.strides[0], O.U8, want.ctypes.data, 29 * scale, 23 * scale, want.strides[0]) == 0
    assert np.array_equal(session.resize_catmull_rom(img, 29 * scale, 23 * scale), want)


def test_full_size_1080p_properties(session):
    # BASELINE config size: the oracle is too slow for every model here, so check one model against the oracle on the
    # full frame and use size-independent properties for the rest: tile seams (a crop with full halo reproduces the
    # whole-frame result bit for bit) and determinism.
    img = O.smooth_u8(1080, 1920, 1, seed=11)
    m = gpu_model("acnet-legacy-hdn0")
    full = session.process_host(m, img, 2.0)
    assert full.shape == (2160, 3840)
    check_int(full, O.oracle_process("acnet-legacy-hdn0", img, 2.0))
    O.set_order(O.ORDER_FMA)
    assert np.array_equal(full, O.oracle_process("acnet-legacy-hdn0", img, 2.0))       # 8.3 M samples, bit for bit
    O.set_order(O.ORDER_GENERIC)
    assert np.array_equal(full, session.process_host(m, img, 2.0))
    halo = 9
    y0, y1, x0, x1 = 300, 420, 700, 860
    crop = img[y0 - halo:y1 + halo, x0 - halo:x1 + halo]
    sub = session.process_host(m, np.ascontiguousarray(crop), 2.0)
    assert np.array_equal(sub[2 * halo:-2 * halo, 2 * halo:-2 * halo], full[2 * y0:2 * y1, 2 * x0:2 * x1])


def test_full_size_1080p_rgb_exact_engine_equals_the_fma_order_oracle(session):
    """BASELINE configs 1 / 2 name a 1920x1080 RGB image: the whole frame through colour split, network (exact engine), Catmull-Rom chroma
    and merge equals the FMA-order oracle bit for bit (24.9 M samples); the default engine (fused colour path) is within the 8-bit bar."""
    img = O.smooth_u8(1080, 1920, 3, seed=12)
    m = gpu_model("acnet-legacy-hdn0")
    O.set_order(O.ORDER_FMA)
    want = O.oracle_process("acnet-legacy-hdn0", img, 2.0)
    O.set_order(O.ORDER_GENERIC)
    session.set_engine(ENGINE_EXACT)
    got = session.process_host(m, img, 2.0)
    assert got.shape == (2160, 3840, 3) and np.array_equal(got, want)
    session.set_engine(ENGINE_AUTO)
    mx, exact = O.compare_u8(session.process_host(m, img, 2.0), want)
    assert mx <= 1 and exact >= 0.999, (mx, exact)


FUSED_SHAPES = [(96, 128), (150, 96), (57, 71), (5, 7), (1, 9), (9, 1), (2, 2), (3, 200), (200, 3), (40, 301), (131, 260)]


@pytest.mark.parametrize("name", ["acnet-legacy-hdn0", "acnet-legacy-hdn2", "acnet-f8b8-hdn", "acnet-f8b8-gan", "acnet-f8b18", "arnet-f8b8", "arnet-f8b16-hdn"])
def test_fused_colour_path_is_bit_identical_to_the_separate_kernels(session, name):
    """Colour split inside the first segment's tile load, chroma resize + merge inside the last segment's tail (TmParams): the RGB
    result must equal the three-extra-kernel path bit for bit on every shape -- tile seams, image borders (folded Catmull-Rom taps),
    images smaller than a strip -- and take exactly one launch per segment."""
    m = gpu_model(name)
    session.set_engine(ENGINE_AUTO)
    for i, (h, w) in enumerate(FUSED_SHAPES):
        img = O.noise_u8(h, w, 3, seed=500 + i)
        session.set_fusion(False)
        n0 = A.launch_count()
        want = session.process_host(m, img, 2.0)
        n1 = A.launch_count()
        session.set_fusion(True)
        got = session.process_host(m, img, 2.0)
        n2 = A.launch_count()
        assert np.array_equal(got, want), (name, h, w, int((got != want).sum()))
        assert (n1 - n0) - (n2 - n1) == 2, (name, n1 - n0, n2 - n1)       # rgb2yuv and chroma_merge are gone
    session.set_fusion(True)


@pytest.mark.parametrize("name", ["acnet-legacy-hdn0", "acnet-f8b8-hdn", "arnet-f8b8"])
def test_fused_colour_path_rgba_is_bit_identical_to_the_separate_kernels_and_the_oracle(session, name):
    """RGBA (rgba2yuva / yuva2rgba, ImageProcess.cpp:113-138, 275-308): colour premultiplied by alpha in the first segment's tile load,
    the (u, v, a) plane resized and merged -- with the un-premultiply division -- in the last segment's tail.  Bit-identical to the
    separate kernels on every shape (transparent and opaque pixels included), two launches less, and within the 8-bit bar of the oracle."""
    m = gpu_model(name)
    session.set_engine(ENGINE_AUTO)
    for i, (h, w) in enumerate(FUSED_SHAPES + [(270, 480)]):
        img = O.noise_u8(h, w, 4, seed=700 + i)
        img[::3, ::5, 3] = 0            # fully transparent pixels: the merge's alpha <= 1e-6 branch
        img[1::4, 2::3, 3] = 255
        session.set_fusion(False)
        n0 = A.launch_count()
        want = session.process_host(m, img, 2.0)
        n1 = A.launch_count()
        session.set_fusion(2)           # RGBA fusion is opt-in (bit-identical, but slower than the separate kernels)
        got = session.process_host(m, img, 2.0)
        n2 = A.launch_count()
        assert np.array_equal(got, want), (name, h, w, int((got != want).sum()))
        assert (n1 - n0) - (n2 - n1) == 2, (name, n1 - n0, n2 - n1)
    session.set_fusion(2)
    # against the oracle: the tensor engine's 1-LSB luma differences are amplified by the division on nearly transparent pixels, so the
    # 1-LSB bar is held where alpha is opaque and the identical-sample bar everywhere
    img = O.noise_u8(64, 80, 4, seed=42)
    img[:, 40:, 3] = 255
    got, want = session.process_host(m, img, 2.0), O.oracle_process(name, img, 2.0)
    mx, exact = O.compare_u8(got, want)
    assert exact >= EXACT_MIN, (name, mx, exact)
    mxo, _ = O.compare_u8(np.ascontiguousarray(got[:, 84:]), np.ascontiguousarray(want[:, 84:]))
    assert mxo <= LSB_MAX, (name, mxo)
    session.set_fusion(True)


def test_fused_colour_path_1080p_and_strided_device_buffers(session):
    import torch
    m = gpu_model("acnet-legacy-hdn0")
    session.set_engine(ENGINE_AUTO)
    img = O.noise_u8(1080, 1920, 3, seed=77)
    session.set_fusion(False)
    want = session.process_host(m, img, 2.0)
    session.set_fusion(True)
    n0 = A.launch_count()
    got = session.process_host(m, img, 2.0)
    assert A.launch_count() - n0 == 2                       # two launches for the whole RGB frame
    assert np.array_equal(got, want)
    # device-resident, padded pitches (odd destination pitches fall back to the separate kernels: the fused stores are 16-bit)
    srcbuf = torch.zeros((1080, 1920 * 3 + 64), dtype=torch.uint8, device="cuda")
    srcbuf[:, :1920 * 3] = torch.from_numpy(img.reshape(1080, -1)).cuda()
    src = srcbuf[:, :1920 * 3].unflatten(1, (1920, 3))
    for pad in (0, 32, 7):
        dstbuf = torch.zeros((2160, 3840 * 3 + pad), dtype=torch.uint8, device="cuda")
        session.process_device(m, src, 2.0, out=dstbuf[:, :3840 * 3].unflatten(1, (3840, 3)))
        session.sync()
        assert np.array_equal(dstbuf[:, :3840 * 3].cpu().numpy().reshape(2160, 3840, 3), want), pad
        assert int(dstbuf[:, 3840 * 3:].sum()) == 0         # nothing written past the row


@pytest.mark.parametrize("w,h,batch,threads", [(720, 480, 600, 1), (720, 480, 600, 8), (1920, 1080, 120, 8)])
def test_reference_benchmark_tool_runs_unchanged_against_the_drop_in(w, h, batch, threads):
    """tools/benchmark/src/Benchmark.cpp of the reference, compiled unchanged (oracle/Makefile ref_callers), at its default 720x480 gray
    shape and at 1080p: creates the "cuda" processor through ac::core::Processor::create, warms up, runs `batch` images over a thread
    pool sharing the processor, prints FPS."""
    import re
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "ac_benchmark")
    if not os.path.isfile(exe):
        pytest.skip("oracle/_ref/ac_benchmark not built (needs /root/reference at build time)")
    out = subprocess.run([exe, "acnet-legacy-hdn0", "cuda", "0", str(w), str(h), str(batch), str(threads)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"FPS: ([0-9.]+)", out.stdout)
    assert m and "processor: CUDA" in out.stdout, out.stdout
    print("reference benchmark tool %dx%d batch %d threads %d: %s FPS" % (w, h, batch, threads, m.group(1)))
    assert float(m.group(1)) > 50.0


@pytest.mark.parametrize("name", ARNETS)
def test_arnet_on_the_tmem_engine_vs_exact_engine_and_mma_engine(session, name):
    """ARNet on the TMEM-resident engine (residual through the per-lane store in shared memory, 1x1 + long skip in the tail's epilogue):
    within the 8-bit bar of the exact engine on odd shapes that cross tile seams and image borders, as close to it as the mma.sync
    implementation is, and identical from run to run."""
    m = gpu_model(name)
    for h, w in ((150, 96), (57, 131), (9, 200), (64, 3)):
        img = O.noise_u8(h, w, 1, seed=900 + h)
        session.set_engine(ENGINE_EXACT)
        want = session.process_host(m, img, 2.0)
        session.set_engine(ENGINE_TENSOR)
        got = {}
        for impl in (0, 2):
            session.set_tensor_impl(impl)
            got[impl] = session.process_host(m, img, 2.0)
            mx, exact = O.compare_u8(got[impl], want)
            assert mx <= LSB_MAX and exact >= EXACT_MIN, (name, h, w, impl, mx, exact)
        assert np.array_equal(got[2], session.process_host(m, img, 2.0))
    session.set_tensor_impl(TENSOR_IMPL)


def test_device_resident_path_matches_host_path(session):
    import torch
    img = O.noise_u8(120, 200, 3, seed=21)
    m = gpu_model("acnet-f8b8-hdn")
    want = session.process_host(m, img, 2.0)
    d = torch.from_numpy(img).cuda()
    out = session.process_device(m, d, 2.0)
    session.sync()
    assert np.array_equal(out.cpu().numpy(), want)


def test_processor_api_threads_and_errors():
    # one processor shared by concurrent callers (tools/benchmark/src/Benchmark.cpp:58-62): per-thread sessions
    import threading
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "anime4kcpp_b200"))
    import pyac
    p = pyac.core.Processor("cuda", 0, "acnet-legacy-hdn0")
    assert p.ok() and p.error() == "NO ERROR" and "NVIDIA" in p.name()
    imgs = [O.noise_u8(64, 80, 1, seed=s) for s in range(6)]
    outs = [None] * 6

    def work(i):
        outs[i] = p(imgs[i], 2.0)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for i in range(6):
        check_int(outs[i], O.oracle_process("acnet-legacy-hdn0", imgs[i], 2.0))
    p.process(imgs[0], 0.5)                     # factors below 1 would need a down-scale by less than 1/2: reported, not aborted
    assert not p.ok() and "at least 1" in p.error()
    out3 = p.process(imgs[0], 3.0)              # non power-of-two: two 2x passes, then the Catmull-Rom luma down-scale by 0.75
    assert p.ok() and out3.shape == (192, 240)
    out4 = p.process(np.ascontiguousarray(imgs[0][:16, :16]), 4.0)      # reference ProcessorTest.cpp:93-104: 4x dims
    assert out4.shape == (64, 64)
    auto = pyac.core.Processor("auto", -1, "acnet-hdn")
    assert auto.ok()
    rgb = O.noise_u8(32, 32, 3, seed=2)
    check_int(auto(rgb), O.oracle_process("acnet-f8b8-hdn", rgb, 2.0))
    assert pyac.core.Processor.InfoList[1].startswith("CUDA:\n  [0] ")


# ---- multi-GPU sharding primitives on one device: bands and the ordered frame stream -----------------------------------------
@pytest.mark.parametrize("name,factor,c", [("acnet-legacy-hdn0", 2.0, 1), ("acnet-legacy-hdn0", 4.0, 3), ("acnet-f8b8-hdn", 2.0, 3), ("arnet-f8b8", 2.0, 1)])
def test_row_bands_reproduce_the_whole_image_bit_for_bit(session, name, factor, c):
    img = O.noise_u8(150, 96, c, seed=31)
    m = gpu_model(name)
    # Bit for bit on the exact engine and on the mma.sync tensor implementation (a pixel's sum order does not depend on where its
    # tile lies).  The TMEM-resident implementation accumulates a row's three input rows in the issue order of its chunks, which
    # depends on the row's position in the tile grid, and a band has its own tile grid: the last bit of an fp32 sum can differ, so
    # there the bands are held to the engine's own bar against the whole image (<= 1 LSB, >= 99.9 % identical).
    for engine, impl in ((ENGINE_EXACT, TENSOR_IMPL), (ENGINE_AUTO, 0), (ENGINE_AUTO, TENSOR_IMPL)):
        session.set_engine(engine)
        session.set_tensor_impl(impl)
        whole = session.process_host(m, img, factor)
        for n_bands in (2, 3, 8):
            out = np.zeros_like(whole)
            for b in range(n_bands):
                A.process_band(session, m, img, factor, n_bands, b, out)
            if engine == ENGINE_EXACT or impl == 0:
                assert np.array_equal(out, whole), (engine, impl, n_bands)
            else:
                mx, exact = O.compare_u8(out, whole)
                assert mx <= 1 and exact >= 0.999, (engine, impl, n_bands, mx, exact)
    session.set_tensor_impl(TENSOR_IMPL)
    # the multi-GPU shape: every band written into a buffer that holds just its own rows (process_band(out_y0=)), then concatenated
    session.set_engine(ENGINE_EXACT)
    whole = session.process_host(m, img, factor)
    parts = []
    for b in range(3):
        _, _, oy0, oy1 = A.band_plan(img.shape[0], factor, m.halo(), 3, b)
        part = np.zeros((oy1 - oy0,) + whole.shape[1:], np.uint8)
        A.process_band(session, m, img, factor, 3, b, part, out_y0=oy0)
        parts.append(part)
    assert np.array_equal(np.concatenate(parts, axis=0), whole)


def test_frame_stream_delivers_in_order_and_matches_single_calls(session):
    m = gpu_model("acnet-legacy-hdn0")
    frames = [O.noise_u8(72, 120, 3, seed=100 + i) for i in range(12)]
    outs = [np.zeros((144, 240, 3), np.uint8) for _ in frames]
    session.set_tensor_impl(TENSOR_IMPL)   # the stream's own sessions run the library default implementation
    stream = A.FrameStream(m, [0, 0], workers_per_device=2, queue_depth=2)      # two lanes on the one device of the test box
    got = []
    for i, (f, o) in enumerate(zip(frames, outs)):
        stream.submit(f, 2.0, o)
        if i >= 4:
            got.append(stream.next()[0])
    while len(got) < len(frames):
        got.append(stream.next()[0])
    stream.close()
    assert got == list(range(len(frames)))
    session.set_engine(ENGINE_AUTO)
    for f, o in zip(frames, outs):
        assert np.array_equal(o, session.process_host(m, f, 2.0))


@pytest.mark.parametrize("name", ["acnet-legacy-hdn0", "acnet-f8b8-hdn"])
def test_default_tensor_engine_is_deterministic(session, name):
    """The TMEM-resident engine issues MMAs from four warps; the accumulation order into a row must not depend on their timing
    (acb200_tm.cuh, the issuers' tickets): the same input gives the same bits on every run and on every session."""
    m = gpu_model(name)
    session.set_engine(ENGINE_TENSOR)
    session.set_tensor_impl(2)
    other = A.Session(0)
    other.set_engine(ENGINE_TENSOR)
    for shape in ((72, 120), (300, 500), (1080, 1920)):
        img = O.noise_u8(shape[0], shape[1], 1, seed=shape[1])
        first = session.process_host(m, img, 2.0)
        for _ in range(6):
            assert np.array_equal(session.process_host(m, img, 2.0), first), shape
        assert np.array_equal(other.process_host(m, img, 2.0), first), shape


# ---- planar / semi-planar video frames (cli/src/Main.cpp:183-206) -----------------------------
def _yuv_frame(h, w, layout, dtype, bits, seed):
    """Synthetic frame: luma (h,w) + chroma planes for I420 / I444 / NV12, samples LSB-aligned in `bits` bits."""
    rs = np.random.RandomState(seed)
    top = 1 << bits
    y = (O.smooth_u8(h, w, 1, seed).astype(np.uint32) * top // 256 + rs.randint(0, max(top // 64, 2), (h, w))).clip(0, top - 1).astype(dtype)
    ch, cw = (h, w) if layout == "i444" else (h // 2, w // 2)
    u = rs.randint(0, top, (ch, cw)).astype(dtype)
    v = (O.smooth_u8(ch, cw, 1, seed + 1).astype(np.uint32) * top // 256).astype(dtype)
    if layout == "nv12":
        return [y, np.ascontiguousarray(np.stack([u, v], axis=-1))]
    if layout == "gray":
        return [y]
    return [y, u, v]


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["i420", "i444", "nv12", "gray"])
@pytest.mark.parametrize("name", ["acnet-legacy-hdn0", "acnet-f8b8-hdn", "arnet-f8b8"])
def test_video_frame_u8_bit_identical_to_oracle(session, layout, name):
    planes = _yuv_frame(54, 98, layout, np.uint8, 8, seed=5)
    O.set_order(O.ORDER_FMA)
    want = O.oracle_frame(name, planes, 2.0)
    got = session.process_frame(gpu_model(name), planes, 2.0)
    assert len(got) == len(want)
    for g, w_ in zip(got, want):
        assert g.shape == w_.shape and np.array_equal(g, w_)


@pytest.mark.gpu
@pytest.mark.parametrize("bits,shift", [(10, 6), (12, 4), (16, 0)])
@pytest.mark.parametrize("layout", ["i420", "nv12"])
def test_video_frame_high_bit_depth_shift_normalisation(session, bits, shift, layout):
    """10/12-bit samples LSB-aligned in 16-bit words: shl before the network, shr after (cli/src/Main.cpp:175,188,195)."""
    planes = _yuv_frame(40, 66, layout, np.uint16, bits, seed=9)
    O.set_order(O.ORDER_FMA)
    want = O.oracle_frame("acnet-legacy-hdn0", planes, 2.0, shift)
    src_copy = [p.copy() for p in planes]
    got = session.process_frame(gpu_model("acnet-legacy-hdn0"), planes, 2.0, shift)
    for g, w_ in zip(got, want):
        assert np.array_equal(g, w_)
    assert int(got[0].max()) < (1 << bits)
    for p, q in zip(planes, src_copy):
        assert np.array_equal(p, q)          # the decoded source frame is not modified


@pytest.mark.gpu
def test_video_frame_4x_tensor_engine_and_strided_planes(session):
    planes = _yuv_frame(46, 70, "i420", np.uint8, 8, seed=21)
    O.set_order(O.ORDER_FMA)
    want = O.oracle_frame("acnet-legacy-hdn1", planes, 4.0)
    m = gpu_model("acnet-legacy-hdn1")
    # padded source rows (a decoder's linesize) and a padded destination luma plane
    padded = [np.zeros((p.shape[0], p.shape[1] + 13), p.dtype) for p in planes]
    views = []
    for p, q in zip(planes, padded):
        q[:, :p.shape[1]] = p
        views.append(q[:, :p.shape[1]])
    out = A.frame_result_planes(planes, 4.0)
    big = np.zeros((out[0].shape[0], out[0].shape[1] + 32), np.uint8)
    out[0] = big[:, :out[0].shape[1]]
    got = session.process_frame(m, views, 4.0, out=out)
    for g, w_ in zip(got, want):
        assert np.array_equal(g, w_)
    assert not big[:, out[0].shape[1]:].any()
    for impl in TENSOR_IMPLS:
        session.set_engine(ENGINE_AUTO)
        session.set_tensor_impl(impl)
        got = session.process_frame(m, planes, 4.0)
        mx, same = O.compare_u8(got[0], want[0])
        assert mx <= 1 and same >= 0.999, (impl, mx, same)
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])


@pytest.mark.gpu
def test_video_frame_1080p_i420_properties(session):
    """Full-size frame: luma equals the gray-image path bit for bit; chroma equals the stand-alone resize."""
    planes = _yuv_frame(1080, 1920, "i420", np.uint8, 8, seed=33)
    m = gpu_model("acnet-legacy-hdn0")
    session.set_engine(ENGINE_AUTO)
    got = session.process_frame(m, planes, 2.0)
    assert np.array_equal(got[0], session.process_host(m, planes[0], 2.0))
    for i in (1, 2):
        assert np.array_equal(got[i], session.resize_catmull_rom(planes[i], 1920, 1080))


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [3, 5, 6])
def test_video_frame_chroma_planes_at_integer_scales_with_zero_taps(session, scale):
    """Chroma destination planes need not be 2^k x the source: 3x / 5x / 6x have phases whose outer Catmull-Rom taps are exactly
    zero (the tiled kernel's source window must not depend on them)."""
    planes = _yuv_frame(40, 140, "i444", np.uint8, 8, seed=15)
    out = [np.empty((80, 280), np.uint8), np.empty((40 * scale, 140 * scale), np.uint8), np.empty((40 * scale, 140 * scale), np.uint8)]
    got = session.process_frame(gpu_model("acnet-legacy-hdn0"), planes, 2.0, out=out)
    for i in (1, 2):
        assert np.array_equal(got[i], O.oracle_resize(planes[i], 140 * scale, 40 * scale))


@pytest.mark.gpu
def test_video_frame_errors(session):
    m = gpu_model("acnet-legacy-hdn0")
    planes = _yuv_frame(16, 16, "i420", np.uint8, 8, seed=1)
    with pytest.raises(A.Acb200Error):
        session.process_frame(m, planes, 0.5)                                   # factors below 1 are not provided
    with pytest.raises(A.Acb200Error):
        session.process_frame(m, [planes[0]] * 4, 2.0)                          # too many planes
    bad = A.frame_result_planes(planes, 2.0)
    bad[0] = np.zeros((31, 32), np.uint8)
    with pytest.raises(A.Acb200Error):
        session.process_frame(m, planes, 2.0, out=bad)                          # luma destination of the wrong size
    with pytest.raises(A.Acb200Error):
        session.process_frame(m, planes, 2.0, shift=8)                          # shift outside the element width
    # the session still works after the errors
    assert session.process_frame(m, planes, 2.0)[0].shape == (32, 32)


@pytest.mark.gpu
def test_frame_stream_planar_frames_in_order(session):
    m = gpu_model("acnet-legacy-hdn0")
    frames = [_yuv_frame(36, 64, "i420" if i % 2 else "nv12", np.uint8, 8, seed=200 + i) for i in range(10)]
    outs = [A.frame_result_planes(f, 2.0) for f in frames]
    stream = A.FrameStream(m, [0, 0], workers_per_device=2, queue_depth=2)
    got = []
    for i, (f, o) in enumerate(zip(frames, outs)):
        stream.submit_frame(f, 2.0, o)
        if i >= 3:
            got.append(stream.next()[0])
    while len(got) < len(frames):
        got.append(stream.next()[0])
    stream.close()
    assert got == list(range(len(frames)))
    session.set_engine(ENGINE_AUTO)
    for f, o in zip(frames, outs):
        for a, b in zip(o, session.process_frame(m, f, 2.0)):
            assert np.array_equal(a, b)


@pytest.mark.gpu
def test_c_binding_process_frame_extension(session):
    """ac_processor_process_frame (include/AC/Core/Processor.h): the reference's per-frame video callback as one C call."""
    import ctypes as C
    L = A.lib()

    class ACProcessor(C.Structure):
        _fields_ = [("device", C.c_int), ("type", C.c_char_p), ("model", C.c_char_p), ("hptr", C.c_void_p)]
    L.ac_processor_alloc.restype = C.POINTER(ACProcessor)
    L.ac_processor_error.restype = C.c_char_p
    L.ac_processor_process_frame.argtypes = [C.POINTER(ACProcessor), C.POINTER(A.Plane), C.POINTER(A.Plane), C.c_int, C.c_int, C.c_int, C.c_double]
    p = L.ac_processor_alloc()
    p.contents.type, p.contents.model, p.contents.device = b"cuda", b"acnet-legacy-hdn0", 0
    assert L.ac_processor_create(p) == 0
    planes = _yuv_frame(30, 44, "i420", np.uint16, 10, seed=77)
    out = A.frame_result_planes(planes, 2.0)
    src, dst = A._planes_of(planes), A._planes_of(out)
    assert L.ac_processor_process_frame(p, src, dst, 3, 0x002, 6, 2.0) == 0
    session.set_engine(ENGINE_AUTO)
    for a, b in zip(out, session.process_frame(gpu_model("acnet-legacy-hdn0"), planes, 2.0, 6)):
        assert np.array_equal(a, b)
    # failure is reported through the processor's sticky status, like ac_processor_process
    assert L.ac_processor_process_frame(p, src, dst, 3, 0x002, 6, 3.0) == -256
    assert b"factor x the source" in L.ac_processor_error(p)        # the destination planes were sized for 2x
    assert L.ac_processor_process_frame(p, src, dst, 3, 0x002, 6, 2.0) == 0
    assert L.ac_processor_process_frame(None, src, dst, 3, 0x002, 6, 2.0) == -22
    L.ac_processor_free(C.byref(p))


# ---- ArtCNN<16/32>, FSRCNNX<8/16>: per-layer fp32 kernels in the reference FMA-backend order -----------
WIDE = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wide_vectors.npz"))
WIDE_MODELS = ["artcnn-c4f16", "artcnn-c4f32-ds", "fsrcnnx-f8b4", "fsrcnnx-f16b4-distort-plus"]


def wide_src(kind):
    g = WIDE["in_gray_noise"]
    return {"gray_noise_2x": g, "gray_smooth_2x": WIDE["in_gray_smooth"], "rgb_2x": WIDE["in_rgb"], "gray_4x": np.ascontiguousarray(g[:16, :20]),
            "gray_f32_2x": g.astype(np.float32) / np.float32(255), "gray_u16_2x": g.astype(np.uint16) * 257}[kind]


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(k.split(":", 1)[1] for k in WIDE.files if k.startswith("fma:")))
def test_wide_families_bit_identical_to_reference_fma_backend_and_close_to_the_others(session, key):
    kind, name = key.split("/", 1)
    out = session.process_host(gpu_model(name), wide_src(kind), 4.0 if kind.endswith("4x") else 2.0)
    assert np.array_equal(out, WIDE["fma:" + key])
    check_close(out, WIDE["generic:" + key], x4=kind.endswith("4x"))
    if "avx512:" + key in WIDE.files:
        check_close(out, WIDE["avx512:" + key])


@pytest.mark.gpu
@pytest.mark.parametrize("name", WIDE_MODELS)
@pytest.mark.parametrize("shape", [(3, 3), (1, 9), (7, 1), (33, 65), (64, 32), (70, 97)])
def test_wide_families_odd_sizes_vs_oracle(session, name, shape):
    O.set_order(O.ORDER_FMA)
    img = O.noise_u8(shape[0], shape[1], 1, seed=shape[0] * 131 + shape[1])
    assert np.array_equal(session.process_host(gpu_model(name), img, 2.0), O.oracle_process(name, img, 2.0))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["artcnn-c4f16-dn", "fsrcnnx-f16b4"])
def test_wide_families_colour_4x_bands_frames_and_processor_api(session, name):
    O.set_order(O.ORDER_FMA)
    m = gpu_model(name)
    rgba = O.noise_u8(26, 38, 4, seed=3)
    got, want = session.process_host(m, rgba, 4.0), O.oracle_process(name, rgba, 4.0)
    opaque = np.repeat((want[..., 3:4] >= 128), 4, axis=2)          # un-premultiply amplifies 1-LSB chroma noise where alpha is small
    assert np.array_equal(got[opaque], want[opaque])
    # row bands reproduce the whole image (halo = blocks + 3 layers, 5x5 head included)
    gray = O.noise_u8(90, 50, 1, seed=8)
    whole = session.process_host(m, gray, 2.0)
    out = np.zeros_like(whole)
    for b in range(3):
        A.process_band(session, m, gray, 2.0, 3, b, out)
    assert np.array_equal(out, whole)
    # planar video frame and the Processor front door
    planes = _yuv_frame(40, 56, "i420", np.uint8, 8, seed=4)
    for a, b in zip(session.process_frame(m, planes, 2.0), O.oracle_frame(name, planes, 2.0)):
        assert np.array_equal(a, b)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "anime4kcpp_b200"))
    import pyac
    p = pyac.core.Processor("cuda", 0, name)
    assert p.ok(), p.error()
    check_int(p(gray, 2.0), whole)          # the Processor runs the default engine (tensor cores for F >= 16): tolerance, not identity


@pytest.mark.gpu
def test_wide_families_1080p_frame_against_oracle_sample(session):
    """Full-size frame: the top-left 160x160 of the result equals the oracle on a 96x96 crop's interior-independent region."""
    O.set_order(O.ORDER_FMA)
    img = O.smooth_u8(1080, 1920, 1, seed=2)
    for name in ("artcnn-c4f16", "fsrcnnx-f8b4"):
        got = session.process_host(gpu_model(name), img, 2.0)
        assert got.shape == (2160, 3840)
        want = O.oracle_process(name, np.ascontiguousarray(img[:96, :96]), 2.0)
        assert np.array_equal(got[:160, :160], want[:160, :160])     # 96 - 7 layers of context - margin = 80 source pixels


@pytest.mark.gpu
def test_config5_8192_square_factor4_in_eight_bands(session):
    """SURVEY.md 8d config 5 at full size: one 8192x8192 1-channel u8 image, factor 4 (two 2x passes), 8 row bands.  The bands
    equal the whole-image result bit for bit, and halo'd crops equal the oracle (which cannot take the whole image)."""
    name = "acnet-legacy-hdn0"
    m = gpu_model(name)
    rs = np.random.RandomState(7)
    base = O.smooth_u8(1024, 1024, 1, seed=5)
    img = np.tile(base, (8, 8))
    img[::7, ::5] = rs.randint(0, 256, img[::7, ::5].shape, dtype=np.uint8)     # break the periodicity
    session.set_engine(ENGINE_EXACT)
    whole = session.process_host(m, img, 4.0)
    assert whole.shape == (32768, 32768)
    out = np.zeros_like(whole)
    for b in range(8):
        A.process_band(session, m, img, 4.0, 8, b, out)
    assert np.array_equal(out, whole)
    del out
    # crops with full context: 2 passes x 9 layers -> 9 + 5 source pixels of halo; compare the interior
    O.set_order(O.ORDER_FMA)
    for (y0, x0) in ((0, 0), (4000, 4100), (8192 - 96, 8192 - 96), (1024 * 3 - 40, 17)):
        crop = np.ascontiguousarray(img[y0:y0 + 96, x0:x0 + 96])
        want = O.oracle_process(name, crop, 4.0)
        mt = 0 if y0 == 0 else 16
        ml = 0 if x0 == 0 else 16
        mb = 0 if y0 + 96 == 8192 else 16
        mr = 0 if x0 + 96 == 8192 else 16
        got = whole[4 * (y0 + mt):4 * (y0 + 96 - mb), 4 * (x0 + ml):4 * (x0 + 96 - mr)]
        assert np.array_equal(got, want[4 * mt:4 * (96 - mb), 4 * ml:4 * (96 - mr)]), (y0, x0)


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(k.split(":", 1)[1] for k in WIDE.files if k.startswith("fma:") and ("f16" in k or "f32" in k)))
def test_wide_families_tensor_engine_golden(session, key):
    """tcgen05 split-fp16 engine of the F -> F layers (F = 16 / 32): the north_star tolerance against every reference backend."""
    kind, name = key.split("/", 1)
    session.set_engine(ENGINE_TENSOR)
    out = session.process_host(gpu_model(name), wide_src(kind), 4.0 if kind.endswith("4x") else 2.0)
    for order in ("fma:", "generic:", "avx512:"):
        if order + key in WIDE.files:
            check_close(out, WIDE[order + key], x4=kind.endswith("4x"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["artcnn-c4f16", "artcnn-c4f32-dn", "fsrcnnx-f16b4-distort-plus"])
@pytest.mark.parametrize("shape", [(3, 3), (1, 9), (15, 64), (16, 65), (33, 130), (70, 97)])
def test_wide_families_tensor_engine_odd_sizes(session, name, shape):
    O.set_order(O.ORDER_FMA)
    img = O.noise_u8(shape[0], shape[1], 1, seed=shape[0] * 7 + shape[1])
    session.set_engine(ENGINE_TENSOR)
    mx, same = O.compare_u8(session.process_host(gpu_model(name), img, 2.0), O.oracle_process(name, img, 2.0))
    assert mx <= 1 and same >= 0.998, (mx, same)        # tiny images: one differing sample is already 0.1-1 %


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["artcnn-c4f16", "artcnn-c4f32", "fsrcnnx-f16b4"])
def test_wide_families_tensor_engine_1080p_against_exact_engine(session, name):
    img = O.smooth_u8(1080, 1920, 1, seed=9)
    m = gpu_model(name)
    session.set_engine(ENGINE_EXACT)
    exact = session.process_host(m, img, 2.0)
    session.set_engine(ENGINE_AUTO)                 # the default: tensor engine on the last (here: only) pass
    got = session.process_host(m, img, 2.0)
    mx, same = O.compare_u8(got, exact)
    assert mx <= LSB_MAX and same >= EXACT_MIN, (mx, same)
    f = (img[:256, :256] / 255.0).astype(np.float32)
    session.set_engine(ENGINE_EXACT)
    e32 = session.process_host(m, f, 2.0)
    session.set_engine(ENGINE_TENSOR)
    assert float(np.abs(session.process_host(m, f, 2.0) - e32).max()) <= F32_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["acnet-legacy-hdn0", "acnet-f8b4", "arnet-f8b8", "artcnn-c4f16", "fsrcnnx-f8b4"])
def test_reference_processor_test_criterion_psnr_48db(session, name):
    """The reference's own acceptance test (tests/core/src/ProcessorTest.cpp:113-217): a 64x64 RGB noise image, the processor
    under test against the Generic CPU backend, PSNR > 48 dB -- here through the Processor front door with its default engine."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "anime4kcpp_b200"))
    import pyac
    src = O.noise_u8(64, 64, 3, seed=20)
    O.set_order(O.ORDER_GENERIC)
    ref = O.oracle_process(name, src, 2.0).astype(np.float64)
    p = pyac.core.Processor("cuda", 0, name)
    assert p.ok(), p.error()
    dst = p.process(src, 2.0).astype(np.float64)
    assert p.ok()
    mse = float(((dst - ref) ** 2).mean())
    psnr = float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)
    assert psnr > 48.0, psnr



# ---- factors that are not powers of two (Processor.cpp:203-204, 237, 249): passes up to the next power of two, then the
#      Catmull-Rom luma down-scale by fxy = factor / 2^power and the chroma resize by the full factor ---------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("factor", [1.0, 1.5, 2.5, 3.0, 3.7])
@pytest.mark.parametrize("name", ["acnet-legacy-hdn0", "acnet-f8b8-hdn", "artcnn-c4f16"])
def test_non_power_of_two_factors_gray_and_colour_vs_oracle(session, name, factor):
    O.set_order(O.ORDER_FMA)
    m = gpu_model(name)
    gray = O.smooth_u8(37, 53, 1, seed=3)
    want = O.oracle_process(name, gray, factor)
    got = session.process_host(m, gray, factor)
    assert got.shape == want.shape == (int(37 * factor), int(53 * factor))
    assert np.array_equal(got, want)
    rgb = O.smooth_u8(26, 34, 3, seed=4)
    assert np.array_equal(session.process_host(m, rgb, factor), O.oracle_process(name, rgb, factor))


@pytest.mark.gpu
def test_non_power_of_two_factor_types_engines_and_device_path(session):
    import torch
    O.set_order(O.ORDER_FMA)
    name = "acnet-f8b4"
    m = gpu_model(name)
    g8 = O.smooth_u8(40, 48, 1, seed=6)
    for img in (g8.astype(np.uint16) * 257, (g8 / 255.0).astype(np.float32), O.noise_u8(22, 30, 4, seed=7)):
        want = O.oracle_process(name, img, 1.5)
        got = session.process_host(m, img, 1.5)
        if img.ndim == 3:
            opaque = np.repeat((want[..., 3:4] >= 128), 4, axis=2)
            assert np.array_equal(got[opaque], want[opaque])
        else:
            assert np.array_equal(got, want)
    # default engine (tensor cores on the last pass): the north_star tolerance
    session.set_engine(ENGINE_AUTO)
    rgb = O.smooth_u8(64, 80, 3, seed=8)
    check_close(session.process_host(m, rgb, 3.0), O.oracle_process(name, rgb, 3.0), x4=True)
    # device-resident entry
    session.set_engine(ENGINE_EXACT)
    d = torch.from_numpy(rgb).cuda()
    out = session.process_device(m, d, 1.5)
    torch.cuda.synchronize()
    session.sync()
    assert np.array_equal(out.cpu().numpy(), O.oracle_process(name, rgb, 1.5))
    with pytest.raises(A.Acb200Error):
        session.process_host(m, g8, 0.75)


@pytest.mark.gpu
@pytest.mark.parametrize("factor,bits,shift", [(1.5, 8, 0), (3.0, 10, 6)])
def test_video_frame_non_power_of_two_factor(session, factor, bits, shift):
    """processor->process(srcy, dsty, factor) with any factor >= 1 (cli/src/Main.cpp:188): luma through the passes and the luma
    down-scale, chroma straight to factor x its size."""
    O.set_order(O.ORDER_FMA)
    planes = _yuv_frame(40, 56, "i420", np.uint8 if bits == 8 else np.uint16, bits, seed=31)
    want = O.oracle_frame("acnet-legacy-hdn0", planes, factor, shift)
    got = session.process_frame(gpu_model("acnet-legacy-hdn0"), planes, factor, shift)
    assert got[0].shape == (int(40 * factor), int(56 * factor)) and got[1].shape == (int(20 * factor), int(28 * factor))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("out_shape", [(20, 30), (23, 37), (29, 50), (15, 22), (45, 30)])
@pytest.mark.parametrize("c", [1, 3])
def test_catmull_rom_downscale_and_mixed_resize_vs_oracle(session, out_shape, c):
    """ac::core::resize with the Catmull-Rom filter below 1x (down to 1/2, what the driver needs for non-2^k factors) and with
    one axis growing while the other shrinks."""
    img = O.noise_u8(30, 44, c, seed=12)
    oh, ow = out_shape
    if ow >= 44 and oh >= 30:
        pytest.skip("pure up-scale is covered elsewhere")
    want = O.oracle_resize(img, ow, oh)
    assert np.array_equal(session.resize_catmull_rom(img, ow, oh), want)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(40))
def test_randomized_model_shape_type_factor_vs_oracle(session, seed):
    """Seeded random draws over the whole input space of Processor::process -- model family, image size (down to one pixel),
    channels, element type, factor (powers of two or not) -- exact engine against the FMA-order oracle, bit for bit."""
    rs = np.random.RandomState(1000 + seed)
    name = str(rs.choice(["acnet-legacy-gan", "acnet-legacy-hdn2", "acnet-f8b4-box", "acnet-f8b8", "acnet-f8b18-hdn", "arnet-f8b8-hdn", "arnet-f8b16-box",
                          "artcnn-c4f16-ds", "artcnn-c4f32", "fsrcnnx-f8b4-distort-plus", "fsrcnnx-f16b4"]))
    h, w = int(rs.randint(1, 90)), int(rs.randint(1, 120))
    c = int(rs.choice([1, 1, 3, 4]))
    factor = float(rs.choice([1.0, 1.25, 1.5, 2.0, 2.0, 2.0, 3.0, 4.0]))
    dtype = [np.uint8, np.uint8, np.uint16, np.float32][int(rs.randint(0, 4))]
    img = O.noise_u8(h, w, c, seed=seed) if rs.rand() < 0.5 else O.smooth_u8(h, w, c, seed=seed)
    if dtype == np.uint16:
        img = img.astype(np.uint16) * 257
    elif dtype == np.float32:
        img = (img / 255.0).astype(np.float32)
    O.set_order(O.ORDER_FMA)
    want = O.oracle_process(name, img, factor)
    got = session.process_host(gpu_model(name), img, factor)
    assert got.shape == want.shape and got.dtype == want.dtype
    if c == 4:
        thr = {np.uint8: 128, np.uint16: 32768, np.float32: 0.5}[dtype]
        opaque = np.repeat(want[..., 3:4] >= thr, 4, axis=2)
        assert np.array_equal(got[opaque], want[opaque]), (name, h, w, c, factor, np.dtype(dtype).name)
    else:
        assert np.array_equal(got, want), (name, h, w, c, factor, np.dtype(dtype).name)


# ---- SURVEY.md 8d configs 2 and 3 at full size: every real-weight 8-feature model and the largest ARNet on a 1080p frame ------------
ALL_REAL_8 = ["acnet-legacy-gan", "acnet-legacy-hdn0", "acnet-legacy-hdn1", "acnet-legacy-hdn2", "acnet-legacy-hdn3",
              "acnet-f8b4", "acnet-f8b4-hdn", "acnet-f8b4-box", "acnet-f8b4-box-hdn", "acnet-f8b8", "acnet-f8b8-hdn", "acnet-f8b8-box",
              "acnet-f8b8-box-hdn", "acnet-f8b18", "acnet-f8b18-hdn", "acnet-f8b18-box", "acnet-f8b18-box-hdn"]
_frame_1080p = {}


def frame_1080p():
    if "img" not in _frame_1080p:
        _frame_1080p["img"] = O.smooth_u8(1080, 1920, 1, seed=77)
    return _frame_1080p["img"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL_REAL_8)
def test_config2_every_real_model_1080p_default_engine_within_the_bar(session, name):
    img = frame_1080p()
    m = gpu_model(name)
    session.set_engine(ENGINE_EXACT)
    exact = session.process_host(m, img, 2.0)
    # the exact engine against the oracle on a crop whose interior does not see the crop border (<= 20 layers of context)
    O.set_order(O.ORDER_FMA)
    want = O.oracle_process(name, np.ascontiguousarray(img[500:596, 900:996]), 2.0)
    assert np.array_equal(exact[2 * 524:2 * 572, 2 * 924:2 * 972], want[48:144, 48:144])
    session.set_engine(ENGINE_AUTO)
    mx, same = O.compare_u8(session.process_host(m, img, 2.0), exact)
    assert mx <= LSB_MAX and same >= EXACT_MIN, (name, mx, same)


@pytest.mark.gpu
def test_config3_arnet_f8b64_small_exact_and_1080p_engines(session):
    """The largest model of the reference's catalogue (130 3x3 layers, synthetic seeded weights: ARNet.p is absent)."""
    name = "arnet-f8b64"
    m = gpu_model(name)
    O.set_order(O.ORDER_FMA)
    small = O.noise_u8(40, 56, 1, seed=64)
    session.set_engine(ENGINE_EXACT)
    assert np.array_equal(session.process_host(m, small, 2.0), O.oracle_process(name, small, 2.0))
    img = frame_1080p()
    exact = session.process_host(m, img, 2.0)
    session.set_engine(ENGINE_AUTO)
    mx, same = O.compare_u8(session.process_host(m, img, 2.0), exact)
    assert mx <= LSB_MAX and same >= EXACT_MIN, (mx, same)
