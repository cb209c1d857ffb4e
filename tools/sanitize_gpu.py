"""Small end-to-end runs of every engine for compute-sanitizer (memcheck / racecheck / initcheck) under gpurun:
    compute-sanitizer --tool racecheck python tools/sanitize_gpu.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import anime4kcpp_b200 as A
import oracle_lib as O

s = A.Session(0)
O.set_order(O.ORDER_FMA)
# (150 x 150 has interior CTAs -- tiles whose 56 x 56 frame lies inside the image -- which take the unclamped tile loop)
for name, shape, c in (("acnet-legacy-hdn0", (70, 90), 3), ("acnet-legacy-hdn0", (150, 150), 1), ("arnet-f8b8", (140, 150), 1), ("acnet-f8b8-hdn", (50, 64), 1), ("acnet-f8b18", (48, 48), 1), ("arnet-f8b8", (60, 44), 4)):
    img = O.noise_u8(shape[0], shape[1], c, seed=1)
    want = O.oracle_process(name, img, 2.0)
    m = A.Model(name)
    for engine, impl in ((0, 0), (1, 0), (1, 1)):
        s.set_engine(engine)
        s.set_tensor_impl(impl)
        got = s.process_host(m, img, 2.0)
        ok = c == 4 or O.compare_u8(got, want)[0] <= 1
        print(name, "engine", engine, "impl", impl, O.compare_u8(got, want), "ok" if ok else "MISMATCH")
# ArtCNN / FSRCNNX per-layer kernels: exact FFMA and the tcgen05 engine (F >= 16), plus a non-power-of-two factor
for eng in (0, 1):
    s.set_engine(eng)
    for name, shape in (("artcnn-c4f16", (37, 70)), ("artcnn-c4f32", (20, 66)), ("fsrcnnx-f16b4", (34, 50))):
        img = O.noise_u8(shape[0], shape[1], 1, seed=3)
        mx, same = O.compare_u8(s.process_host(A.Model(name), img, 2.0), O.oracle_process(name, img, 2.0))
        print(name, "engine", eng, (mx, same), "ok" if mx <= 1 else "MISMATCH")
s.set_engine(0)
img = O.noise_u8(30, 44, 3, seed=4)
print("factor 1.5:", "identical" if np.array_equal(s.process_host(A.Model("acnet-f8b4"), img, 1.5), O.oracle_process("acnet-f8b4", img, 1.5)) else "MISMATCH")
s.set_engine(0)
for name, shape, c in (("artcnn-c4f16", (41, 70), 1), ("artcnn-c4f32-dn", (35, 33), 3), ("fsrcnnx-f8b4", (66, 37), 1), ("fsrcnnx-f16b4", (34, 50), 1)):
    img = O.noise_u8(shape[0], shape[1], c, seed=2)
    got = s.process_host(A.Model(name), img, 2.0)
    print(name, "identical" if np.array_equal(got, O.oracle_process(name, img, 2.0)) else "MISMATCH")
# planar / semi-planar video frames: shift kernels, the tiled chroma-plane kernel (u8 / u16, 1 and 2 channels), 4x
rs = np.random.RandomState(5)
m = A.Model("acnet-legacy-hdn0")
s.set_engine(2)
s.set_tensor_impl(0)
for dtype, bits, shift, layout, factor in ((np.uint8, 8, 0, "i420", 2.0), (np.uint16, 10, 6, "nv12", 2.0), (np.uint16, 16, 0, "i444", 4.0), (np.float32, 0, 0, "i420", 2.0)):
    h, w = 38, 70
    ch, cw = (h, w) if layout == "i444" else (h // 2, w // 2)
    if dtype == np.float32:
        planes = [rs.rand(h, w).astype(dtype), rs.rand(ch, cw).astype(dtype), rs.rand(ch, cw).astype(dtype)]
    else:
        planes = [rs.randint(0, 1 << bits, (h, w)).astype(dtype), rs.randint(0, 1 << bits, (ch, cw)).astype(dtype), rs.randint(0, 1 << bits, (ch, cw)).astype(dtype)]
    if layout == "nv12":
        planes = [planes[0], np.ascontiguousarray(np.stack(planes[1:], axis=-1))]
    got = s.process_frame(m, planes, factor, shift)
    want = O.oracle_frame("acnet-legacy-hdn0", planes, factor, shift)
    ok = all(np.array_equal(a, b) for a, b in zip(got[1:], want[1:]))
    print("frame", layout, np.dtype(dtype).name, "shift", shift, "factor", factor, "chroma", "ok" if ok else "MISMATCH")
print("done")
