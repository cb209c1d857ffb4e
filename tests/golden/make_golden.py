"""Mints tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libac_ref.so, Generic backend = arch 1,
the backend the reference's own ProcessorTest.cpp:129 treats as ground truth).  Run in the build container,
where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The GPU box has no /root/reference; tests there read only the committed vectors.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

REAL = ["acnet-legacy-gan", "acnet-legacy-hdn0", "acnet-legacy-hdn1", "acnet-legacy-hdn2", "acnet-legacy-hdn3",
        "acnet-f8b4", "acnet-f8b4-hdn", "acnet-f8b4-box", "acnet-f8b4-box-hdn",
        "acnet-f8b8", "acnet-f8b8-hdn", "acnet-f8b8-box", "acnet-f8b8-box-hdn",
        "acnet-f8b18", "acnet-f8b18-hdn", "acnet-f8b18-box", "acnet-f8b18-box-hdn"]
ARNET = ["arnet-f8b8", "arnet-f8b16-hdn", "arnet-f8b32-box", "arnet-f8b64-box-hdn"]


def main():
    assert O.ref() is not None, "build oracle/_ref first (make -C oracle ref)"
    out = {}
    gray = O.noise_u8(40, 48, 1, seed=1234)
    smooth = O.smooth_u8(36, 52, 1, seed=7)
    out["in_gray_noise"] = gray
    out["in_gray_smooth"] = smooth
    for name in REAL + ARNET:
        out["gray_noise_2x/" + name] = O.ref_process(name, gray, 2.0, arch=1)
    for name in ["acnet-legacy-hdn0", "acnet-f8b8-hdn", "acnet-f8b18", "arnet-f8b8"]:
        out["gray_smooth_2x/" + name] = O.ref_process(name, smooth, 2.0, arch=1)
    rgb = O.noise_u8(24, 32, 3, seed=99)
    rgba = O.noise_u8(20, 28, 4, seed=100)
    out["in_rgb"] = rgb
    out["in_rgba"] = rgba
    for name in ["acnet-legacy-hdn0", "acnet-f8b4", "arnet-f8b8"]:
        # RGB: the chroma resize inside is the oracle restatement (stb is not in the container) -- unpinned stage
        out["rgb_2x/" + name] = O.ref_process(name, rgb, 2.0, arch=1)
        out["rgb_4x/" + name] = O.ref_process(name, rgb, 4.0, arch=1)
        out["rgba_2x/" + name] = O.ref_process(name, rgba, 2.0, arch=1)
        out["gray_4x/" + name] = O.ref_process(name, gray[:20, :24], 4.0, arch=1)
        out["gray_f32_2x/" + name] = O.ref_process(name, gray.astype(np.float32) / np.float32(255), 2.0, arch=1)
        out["gray_u16_2x/" + name] = O.ref_process(name, gray.astype(np.uint16) * 257, 2.0, arch=1)
    # colour conversion alone (reference code, pinned)
    y = np.empty(rgb.shape[:2], np.uint8)
    uv = np.empty(rgb.shape[:2] + (2,), np.uint8)
    O.ref().ref_rgb2yuv(rgb.ctypes.data, 32, 24, 3, rgb.strides[0], O.U8, y.ctypes.data, y.strides[0], uv.ctypes.data, uv.strides[0])
    back = np.empty_like(rgb)
    O.ref().ref_yuv2rgb(y.ctypes.data, y.strides[0], uv.ctypes.data, uv.strides[0], 32, 24, 3, O.U8, back.ctypes.data, back.strides[0])
    out["rgb2yuv_y"], out["rgb2yuv_uv"], out["yuv2rgb_back"] = y, uv, back
    # the same cases on the reference's FMA backend (arch 4 = OpImplX86SIMD256<true>; bit-identical to what the auto-ISA
    # choice, AVX512, executes for these 8-channel models) -- the order the GPU's exact engine reproduces bit for bit
    fma = {}
    for key in list(out.keys()):
        if "/" not in key:
            continue
        kind, name = key.split("/", 1)
        src = {"gray_noise_2x": gray, "gray_smooth_2x": smooth, "rgb_2x": rgb, "rgb_4x": rgb, "rgba_2x": rgba, "gray_4x": gray[:20, :24],
               "gray_f32_2x": gray.astype(np.float32) / np.float32(255), "gray_u16_2x": gray.astype(np.uint16) * 257}[kind]
        fma["fma:" + key] = O.ref_process(name, src, 4.0 if kind.endswith("4x") else 2.0, arch=4)
    out.update(fma)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
