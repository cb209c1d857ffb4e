"""Mints tests/golden/wide_vectors.npz -- ArtCNN<16/32> and FSRCNNX<8/16> -- from the COMPILED REFERENCE
(oracle/_ref/libac_ref.so): its Generic backend (arch 1, the ground truth of the reference's own ProcessorTest.cpp:129), its
256-bit FMA backend (arch 4, the order the GPU kernels reproduce bit for bit) and its AVX512 backend (arch 5: for these
16/32-feature models it sums in a third order, X86/AVX512.hpp:27-52).  Run in the build container:

    make -C oracle ref && python tests/golden/make_golden_wide.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

ALL = ["artcnn-c4f16", "artcnn-c4f16-dn", "artcnn-c4f16-ds", "artcnn-c4f32", "artcnn-c4f32-dn", "artcnn-c4f32-ds",
       "fsrcnnx-f8b4", "fsrcnnx-f8b4-distort-plus", "fsrcnnx-f16b4", "fsrcnnx-f16b4-distort-plus"]
DEEP = ["artcnn-c4f16", "artcnn-c4f32-ds", "fsrcnnx-f8b4", "fsrcnnx-f16b4-distort-plus"]
ARCH = {"generic": 1, "fma": 4, "avx512": 5}


def main():
    assert O.ref() is not None, "build oracle/_ref first (make -C oracle ref)"
    out = {}
    gray = O.noise_u8(34, 46, 1, seed=4321)
    smooth = O.smooth_u8(30, 44, 1, seed=11)
    rgb = O.noise_u8(22, 30, 3, seed=98)
    out["in_gray_noise"], out["in_gray_smooth"], out["in_rgb"] = gray, smooth, rgb
    cases = {"gray_noise_2x": (gray, 2.0), "gray_smooth_2x": (smooth, 2.0), "rgb_2x": (rgb, 2.0), "gray_4x": (np.ascontiguousarray(gray[:16, :20]), 4.0),
             "gray_f32_2x": (gray.astype(np.float32) / np.float32(255), 2.0), "gray_u16_2x": (gray.astype(np.uint16) * 257, 2.0)}
    have512 = O.ref().ref_processor_name(b"artcnn-c4f16", 0) == b"AVX512"
    for name in ALL:
        for kind, (img, factor) in cases.items():
            if kind != "gray_noise_2x" and name not in DEEP:
                continue
            for order, arch in ARCH.items():
                if order == "avx512" and (not have512 or kind != "gray_noise_2x"):
                    continue
                out["%s:%s/%s" % (order, kind, name)] = O.ref_process(name, img, factor, arch=arch)
    path = os.path.join(HERE, "wide_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", len(out), "arrays,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
