#!/usr/bin/env python
"""Per-frame GPU time of the TMEM engine on device-resident 1080p frames, 32 frames back to back on one stream (CUDA events): the luma
pass alone (gray) and the whole RGB frame with the colour handling fused or as separate kernels; plus a checksum of one result so that
build variants can be compared bit for bit.  GPU box only.

    ACB200_LIB=anime4kcpp_b200/lib/libvariant_X.so python tools/time_tm.py [model ...]
"""
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import anime4kcpp_b200 as A
import oracle_lib as O

models = sys.argv[1:] or ["acnet-legacy-hdn0"]
gray = torch.from_numpy(O.noise_u8(1080, 1920, 1, seed=3)).cuda()
rgb = torch.from_numpy(O.noise_u8(1080, 1920, 3, seed=4)).cuda()
tag = os.path.basename(os.environ.get("ACB200_LIB", "libac_b200.so"))
stream = torch.cuda.Stream()
for name in models:
    m = A.Model(name)
    for label, img, fuse in (("gray", gray, 1), ("rgb fused", rgb, 1), ("rgb separate", rgb, 0)):
        s = A.Session(0)
        s.set_engine(2)
        s.set_fusion(fuse)
        out = torch.empty((2160, 3840) + tuple(img.shape[2:]), dtype=torch.uint8, device="cuda")
        best = []
        with torch.cuda.stream(stream):
            for rep in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(32):
                    s.process_device(m, img, 2.0, out=out, stream=stream.cuda_stream)
                e1.record(stream)
                stream.synchronize()
                best.append(e0.elapsed_time(e1) / 32)
        best = sorted(best[1:])
        print("%-24s %-18s %-13s ms / frame: min %.4f median %.4f  crc %08x" % (tag, name, label, best[0], best[len(best) // 2], zlib.crc32(out.cpu().numpy().tobytes())), flush=True)
