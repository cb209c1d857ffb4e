#!/usr/bin/env python
"""Luma-network kernel time per engine / tensor implementation on one 1920x1080 gray u8 frame (CUDA events inside the
library, acb200_session_last_kernel_ms), plus the agreement between the implementations.  GPU box only.

    python tools/time_engines.py [model ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import anime4kcpp_b200 as A
import oracle_lib as O

models = sys.argv[1:] or ["acnet-legacy-hdn0", "acnet-f8b8-hdn", "arnet-f8b8-hdn"]
img = O.noise_u8(1080, 1920, 1, seed=3)
for name in models:
    m = A.Model(name)
    outs = {}
    for label, engine, impl in (("exact", 0, 0), ("mma.sync", 1, 0), ("tcgen05", 1, 1)):
        s = A.Session(0)
        s.set_engine(engine)
        s.set_tensor_impl(impl)
        ts = []
        for _ in range(12):
            outs[label] = s.process_host(m, img, 2.0)
            ts.append(s.last_kernel_ms())
        ts = sorted(ts[2:])
        print("%-18s %-9s luma kernel ms: min %.4f median %.4f" % (name, label, ts[0], ts[len(ts) // 2]), flush=True)
    for label in ("mma.sync", "tcgen05"):
        d = np.abs(outs[label].astype(np.int32) - outs["exact"].astype(np.int32))
        print("%-18s %-9s vs exact: max |diff| %d, identical %.4f %%" % (name, label, d.max(), 100.0 * (d == 0).mean()), flush=True)
