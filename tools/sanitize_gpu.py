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
for name, shape, c in (("acnet-legacy-hdn0", (70, 90), 3), ("acnet-f8b8-hdn", (50, 64), 1), ("acnet-f8b18", (48, 48), 1), ("arnet-f8b8", (60, 44), 4)):
    img = O.noise_u8(shape[0], shape[1], c, seed=1)
    want = O.oracle_process(name, img, 2.0)
    m = A.Model(name)
    for engine, impl in ((0, 0), (1, 0), (1, 1)):
        s.set_engine(engine)
        s.set_tensor_impl(impl)
        got = s.process_host(m, img, 2.0)
        ok = c == 4 or O.compare_u8(got, want)[0] <= 1
        print(name, "engine", engine, "impl", impl, O.compare_u8(got, want), "ok" if ok else "MISMATCH")
print("done")
