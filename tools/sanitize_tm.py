"""Small end-to-end runs of the TMEM-resident engine for compute-sanitizer under gpurun: gray and fused-colour RGB, interior and border
CTAs, ACNetLegacy / ACNet / ARNet (residual store in shared memory, 1x1 tail):
    compute-sanitizer --tool memcheck python tools/sanitize_tm.py
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
s.set_engine(1)
s.set_tensor_impl(2)
for name, shape, c in (("acnet-legacy-hdn0", (70, 90), 3), ("acnet-legacy-hdn0", (150, 150), 1), ("acnet-f8b8-hdn", (50, 200), 3), ("arnet-f8b8", (140, 150), 1),
                       ("arnet-f8b8", (60, 44), 3), ("acnet-f8b18", (48, 48), 3), ("arnet-f8b16-hdn", (5, 7), 1)):
    img = O.noise_u8(shape[0], shape[1], c, seed=1)
    want = O.oracle_process(name, img, 2.0)
    got = s.process_host(A.Model(name), img, 2.0)
    mx, same = O.compare_u8(got, want)
    print(name, shape, c, (mx, same), "ok" if mx <= 1 else "MISMATCH", flush=True)
print("done")
