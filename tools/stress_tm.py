#!/usr/bin/env python
"""Run-to-run identity of the TMEM engine under load: the same 1080p frame many times, on two streams at once, every result compared
bit for bit with the first (a lost flag / residual-store ordering would show up as a sporadic difference).  GPU box only."""
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import anime4kcpp_b200 as A
import oracle_lib as O

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for name, c in (("acnet-legacy-hdn0", 3), ("acnet-f8b8-hdn", 3), ("arnet-f8b64", 1), ("arnet-f8b8", 3)):
    m = A.Model(name)
    img = torch.from_numpy(O.noise_u8(1080, 1920, c, seed=11)).cuda()
    sess = [A.Session(0), A.Session(0)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[torch.empty((2160, 3840) + tuple(img.shape[2:]), dtype=torch.uint8, device="cuda") for _ in range(reps)] for _ in range(2)]
    for r in range(reps):
        for k in range(2):
            sess[k].process_device(m, img, 2.0, out=outs[k][r], stream=streams[k].cuda_stream)
    torch.cuda.synchronize()
    ref = outs[0][0]
    bad = sum(int(not torch.equal(o, ref)) for k in range(2) for o in outs[k])
    print("%-18s c=%d  %d results on 2 streams, %d differ from the first  crc %08x" % (name, c, 2 * reps, bad, zlib.crc32(ref.cpu().numpy().tobytes())), flush=True)
    assert bad == 0
print("stress ok")
