import sys, time, threading
sys.path.insert(0, "anime4kcpp_b200"); sys.path.insert(0, "tests")
import numpy as np, pyac, oracle_lib as O
p = pyac.core.Processor("cuda", 0, "acnet-legacy-hdn0")
imgs = [O.noise_u8(1080, 1920, 3, seed=i) for i in range(4)]
for _ in range(4): out = p(imgs[0])
assert out.flags["C_CONTIGUOUS"] and out.shape == (2160, 3840, 3)
for nt in (1, 4):
    N = 64
    def work(k):
        for i in range(N // nt): p(imgs[k % 4])
    ts = [threading.Thread(target=work, args=(k,)) for k in range(nt)]
    t0 = time.perf_counter(); [t.start() for t in ts]; [t.join() for t in ts]
    print("pyac 1080p RGB, %d python thread(s): %.0f frames/s" % (nt, N / (time.perf_counter() - t0)))
