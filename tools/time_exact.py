import os, sys, zlib
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, anime4kcpp_b200 as A, oracle_lib as O
tag = os.path.basename(os.environ.get("ACB200_LIB", "libac_b200.so"))
img = torch.from_numpy(O.noise_u8(1080, 1920, 1, seed=3)).cuda()
st = torch.cuda.Stream()
for name in ("acnet-legacy-hdn0", "acnet-f8b8-hdn", "arnet-f8b8"):
    m = A.Model(name); s = A.Session(0); s.set_engine(0)
    out = torch.empty((2160, 3840), dtype=torch.uint8, device="cuda")
    best = []
    with torch.cuda.stream(st):
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(8): s.process_device(m, img, 2.0, out=out, stream=st.cuda_stream)
            e1.record(st); st.synchronize(); best.append(e0.elapsed_time(e1) / 8)
    print("%-22s %-18s exact engine ms / frame: %.4f  crc %08x" % (tag, name, min(best[1:]), zlib.crc32(out.cpu().numpy().tobytes())), flush=True)
