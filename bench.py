#!/usr/bin/env python
"""Benchmark of the B200 backend on BASELINE.json's metric: output megapixels/s (and 1080p->4K frames/s) of the
CNN upscaling hot path, with the kernel roofline and the reference CPU processor timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model M] [--batch B]

A "step" is one pass of the hot path -- Processor::process(frame, 2.0): RGB->YUV split, the fused luma network,
Catmull-Rom chroma resize, YUV->RGB merge -- over one batch of B synthetic 1920x1080 RGB u8 frames.
 * value : whole-job output MP/s with the batch already resident in HBM (device-resident C-ABI entry, CUDA events
           on the launching stream, max over ranks).
 * e2e   : the same metric through the reference-facing C binding (ac_processor_process) with HOST buffers: pinned
           host images in, pinned host images out, H2D and D2H inside the timed region, caller threads sharing one
           processor exactly as tools/benchmark does.
 * roofline : the luma-network kernel, algorithmic FLOPs (2 x kernelLength() per input pixel) over
           its CUDA-event launch duration, against the measured peaks in MEASURED_PEAKS.json.
 * cpu_baseline : the reference's CPU processor (oracle/_ref, auto-ISA backend, OpenMP) on the box's host cores.
Multi-GPU (torchrun, one rank per GPU): frames are independent, every rank upscales its own batch, no data-path
collective (weak scaling); NCCL is used only for the timing barrier / max-reduce.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

W, H, CH = 1920, 1080, 3
FACTOR = 2.0
OUT_MP = (W * 2) * (H * 2) / 1e6
MACS_PER_PIXEL = {"acnet-legacy": 4712, "acnet-f8b4": 2664, "acnet-f8b8": 4968, "acnet-f8b18": 10728,
                  "arnet-f8b8": 9640, "arnet-f8b16": 18856, "arnet-f8b32": 37288, "arnet-f8b64": 74152,
                  "artcnn-c4f16": 12240, "artcnn-c4f32": 47520, "fsrcnnx-f8b4": 2856, "fsrcnnx-f16b4": 10448}


def macs_for(model):
    for k, v in sorted(MACS_PER_PIXEL.items(), key=lambda kv: -len(kv[0])):
        if model.startswith(k):
            return v
    raise KeyError(model)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arms: the reference's own CPU processor (oracle/_ref) when it travelled with the repo, else the oracle port
# ------------------------------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms must use every host thread the box has (and say how many)."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def cpu_time_frames(model, frames, threads_in_flight=1):
    """Seconds for `frames` 1080p RGB u8 frames through Processor::process(img, 2.0) on the host cores."""
    import numpy as np
    import oracle_lib as O
    use_all_host_threads()
    ref = O.ref()
    if ref is not None:
        t = ref.ref_benchmark(model.encode(), 0, W, H, CH, frames, threads_in_flight, 1234)
        return t, "reference", ref.ref_processor_name(model.encode(), 0).decode()
    img = O.noise_u8(H, W, CH, seed=1234)
    O.oracle_process(model, img[:64, :64], FACTOR)
    t0 = time.perf_counter()
    for _ in range(frames):
        O.oracle_process(model, img, FACTOR)
    return time.perf_counter() - t0, "port", "oracle/ac_oracle.c (Generic order)"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_host_threads()
    frames_per_step = 1
    for _ in range(max(args.warmup, 0)):
        cpu_time_frames(args.model, frames_per_step)
    t, kind, backend = cpu_time_frames(args.model, frames_per_step * args.steps)
    value = OUT_MP * frames_per_step * args.steps / t
    line = {
        "impl": "reference", "metric": "output megapixels/s, 1080p->2160p RGB u8, 2x", "value": value, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s 2x on synthetic 1920x1080 RGB u8 frames (BASELINE configs[1])" % args.model, "frames_per_step": frames_per_step,
                   "fps": frames_per_step * args.steps / t},
        "cpu_baseline": {"value": value, "unit": "MP/s", "cores": cores, "kind": kind, "sample": "%d frame(s) per step, %d steps, backend %s, OpenMP rows over all host threads"
                         % (frames_per_step, args.steps, backend)},
        "e2e": {"value": value, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
class ACImage(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("stride", C.c_int), ("element_type", C.c_int),
                ("ptr", C.c_void_p), ("hptr", C.c_void_p)]


class ACProcessor(C.Structure):
    _fields_ = [("device", C.c_int), ("type", C.c_char_p), ("model", C.c_char_p), ("hptr", C.c_void_p)]


def map_image(lib, arr):
    lib.ac_image_alloc.restype = C.POINTER(ACImage)
    img = lib.ac_image_alloc()
    h, w = arr.shape[:2]
    img.contents.width, img.contents.height, img.contents.channels = w, h, (1 if arr.ndim == 2 else arr.shape[2])
    img.contents.element_type, img.contents.stride, img.contents.ptr = 1, arr.strides[0], arr.ctypes.data
    assert lib.ac_image_map(img) == 0
    return img


IMPL_NAMES = {0: "mma.sync", 1: "tcgen05, maps in shared memory", 2: "tcgen05, maps resident in TMEM"}


def impl_label(model, tensor_impl):
    """Name of the tensor-engine implementation that runs `model` (library default: 2, the TMEM-resident engine, for ACNetLegacy / ACNet / ARNet)."""
    impl = 2 if tensor_impl is None else tensor_impl
    if model.startswith(("artcnn", "fsrcnnx")):
        return "per-layer tcgen05 MMA (wide families)"
    return IMPL_NAMES[impl]


# dram__bytes_read.sum + dram__bytes_write.sum of one luma pass (both segment launches), ncu --cache-control none (caches NOT flushed
# between replays: the steady-state figure), 1920x1080 u8 plane, ACNetLegacy as head + 4 | 3 + tail: profiles/r02_final_luma_pass_traffic.csv
# (two consecutive passes: 66.8 and 61.3 MB; segment A writes the inter-segment map through to DRAM, segment B reads it out of L2)
STEADY_TRAFFIC = {("acnet-legacy", 2): (66821632 + 61305856) // 2}


def luma_roofline(A, torch, sess, model, name, d_plane, stream, peaks, reps):
    """The luma network alone on a device-resident 1080p Y plane: CUDA-event time per pass, algorithmic FLOPs against the BURST bf16 peak."""
    y_out = torch.empty((2 * H, 2 * W), dtype=torch.uint8, device="cuda")
    for i in range(3):
        sess.process_device(model, d_plane[i % len(d_plane)], FACTOR, out=y_out, stream=stream)
    torch.cuda.synchronize()
    l0 = A.launch_count()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(reps):
        sess.process_device(model, d_plane[i % len(d_plane)], FACTOR, out=y_out, stream=stream)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / reps
    launches = (A.launch_count() - l0) / reps
    flop = 2.0 * macs_for(name) * W * H
    tf = flop / (kernel_ms / 1e3) / 1e12
    return {"kernel_ms": kernel_ms, "launches_per_pass": launches, "flop_per_pass": flop, "flop_per_launch": flop / max(launches, 1.0),
            "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"]}


def run_ours(args):
    import numpy as np
    import torch
    import anime4kcpp_b200 as A
    import oracle_lib as O

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert A.device_count() > local, "no CUDA device: the B200 backend has no CPU fallback"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL_DEBUG=VERSION makes NCCL print its banner on STDOUT, ahead of the one JSON line this script owes its caller
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lib = A.lib()
    B = args.batch
    model = A.Model(args.model)
    sess = A.Session(local)
    sess.set_engine(args.engine)
    if args.tensor_impl is not None:
        sess.set_tensor_impl(args.tensor_impl)
    rs = np.random.RandomState(1234 + rank)
    host_frames = rs.randint(0, 256, size=(B, H, W, CH), dtype=np.uint8)
    d_in = torch.from_numpy(host_frames).cuda()
    d_out = torch.empty((B, 2 * H, 2 * W, CH), dtype=torch.uint8, device="cuda")
    work_stream = torch.cuda.Stream()
    torch.cuda.set_stream(work_stream)          # kernels and torch.cuda.Event timing share this stream
    stream = work_stream.cuda_stream

    # The batch is dealt over `--streams` sessions (stream + scratch each), frame i on stream i mod S: the tail wave of one
    # frame's luma kernel and the small colour kernels run beside the next frame's CTAs instead of leaving SMs idle.  The timed
    # region is bracketed on work_stream: the side streams wait for its start event, it waits for their last events.
    n_streams = max(1, args.streams)
    side = [(A.Session(local), torch.cuda.Stream()) for _ in range(n_streams - 1)]
    for ss, _ in side:
        ss.set_engine(args.engine)
        if args.tensor_impl is not None:
            ss.set_tensor_impl(args.tensor_impl)
    lanes = [(sess, work_stream)] + side

    def step():
        fork = torch.cuda.Event()
        fork.record(work_stream)
        for _, st_ in side:
            st_.wait_event(fork)
        for i in range(B):
            ss, st_ = lanes[i % n_streams]
            ss.process_device(model, d_in[i], FACTOR, out=d_out[i], stream=st_.cuda_stream)
        for _, st_ in side:
            join = torch.cuda.Event()
            join.record(st_)
            work_stream.wait_event(join)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = A.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = A.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    frames_total = B * args.steps * world
    value = OUT_MP * frames_total / (ms_max / 1e3)

    # ---- spot parity inside the bench: a crop of the first output frame against the CPU oracle ---------------------
    parity = None
    if rank == 0:
        crop = host_frames[0, :96, :128]
        got = sess.process_host(model, np.ascontiguousarray(crop), FACTOR)
        mx, exact = O.compare_u8(got, O.oracle_process(args.model, np.ascontiguousarray(crop), FACTOR))
        parity = {"max_lsb": mx, "bit_exact_frac": exact}

    # ---- dominant kernel alone: the luma network on a device-resident 1080p Y plane (what the RGB path launches) ----
    y_in = d_in[:, :, :, 0].contiguous()
    peaks = load_peaks()
    lr = luma_roofline(A, torch, sess, model, args.model, y_in, stream, peaks, max(8, min(64, B * args.steps)))
    kernel_ms, launches_per_pass, flop_frame, achieved_tf = lr["kernel_ms"], lr["launches_per_pass"], lr["flop_per_pass"], lr["achieved"]
    # bytes that must cross HBM per RGB frame (in + out) -- the other roofline candidate; compute dominates
    bytes_frame = W * H * CH + 4 * W * H * CH
    t_roof_ms = max(flop_frame / (peaks["bf16_tflops"] * 1e12), bytes_frame / (peaks["hbm_gbs"] * 1e9)) * 1e3

    # ---- end to end through the C binding with host buffers ---------------------------------------------------------
    e2e = None
    # workers of the ordered streams: 4 per GPU; callers of the e2e legs: see --threads
    n_threads = 4
    e2e_threads = args.threads if args.threads else (4 if world <= 2 else 2 if world <= 4 else 1)
    yuv_threads = args.yuv_threads if args.yuv_threads else 4
    pin_in = torch.from_numpy(host_frames).pin_memory()
    pin_out = torch.empty((B, 2 * H, 2 * W, CH), dtype=torch.uint8).pin_memory()
    lib.ac_processor_alloc.restype = C.POINTER(ACProcessor)
    lib.ac_processor_error.restype = C.c_char_p
    proc = lib.ac_processor_alloc()
    proc.contents.type, proc.contents.model, proc.contents.device = b"cuda", args.model.encode(), local
    assert lib.ac_processor_create(proc) == 0, lib.ac_processor_error(proc)
    np_in, np_out = pin_in.numpy(), pin_out.numpy()
    srcs = [map_image(lib, np_in[i]) for i in range(B)]
    dsts = [map_image(lib, np_out[i]) for i in range(B)]

    def e2e_steps(n_steps):
        # n_steps batches of B frames; the caller threads (tools/benchmark's "N concurrent images" mode) live across the steps and
        # pull frame after frame, every call a full ac_processor_process: H2D, the GPU pass, D2H, synchronise
        nxt = [0]
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= n_steps * B:
                    return
                rc = lib.ac_processor_process(proc, srcs[i % B], dsts[i % B], C.c_double(FACTOR))
                assert rc == 0, lib.ac_processor_error(proc)
        ts = [threading.Thread(target=worker) for _ in range(e2e_threads)]
        [x.start() for x in ts]
        [x.join() for x in ts]

    e2e_steps(2)
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = OUT_MP * frames_total / float(te.item())
    if rank == 0:
        want = d_out[0].cpu().numpy()
        if args.engine == 2:
            assert np.array_equal(np_out[0], want), "host path and device path disagree"

    # The ceiling of any host-fed path on this box: the SAME bytes (one pinned 1080p RGB frame in, one pinned 2160p RGB frame out, a
    # synchronise per frame) from the same number of caller threads, each on its own stream, with NO kernel in between.
    def copy_steps(n_steps):
        nxt = [0]
        lock = threading.Lock()

        def worker(k):
            st_ = torch.cuda.Stream()
            with torch.cuda.stream(st_):
                while True:
                    with lock:
                        i = nxt[0]
                        nxt[0] += 1
                    if i >= n_steps * B:
                        return
                    d_in[i % B].copy_(pin_in[i % B], non_blocking=True)
                    pin_out[i % B].copy_(d_out[i % B], non_blocking=True)
                    st_.synchronize()
        ts = [threading.Thread(target=worker, args=(k,)) for k in range(e2e_threads)]
        [x.start() for x in ts]
        [x.join() for x in ts]
    copy_steps(1)
    barrier()
    t0 = time.perf_counter()
    copy_steps(args.steps)
    barrier()
    tcp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tcp, op=dist.ReduceOp.MAX)
    copy_value = OUT_MP * frames_total / float(tcp.item())
    copy_gbs = frames_total * 5 * W * H * CH / float(tcp.item()) / 1e9
    e2e = {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": B * W * H * CH, "d2h_bytes_per_step": B * 4 * W * H * CH,
           "fps": frames_total / float(te.item()), "caller_threads": e2e_threads, "api": "ac_processor_process (libac_c binding), pinned host images",
           "copy_ceiling": {"value": copy_value, "unit": "MP/s", "gb_per_s_both_directions": copy_gbs,
                            "what": "the same pinned H2D + D2H bytes per frame from the same caller threads, no kernels; all ranks at once"},
           "frac_of_copy_ceiling": e2e_value / copy_value}

    # ---- the video callers' format: planar YUV420 u8 frames, Y through the network, U and V
    #      through the Catmull-Rom resize, one submission per frame; same metric, counted on the luma plane ----------------
    yuv = None
    if not args.no_yuv:
        f_y = [d_in[i, :, :, 0].contiguous() for i in range(B)]
        f_u = [d_in[i, ::2, ::2, 1].contiguous() for i in range(B)]
        f_v = [d_in[i, ::2, ::2, 2].contiguous() for i in range(B)]
        o_y = [torch.empty((2 * H, 2 * W), dtype=torch.uint8, device="cuda") for _ in range(B)]
        o_u = [torch.empty((H, W), dtype=torch.uint8, device="cuda") for _ in range(B)]
        o_v = [torch.empty((H, W), dtype=torch.uint8, device="cuda") for _ in range(B)]

        def yuv_step():
            fork = torch.cuda.Event()
            fork.record(work_stream)
            for _, st_ in side:
                st_.wait_event(fork)
            for i in range(B):
                ss, st_ = lanes[i % n_streams]
                ss.process_frame_device(model, [f_y[i], f_u[i], f_v[i]], [o_y[i], o_u[i], o_v[i]], FACTOR, 0, st_.cuda_stream)
            for _, st_ in side:
                join = torch.cuda.Event()
                join.record(st_)
                work_stream.wait_event(join)
        for _ in range(3):
            yuv_step()
        barrier()
        y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        y0.record()
        for _ in range(args.steps):
            yuv_step()
        y1.record()
        barrier()
        ty = torch.tensor([y0.elapsed_time(y1)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(ty, op=dist.ReduceOp.MAX)
        yuv_ms = float(ty.item())
        # end to end: pinned host planes through the C binding's frame call, caller threads sharing the processor
        hp_in = [[t.cpu().pin_memory() for t in (f_y[i], f_u[i], f_v[i])] for i in range(B)]
        hp_out = [[torch.empty(t.shape, dtype=torch.uint8).pin_memory() for t in (o_y[i], o_u[i], o_v[i])] for i in range(B)]
        lib.ac_processor_process_frame.argtypes = [C.POINTER(ACProcessor), C.POINTER(A.Plane), C.POINTER(A.Plane), C.c_int, C.c_int, C.c_int, C.c_double]
        pl_in = [A._planes_of([t.numpy() for t in f]) for f in hp_in]
        pl_out = [A._planes_of([t.numpy() for t in f]) for f in hp_out]

        def yuv_e2e_steps(n_steps):
            nxt = [0]
            lock = threading.Lock()

            def worker():
                while True:
                    with lock:
                        i = nxt[0]
                        nxt[0] += 1
                    if i >= n_steps * B:
                        return
                    rc = lib.ac_processor_process_frame(proc, pl_in[i % B], pl_out[i % B], 3, 1, 0, FACTOR)
                    assert rc == 0, lib.ac_processor_error(proc)
            ts = [threading.Thread(target=worker) for _ in range(yuv_threads)]
            [x.start() for x in ts]
            [x.join() for x in ts]
        yuv_e2e_steps(2)
        barrier()
        t0 = time.perf_counter()
        yuv_e2e_steps(args.steps)
        barrier()
        tye = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tye, op=dist.ReduceOp.MAX)
        if rank == 0 and args.engine == 2:
            assert all(np.array_equal(a.numpy(), b.cpu().numpy()) for a, b in zip(hp_out[0], (o_y[0], o_u[0], o_v[0]))), "yuv host and device paths disagree"
        yuv = {"workload": "planar YUV420 u8 1080p frames: Y through the network, U/V 960x540 -> 1920x1080 Catmull-Rom, one submission per frame",
               "value": OUT_MP * frames_total / (yuv_ms / 1e3), "unit": "MP/s", "fps": frames_total / (yuv_ms / 1e3),
               "e2e": {"value": OUT_MP * frames_total / float(tye.item()), "unit": "MP/s", "fps": frames_total / float(tye.item()),
                       "h2d_bytes_per_step": B * W * H * 3 // 2, "d2h_bytes_per_step": B * 4 * W * H * 3 // 2,
                       "api": "ac_processor_process_frame (C binding extension), pinned host planes", "caller_threads": yuv_threads}}

    # ---- a stream of 256 frames through the ordered multi-worker frame stream (the worker / ordering core
    #      of the reference's video filter), packed RGB and planar YUV420, host buffers in and out; wall clock, max over ranks -------
    stream_res = None
    if not args.no_yuv:
        NFR = 256
        fs = A.FrameStream(model, [local], workers_per_device=n_threads, queue_depth=4)

        def run_stream(submit):
            barrier()
            t0 = time.perf_counter()
            done = 0
            for i in range(NFR):
                submit(i % B)
                if i >= 2 * n_threads:
                    fs.next(); done += 1
            while done < NFR:
                fs.next(); done += 1
            barrier()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return NFR * world / float(tt.item())
        run_stream(lambda i: fs.submit(np_in[i], FACTOR, np_out[i]))                 # warm-up (sessions, scratch)
        rgb_fps = run_stream(lambda i: fs.submit(np_in[i], FACTOR, np_out[i]))
        fy_in = [[t.numpy() for t in f] for f in hp_in]
        fy_out = [[t.numpy() for t in f] for f in hp_out]
        run_stream(lambda i: fs.submit_frame(fy_in[i], FACTOR, fy_out[i]))
        yuv_fps = run_stream(lambda i: fs.submit_frame(fy_in[i], FACTOR, fy_out[i]))
        fs.close()
        stream_res = {"frames": NFR * world, "workers_per_gpu": n_threads, "queue_depth": 4, "rgb_fps": rgb_fps, "yuv420_fps": yuv_fps,
                      "api": "acb200_stream_submit / submit_frame / next (in-order delivery), pinned host frames"}

    # ---- BASELINE configs 2 / 3: the other model families on the same frames (device-resident batch + the luma pass alone) ----------
    models_res = None
    if not args.no_extra:
        models_res = {}
        for name in args.models.split(","):
            if not name or name == args.model:
                continue
            mm = A.Model(name)
            nb = min(B, 4)

            def mstep():
                for i in range(nb):
                    ss, st_ = lanes[i % n_streams]
                    ss.process_device(mm, d_in[i], FACTOR, out=d_out[i], stream=st_.cuda_stream)
                for _, st_ in side:
                    join = torch.cuda.Event()
                    join.record(st_)
                    work_stream.wait_event(join)
            for _ in range(2):
                mstep()
            barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            nrep = 3
            for _ in range(nrep):
                mstep()
            m1.record()
            barrier()
            tm_ = torch.tensor([m0.elapsed_time(m1)], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(tm_, op=dist.ReduceOp.MAX)
            r = luma_roofline(A, torch, sess, mm, name, y_in, stream, peaks, 8)
            crop = np.ascontiguousarray(host_frames[0, :96, :128])
            mx, exact = O.compare_u8(sess.process_host(mm, crop, FACTOR), O.oracle_process(name, crop, FACTOR))
            models_res[name] = {"value": OUT_MP * nb * nrep * world / (float(tm_.item()) / 1e3), "unit": "MP/s",
                                "fps": nb * nrep * world / (float(tm_.item()) / 1e3), "frames_per_step_per_gpu": nb,
                                "tensor_impl": impl_label(name, args.tensor_impl), "parity_spot_check": {"max_lsb": mx, "bit_exact_frac": exact},
                                "roofline": dict(r, bound="tensor", kernel="luma network alone, 1920x1080 Y -> 3840x2160 Y")}

    # ---- BASELINE config 5: one 8192 x 8192 image, 4x (two 2x passes), eight halo-overlapped row bands dealt over the ranks; host image
    #      in, host image out (gray: the 32768^2 RGB result does not fit the reference's int-sized Image) -----------------------------
    bands_res = None
    if not args.no_extra:
        NB_, SZ = 8, args.band_size
        # pinned host memory on both sides, like the other host-fed legs; every rank keeps the bands it owns (the gather of a multi-GPU
        # job is "each band is already where its rank wrote it")
        big = torch.from_numpy(np.random.RandomState(99).randint(0, 256, size=(SZ, SZ), dtype=np.uint8)).pin_memory().numpy()
        mine = [b for b in range(NB_) if b % world == rank]
        halo_rows = model.halo()
        plans = {b: A.band_plan(SZ, 4.0, halo_rows, NB_, b) for b in mine}
        band_out = {b: torch.empty((plans[b][3] - plans[b][2], 4 * SZ), dtype=torch.uint8).pin_memory().numpy() for b in mine}
        A.process_band(sess, model, big, 4.0, NB_, mine[0], band_out[mine[0]], out_y0=plans[mine[0]][2])        # warm-up: scratch at band size
        barrier()
        t0 = time.perf_counter()
        for b in mine:
            A.process_band(sess, model, big, 4.0, NB_, b, band_out[b], out_y0=plans[b][2])
        torch.cuda.synchronize()
        t_mine = time.perf_counter() - t0
        barrier()
        band_parity = None
        if rank == 0 and mine[0] == 0:
            # the top-left 384 x 256 output pixels of band 0 against the oracle on the source crop that contains their whole context
            want = O.oracle_process(args.model, np.ascontiguousarray(big[:64 + 32, :96 + 32]), 4.0)[:256, :384]
            mxb, exb = O.compare_u8(np.ascontiguousarray(band_out[0][:256, :384]), want)
            band_parity = {"max_lsb": mxb, "bit_exact_frac": exb, "what": "256 x 384 output crop of band 0 vs the CPU oracle (Generic order), 4x"}
        tb = torch.tensor([t_mine], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        flop_big = 2.0 * macs_for(args.model) * (SZ * SZ + 4 * SZ * SZ)
        tfb = flop_big / float(tb.item()) / 1e12
        bands_res = {"workload": "%dx%d gray u8 image, 4x, %d row bands, band b on rank b mod N, host image in / host image out" % (SZ, SZ, NB_),
                     "value": 16 * SZ * SZ / 1e6 / float(tb.item()), "unit": "MP/s", "seconds": float(tb.item()), "bands_per_rank": len(mine),
                     "h2d_bytes": SZ * SZ, "d2h_bytes": 16 * SZ * SZ, "host_memory": "pinned", "parity_spot_check": band_parity,
                     "roofline": {"bound": "tensor", "achieved": tfb, "peak": peaks["bf16_tflops"] * world, "unit": "TFLOP/s",
                                  "frac": tfb / (peaks["bf16_tflops"] * world),
                                  "note": "end to end (copies and both passes inside the timed region) against the burst peak of N GPUs"}}
        del band_out

    # ---- BASELINE config 4, single process: ONE frame stream over every GPU of the job (rank 0 drives them all; the other ranks wait) ----
    inproc_res = None
    if not args.no_extra:
        barrier()
        if rank == 0:
            ndev = min(world, A.device_count())
            fs = A.FrameStream(model, list(range(ndev)), workers_per_device=n_threads, queue_depth=4)
            NFR = 128 * ndev

            def run_inproc():
                t0 = time.perf_counter()
                done = 0
                for i in range(NFR):
                    fs.submit(np_in[i % B], FACTOR, np_out[i % B])
                    if i >= 2 * n_threads * ndev:
                        fs.next(); done += 1
                while done < NFR:
                    fs.next(); done += 1
                return time.perf_counter() - t0
            run_inproc()
            tt = run_inproc()
            fs.close()
            tfi = NFR * flop_frame / tt / 1e12
            inproc_res = {"workload": "%d synthetic 1080p RGB u8 frames, one process, frame n on GPU n mod %d, in-order delivery, pinned host frames" % (NFR, ndev),
                          "devices": ndev, "workers_per_gpu": n_threads, "fps": NFR / tt, "value": OUT_MP * NFR / tt, "unit": "MP/s",
                          "roofline": {"bound": "tensor", "achieved": tfi, "peak": peaks["bf16_tflops"] * ndev, "unit": "TFLOP/s", "frac": tfi / (peaks["bf16_tflops"] * ndev),
                                       "note": "luma-network FLOPs of the delivered frames over wall time, host copies included"}}
        barrier()

    # ---- small frames (tools/benchmark's 720x480 gray case): launch-bound -- latency of one synchronous host call, and device-resident rate ----
    small_res = None
    if not args.no_extra and rank == 0:
        sh, sw = 480, 720
        g = np.random.RandomState(5).randint(0, 256, size=(sh, sw), dtype=np.uint8)
        for _ in range(5):
            sess.process_host(model, g, FACTOR)
        t0 = time.perf_counter()
        for _ in range(200):
            sess.process_host(model, g, FACTOR)
        lat = (time.perf_counter() - t0) / 200
        dg = torch.from_numpy(g).cuda()
        dgo = torch.empty((2 * sh, 2 * sw), dtype=torch.uint8, device="cuda")
        for _ in range(5):
            sess.process_device(model, dg, FACTOR, out=dgo, stream=stream)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s0.record()
        for _ in range(400):
            sess.process_device(model, dg, FACTOR, out=dgo, stream=stream)
        s1.record()
        t_submit = (time.perf_counter() - t0) / 400
        torch.cuda.synchronize()
        # the reference tool's harness shape (tools/benchmark/src/Benchmark.cpp:45-66): `batch` images through ONE shared processor from a
        # pool of caller threads, every call a complete ac_processor_process (H2D, the pass, D2H, synchronise); pinned host images
        NPOOL = 16
        sp_in = torch.from_numpy(np.random.RandomState(6).randint(0, 256, size=(NPOOL, sh, sw), dtype=np.uint8)).pin_memory()
        sp_out = torch.empty((NPOOL, 2 * sh, 2 * sw), dtype=torch.uint8).pin_memory()
        s_srcs = [map_image(lib, sp_in.numpy()[i]) for i in range(NPOOL)]
        s_dsts = [map_image(lib, sp_out.numpy()[i]) for i in range(NPOOL)]

        def api_fps(n_frames, n_thr):
            nxt = [0]
            lock = threading.Lock()

            def worker():
                k = threading.get_ident() % NPOOL       # a caller keeps its own destination image
                while True:
                    with lock:
                        i = nxt[0]
                        nxt[0] += 1
                    if i >= n_frames:
                        return
                    rc = lib.ac_processor_process(proc, s_srcs[i % NPOOL], s_dsts[k], C.c_double(FACTOR))
                    assert rc == 0, lib.ac_processor_error(proc)
            ts = [threading.Thread(target=worker) for _ in range(n_thr)]
            t0 = time.perf_counter()
            [x.start() for x in ts]
            [x.join() for x in ts]
            return n_frames / (time.perf_counter() - t0)
        api_fps(60, 8)
        api = {"1": api_fps(600, 1), "8": api_fps(600, 8)}
        small_res = {"workload": "720x480 gray u8, 2x", "host_call_latency_ms": lat * 1e3, "host_call_fps": 1.0 / lat,
                     "device_resident_fps": 400 / (s0.elapsed_time(s1) / 1e3), "device_resident_ms": s0.elapsed_time(s1) / 400,
                     "host_submit_us_per_frame": t_submit * 1e6,
                     "ac_processor_process_fps_by_caller_threads": api,
                     "api_note": "600 images, one shared processor, pinned host images (the reference benchmark tool's harness shape)"}

    cpu = None
    if rank == 0 and not args.no_cpu:
        frames = args.cpu_frames
        tc, kind, backend = cpu_time_frames(args.model, frames)
        cpu = {"value": OUT_MP * frames / tc, "unit": "MP/s", "cores": use_all_host_threads(), "kind": kind,
               "sample": "%d 1080p RGB frames, backend %s, one frame at a time with OpenMP over all host threads" % (frames, backend),
               "fps": frames / tc}

    if rank == 0:
        line = {
            "metric": "output megapixels/s, 1080p->2160p RGB u8, 2x", "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s 2x on a batch of %d synthetic 1920x1080 RGB u8 frames per GPU (BASELINE configs[1])" % (args.model, B),
                       "frames_per_step_per_gpu": B, "fps": frames_total / (ms_max / 1e3), "engine": args.engine, "streams": n_streams,
                       "tensor_impl": impl_label(args.model, args.tensor_impl),
                       "cache": "inputs larger than L2: %d MB in + %d MB out per step" % (B * W * H * CH >> 20, B * 4 * W * H * CH >> 20),
                       "parity_spot_check": parity},
            # the dominant kernel timed ALONE against the BURST tensor peak of MEASURED_PEAKS.json (dense bf16; the split-fp16 scheme
            # spends three tensor products per algorithmic product, so 1/3 is the ceiling of this formulation)
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved_tf / peaks["bf16_tflops"],
                         "traffic": STEADY_TRAFFIC.get((args.model.split("-hdn")[0], 2 if args.tensor_impl is None else args.tensor_impl)) if args.engine != 0 else None,
                         "traffic_note": "dram bytes read + written by the pass's launches in steady state (ncu --cache-control none: the inter-segment map is consumed out of L2)",
                         "algorithmic_bytes": W * H + 4 * W * H,
                         "kernel": ("luma network, one launch per layer (launches_per_pass), 1920x1080 Y -> 3840x2160 Y" if args.model.startswith(("artcnn", "fsrcnnx"))
                                    else "fused luma network (segment kernels back to back: launches_per_pass), 1920x1080 Y -> 3840x2160 Y"), "kernel_ms": kernel_ms,
                         "launches_per_pass": launches_per_pass, "flop_per_launch": flop_frame / max(launches_per_pass, 1.0), "flop_per_pass": flop_frame,
                         "peak_source": peaks["source"] + ", burst",
                         "pipe": "fp32 FFMA (CUDA cores), exact engine" if (args.engine == 0 or args.model.startswith("fsrcnnx-f8"))
                                 else "split-fp16 tensor-core MMA, 3 tensor products per algorithmic product: " + impl_label(args.model, args.tensor_impl),
                         "frame_roofline_ms": t_roof_ms, "frame_frac": t_roof_ms / (ms_max / (B * args.steps))},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "yuv420": yuv, "stream": stream_res,
            "models": models_res, "bands": bands_res, "stream_inproc": inproc_res, "small_frame": small_res,
        }
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The ONE JSON line, on the process's original stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Libraries print on stdout too (NCCL's version banner under NCCL_DEBUG=VERSION, OpenMP / loader notices): keep a private
    # handle on the real stdout for the JSON line and point fd 1 at stderr for everything else.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="acnet-legacy-hdn0")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--threads", type=int, default=None,
                    help="caller threads per rank sharing the processor in the host-fed e2e legs (default: 4 on 1-2 GPUs, 2 on 4, 1 on 8 -- the "
                         "boxes' aggregate copy ceiling is reached with fewer callers as ranks are added, profiles/r02_e2e_caller_thread_sweep_n8.txt)")
    ap.add_argument("--yuv-threads", type=int, default=None, help="caller threads per rank of the planar-YUV420 e2e leg (default 4)")
    ap.add_argument("--streams", type=int, default=2, help="sessions / CUDA streams the device-resident batch is dealt over")
    ap.add_argument("--engine", type=int, default=2, help="0 exact FFMA, 1 tensor MMA, 2 auto")
    ap.add_argument("--tensor-impl", type=int, default=None, help="0 mma.sync, 1 tcgen05 (default: library default)")
    ap.add_argument("--cpu-frames", type=int, default=12)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-yuv", action="store_true", help="skip the planar-YUV420 (video caller) section")
    ap.add_argument("--no-extra", action="store_true", help="skip the models / bands / stream_inproc / small_frame sections")
    ap.add_argument("--models", default="acnet-f8b8-hdn,arnet-f8b64", help="other models measured on the same frames (BASELINE configs 2 / 3)")
    ap.add_argument("--band-size", type=int, default=8192, help="edge of the square image of the bands section (BASELINE config 5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
