#!/usr/bin/env python
"""bench.py — headline benchmark of the registration hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], "C2"): ICP-only point-to-point alignment, 50 000-point source vs 50 000-point
target, exactly 50 iterations (convergence tests evaluated but not acted on), max-correspondence-distance 0.05 m.
A step = one align() call = target index build + 50 fused ICP iterations. metric = ICP iterations per second,
whole job (N ranks each align their own independent pair: weak scaling, no data-path collective).

  value     inputs already resident in HBM; per-step CUDA-event time on the library's stream; L2 flushed between steps
  e2e       the same call through the C ABI with HOST buffers: H2D of both clouds, align, D2H of the 4x4 + aligned cloud
  roofline  the fused icp_kernel: algorithmic bytes (16*Ns + 16*Nt + 64 per iteration x 50) / its CUDA-event duration
  cpu_baseline  the CPU oracle (single-threaded restatement of the PCL path; PCL itself cannot be built here) timed
                on this box's host on the same pair
  pipeline  (N=1 only) frames/s of the full FPFH + SAC-IA + ICP estimateFinalPose on a synthetic 640x480 frame ("C1"),
            device-resident and end-to-end, beside the oracle on the host

--impl reference times the CPU oracle port on all host cores (one independent alignment per core).
"""
import argparse
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # ope_pose_batch: one hardware queue per worker stream (before CUDA starts)
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

N_PTS = 50000
ICP_ITERS = 50
MAX_CORR = 0.05
METRIC = "icp_iterations_per_sec_50k"
UNIT = "iterations/s"
WORKLOAD = "C2: ICP-only point-to-point, 50k-point source vs 50k-point target, 50 iterations, max-corr-distance 0.05"


def icp_kwargs():
    return dict(max_iterations=ICP_ITERS, max_correspondence_distance=MAX_CORR, transformation_epsilon=1e-8,
                euclidean_fitness_epsilon=1e-8, force_all_iterations=1)


# DRAM bytes of ONE launch from ncu --set full captures of these exact workloads (committed under profiles/): what the kernels really
# moved, to set beside the algorithmic bytes. icp_kernel: the 1.6 MB pair is read once and then lives in L2 (80 MB algorithmic over 50
# iterations, 5.2 MB from DRAM); depth_fused_kernel: DRAM traffic equals the algorithmic bytes.
NCU_ICP_DRAM_BYTES = 4826624 + 816128
NCU_DEPTH_DRAM_BYTES = 629331712 + 4976560000


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples SM clock and throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------ reference arm ----
_REF = {}


def _ref_init():
    """per-worker: build the C2 pair once, outside every timed region"""
    import multiprocessing as mp
    import orc_py
    import ope_pkg
    ope_pkg.load()
    from ope_b200 import synth
    ident = mp.current_process()._identity
    seed = ident[0] if ident else 0
    _REF["pair"] = synth.icp_pair(N_PTS, seed=seed)[:2]
    _REF["orc"] = orc_py
    orc_py.lib()


def _ref_worker(iters):
    orc_py = _REF["orc"]
    src, tgt = _REF["pair"]
    kw = icp_kwargs()
    kw["max_iterations"] = iters
    res = orc_py.icp(src, tgt, orc_py.icp_params(**kw))
    return res.iterations


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    iters = 10  # bounded sample: 10 of the 50 iterations per alignment, one alignment per core per step
    with mp.get_context("fork").Pool(cores, initializer=_ref_init) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_ref_worker, [iters] * cores, chunksize=1)
        t0 = time.perf_counter()
        done = 0
        for _ in range(args.steps):
            done += sum(pool.map(_ref_worker, [iters] * cores, chunksize=1))
        wall = time.perf_counter() - t0
    value = done / wall
    sample = ("per step: %d parallel alignments (one per host core) of a C2 pair, %d forced iterations each, kd-tree build "
              "included" % (cores, iters))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU restatement of the PCL path (PCL itself unavailable), all host cores"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ our arm ----
def run_ours(args):
    import torch
    import ope_pkg
    ope_pkg.load()
    from ope_b200 import cuda_lib, synth
    T = cuda_lib.T

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"     # the version banner goes to stdout; this run prints exactly one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    stream = torch.cuda.Stream()       # a real handle: the legacy default stream (0) would make the ctx create its own
    torch.cuda.set_stream(stream)      # torch work (L2 flush, events) and the library's kernels share ONE stream
    ctx = cuda_lib.Context(local, stream.cuda_stream)
    model = synth.make_model()
    # weak scaling: every rank aligns its own copy of the SAME synthetic pair, so that all ranks do equal work and the max over
    # ranks measures the machine, not the luck of a seed (pairs drawn with different seeds differ by 30 % in far-query load)
    src, tgt, _ = synth.icp_pair(N_PTS, seed=0, model=model)
    prm = cuda_lib.icp_params(**icp_kwargs())
    cs, ct = ctx.upload(src), ctx.upload(tgt)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        ctx.invalidate(ct)            # the reference rebuilds its kd-tree per align(); so do we
        return ctx.icp(cs, ct, prm)

    for _ in range(max(args.warmup, 3)):
        res = step_device()
    assert res.iterations == ICP_ITERS, res.iterations

    # ---- value: device-resident inputs, per-step CUDA events, L2 flush between steps ----
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = ctx.launches
    step_ms, kern_ms = [], []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device()
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        kern_ms.append(ctx.last_kernel_ms(0))
    barrier()
    launches = ctx.launches - l0
    total_ms = float(np.sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = world * ICP_ITERS * args.steps / (total_ms_max / 1e3)

    # ---- e2e: host buffers through the C ABI every step ----
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        a, b = ctx.upload(src), ctx.upload(tgt)
        r, aligned = ctx.icp(a, b, prm, want_aligned=True)
        out = aligned.download()
        a.free(); b.free(); aligned.free()
    ev1.record(stream)
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * ICP_ITERS * args.steps / (float(t.item()) / 1e3)
    clocks = sampler.stop()
    h2d = (len(src) + len(tgt)) * 16
    d2h = len(src) * 16 + 104

    # ---- roofline of the dominant kernel (icp_kernel) ----
    peak, peak_src = peaks()
    alg_bytes = ICP_ITERS * (16 * len(src) + 16 * len(tgt) + 64)
    k_ms = float(np.mean(kern_ms))
    achieved = alg_bytes / (k_ms / 1e3) / 1e9
    roofline = {"kernel": "icp_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": NCU_ICP_DRAM_BYTES, "traffic_source": "ncu --set full, profiles/r01_g_icp_full.txt "
                "(dram__bytes_read.sum + dram__bytes_write.sum of one icp_kernel launch on this workload)",
                "peak_source": peak_src, "kernel_ms": k_ms,
                "kernel_share_of_step": k_ms / (total_ms / args.steps),
                "note": "one C2 alignment is 1.6 MB and stays in L2: the loop is search-latency/grid-barrier bound, not HBM bound "
                        "(SURVEY 8d); frac is reported against HBM as the contract asks"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "iterations_per_step": ICP_ITERS, "l2": "flushed between steps (256 MiB write)",
                       "sharding": "one independent pair per rank, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline}

    if rank == 0 and world == 1:
        line["cpu_baseline"] = cpu_baseline(src, tgt)
        try:
            line["feature_gemm"] = feature_gemm_numbers(ctx)
        except Exception as e:
            line["feature_gemm"] = {"error": repr(e)}
        try:
            line["depth_to_cloud"] = depth_numbers(ctx, synth, model, stream, torch)
        except Exception as e:
            line["depth_to_cloud"] = {"error": repr(e)}
        try:
            line["pipeline"] = pipeline_numbers(ctx, cuda_lib, synth, model, stream, torch)
        except Exception as e:  # the headline line must survive a failure of the secondary workload
            line["pipeline"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def cpu_baseline(src, tgt):
    """the oracle (CPU restatement of the PCL path) on one host core, same pair, all 50 iterations"""
    import orc_py
    t0 = time.perf_counter()
    res = orc_py.icp(src, tgt, orc_py.icp_params(**icp_kwargs()))
    dt = time.perf_counter() - t0
    return {"value": res.iterations / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "one full alignment of the same 50k/50k pair (%d iterations, kd-tree build included), %.1f s" % (res.iterations, dt)}


def feature_gemm_numbers(ctx, nq=18944, nt=307200):
    """K6 at C3 scale (FPFH-space 5-NN of nq source descriptors in a 307 200-descriptor scene): the tcgen05/TMA distance GEMM with
    fused candidate selection, timed by its own CUDA events; tensor roofline against the measured bf16 peak."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(p))["bf16_tflops"]) if os.path.exists(p) else 1590.0
    rng = np.random.default_rng(0)
    ft = rng.gamma(0.6, 8.0, size=(nt, 33)).astype(np.float32)
    ft = (100.0 * ft / ft.reshape(nt, 3, 11).sum(2).repeat(11, 1)).astype(np.float32)
    fq = ft[rng.integers(0, nt, nq)] + rng.normal(0, 0.5, size=(nq, 33)).astype(np.float32)
    os.environ["OPE_FEATURE_KNN"] = "gemm"
    ms = []
    for _ in range(4):
        ctx.feature_knn(ft, fq, 5)
        ms.append(ctx.last_kernel_ms(2))
    os.environ.pop("OPE_FEATURE_KNN", None)
    k_ms = float(np.median(ms[1:]))
    g, f = ctx.feature_knn_stats()
    alg = 2.0 * 33 * nq * nt / (k_ms * 1e-3) / 1e12
    issued = 2.0 * 128 * nq * nt / (k_ms * 1e-3) / 1e12
    return {"workload": "C3-scale feature k-NN: %d x %d FPFH descriptors, k = 5" % (nq, nt), "kernel": "featgemm_kernel", "kernel_ms": k_ms,
            "roofline": {"bound": "tensor", "achieved": alg, "peak": peak, "unit": "TFLOP/s", "frac": alg / peak, "traffic": None,
                         "note": "achieved = 2*33*Nq*Nt algorithmic flops; the kernel issues K = 128 (3-way bf16 split + norm columns), "
                                 "i.e. %.0f TFLOP/s of tensor work; ncu sm__pipe_tensor_cycles_active in profiles/" % issued},
            "exact_fallback_queries": f, "gemm_queries": g}


def depth_numbers(ctx, synth, model, stream, torch, frames=1024):
    """The one stage whose inputs exceed the 126 MB L2 (only such launches say anything about HBM): depth image -> cloud
    for a batch of Kinect frames resident in HBM (8f-1), one launch of depth_fused_kernel; roofline against the measured copy peak."""
    base = []
    for f in range(8):
        _, cloud, _ = synth.make_frame(model, 700 + f)
        base.append(np.rint(np.nan_to_num(cloud[..., 2], nan=0.0).astype(np.float64) * 1000.0).astype(np.uint16))
    depth = torch.from_numpy(np.stack(base).astype(np.int16)).cuda().repeat(frames // 8, 1, 1).contiguous()
    B, R, Cc = depth.shape
    out = torch.empty((B * R * Cc, 4), dtype=torch.float32, device="cuda")
    col = torch.empty(B * Cc + 1, dtype=torch.int32, device="cuda")
    ms = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.depth_to_cloud_batch(depth.data_ptr(), B, R, Cc, out.data_ptr(), col.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    kept = int(col[-1].item())
    t = float(np.median(ms[2:])) * 1e-3
    peak, peak_src = peaks()
    npx = B * R * Cc
    alg = (2 * npx + 16 * kept) / t / 1e9
    del depth, out, col
    torch.cuda.empty_cache()
    return {"workload": "depth image -> cloud, %d frames of %dx%d resident in HBM (%.0f MB in, %.0f MB out), one launch" % (B, Cc, R, 2 * npx / 1e6, 16 * kept / 1e6),
            "kernel": "depth_fused_kernel", "ms": t * 1e3, "frames_per_sec": B / t,
            "roofline": {"bound": "hbm", "achieved": alg, "peak": peak, "unit": "GB/s", "frac": alg / peak,
                         "traffic": NCU_DEPTH_DRAM_BYTES if frames == 1024 else None,
                         "traffic_source": "ncu --set full, profiles/r01_e_depth_fused_full.txt", "peak_source": peak_src,
                         "note": "algorithmic = 2 B per pixel in + 16 B per kept pixel out; ncu dram__bytes of the launch equal it (profiles/)"}}


def pipeline_numbers(ctx, cuda_lib, synth, model, stream, torch, n_gpu_frames=12, n_cpu_frames=2):
    """C1: frames/s of estimateFinalPose (UniformSampling + normals + FPFH + SAC-IA + ICP + dense Umeyama) on synthetic
    640x480 frames; every frame starts a fresh tracker so SAC-IA runs (the first-frame path of the reference)."""
    import orc_py
    frames = [synth.make_frame(model, f)[0] for f in range(n_gpu_frames)]
    tables = []
    for f in range(n_gpu_frames):
        orc_py.srand(1 + f)
        sp = model[orc_py.uniform_sample(model, 0.01)]
        tables.append(cuda_lib.rng_table(*orc_py.sacia_draw(sp, 400, 5, 5, 0.01)))
    # device-resident
    model_cloud = ctx.upload(model)
    targets = [ctx.upload(c) for c in frames]
    def one(f):
        tr = cuda_lib.PoseTracker(ctx)
        src = ctx.transform(model_cloud, np.eye(4, dtype=np.float32))
        res = tr.estimate_final_device(src, targets[f], tables[f])
        ms = tr.stage_ms()
        tr.close(); src.free()
        return res, ms
    for f in range(3):
        one(f)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    stage = np.zeros(8)
    for f in range(n_gpu_frames):
        _, ms = one(f)
        stage += np.array(ms)
    e1.record(stream)
    torch.cuda.synchronize()
    dev_fps = n_gpu_frames / (e0.elapsed_time(e1) / 1e3)
    # end to end: host buffers in, pose + aligned cloud out
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for f in range(n_gpu_frames):
        tr = cuda_lib.PoseTracker(ctx)
        src = model.copy()
        tr.estimate_final(src, frames[f], tables[f])
        tr.close()
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_fps = n_gpu_frames / (e0.elapsed_time(e1) / 1e3)
    # batched (C5): ope_pose_batch, worker threads with their own streams, the frame-invariant model side cached
    n_batch, workers = 512, 16
    batch = {}
    for host in (False, True):
        inputs = [(frames if host else targets)[f % n_gpu_frames] for f in range(n_batch)]
        tb = [tables[f % n_gpu_frames] for f in range(n_batch)]
        ctx.pose_batch(model, inputs[:32], tables=tb[:32], workers=workers)   # warm the worker contexts
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res, status = ctx.pose_batch(model, inputs, tables=tb, workers=workers)
        e1.record(stream)
        torch.cuda.synchronize()
        assert (status == 0).all()
        batch["e2e_frames_per_sec" if host else "frames_per_sec"] = n_batch / (e0.elapsed_time(e1) / 1e3)
    batch.update({"frames": n_batch, "workers": workers, "api": "ope_pose_batch"})
    # CPU oracle, one core
    t0 = time.perf_counter()
    for f in range(n_cpu_frames):
        pe = orc_py.PoseEstimator()
        src = model.copy()
        pe.estimate_final(src, frames[f], orc_py.rng_table(*tables[f]._keep))
    cpu_fps = n_cpu_frames / (time.perf_counter() - t0)
    return {"workload": "C1: estimateFinalPose, 157825-point model vs segmented cluster of a synthetic 640x480 frame, "
                        "UniformSampling 1 cm / 8 mm, FPFH r=0.03, SAC-IA 400x5, ICP-with-normals <=100 it",
            "frames_per_sec": dev_fps, "e2e_frames_per_sec": e2e_fps, "cpu_frames_per_sec": cpu_fps, "cpu_cores": 1,
            "batched": batch,
            "stage_ms_per_frame": (stage / n_gpu_frames).round(4).tolist(),
            "stage_names": ["downsample", "normals", "fpfh", "sacia", "icp", "fitness", "umeyama+transforms", "total"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
