#!/usr/bin/env python
"""ope_pose_batch: frame-spanning launches (default) vs the per-frame worker path (OPE_BATCH_MODE=workers), same frames, same
decision tables: frames/s of both and bit-level agreement of every result.   python tools/bench_batch_modes.py [--frames 512]"""
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import json, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import bench

n = int(sys.argv[sys.argv.index("--frames") + 1]) if "--frames" in sys.argv else 512
clusters = bench.make_clusters(range(n))
import ope_pkg
ope_pkg.load()
from ope_b200 import cuda_lib, synth
import ctypes
libc = ctypes.CDLL(None)
ctx = cuda_lib.Context(0)
model = synth.bundled_model()
out = {"frames": n}
res = {}
fused_only = "--fused-only" in sys.argv
for mode in (("fused",) if fused_only else ("fused", "workers")):
    os.environ["OPE_BATCH_MODE"] = mode
    libc.srand(5)
    ctx.pose_batch(model, clusters[:64], workers=16)
    best = 1e9
    for rep in range(1 if fused_only else 3):
        libc.srand(5)
        t0 = time.perf_counter()
        r, st = ctx.pose_batch(model, clusters, workers=16)
        best = min(best, time.perf_counter() - t0)
    assert (st == 0).all()
    if mode == "fused":
        ctx.batch_stage_ms(1)
        libc.srand(5)
        ctx.pose_batch(model, clusters, workers=16)
        out["fused_stage_ms_per_frame"] = {k: round(v / n, 5) for k, v in ctx.batch_stage_ms(0).items()}
    res[mode] = r
    out[mode + "_e2e_frames_per_s"] = n / best
    out[mode + "_launches_per_frame"] = None
if fused_only:
    print(json.dumps(out))
    sys.exit(0)
same = 0
for a, b in zip(res["fused"], res["workers"]):
    same += (np.array_equal(np.array(list(a.final_pose)), np.array(list(b.final_pose))) and a.icp_iterations == b.icp_iterations and
             a.icp_state == b.icp_state and a.sacia_best_iteration == b.sacia_best_iteration and a.n_src_fine == b.n_src_fine and
             a.n_tgt_fine == b.n_tgt_fine and abs(a.fitness - b.fitness) < 1e-12 and a.align_strength == b.align_strength)
out["bit_identical_frames"] = same
T = cuda_lib.T
errs = [synth.pose_error(T.mat4(a.final_pose), T.mat4(b.final_pose)) for a, b in zip(res["fused"], res["workers"])]
out["max_rot_rad"] = max(e[0] for e in errs); out["max_trans_m"] = max(e[1] for e in errs)
out["same_icp_state_and_iterations"] = sum((a.icp_iterations, a.icp_state, a.icp_converged) == (b.icp_iterations, b.icp_state, b.icp_converged)
                                           for a, b in zip(res["fused"], res["workers"]))
print(json.dumps(out))
