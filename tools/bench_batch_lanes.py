#!/usr/bin/env python
"""ope_pose_batch: end-to-end frames/s over the number of concurrent chunk lanes and the chunk size.
   python tools/bench_batch_lanes.py [--frames 1024] [--configs "1:296,2:0,2:148,3:0"]   (chunk 0 = the library's own choice)"""
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import json, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import bench

def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default

n = int(arg("--frames", "1024"))
configs = [tuple(int(v) for v in c.split(":")) for c in arg("--configs", "1:0,2:0,3:0").split(",")]
clusters = bench.make_clusters(range(n))
import ope_pkg
ope_pkg.load()
from ope_b200 import cuda_lib, synth
import ctypes
libc = ctypes.CDLL(None)
ctx = cuda_lib.Context(0)
model = synth.bundled_model()
ref = None
out = {"frames": n, "runs": []}
for lanes, chunk in configs:
    os.environ["OPE_BATCH_LANES"] = str(lanes)
    if chunk > 0:
        os.environ["OPE_BATCH_CHUNK"] = str(chunk)
    else:
        os.environ.pop("OPE_BATCH_CHUNK", None)
    libc.srand(5)
    ctx.pose_batch(model, clusters[:min(n, 128)], workers=16)
    best = 1e9
    for rep in range(3):
        libc.srand(5)
        t0 = time.perf_counter()
        r, st = ctx.pose_batch(model, clusters, workers=16)
        best = min(best, time.perf_counter() - t0)
    assert (st == 0).all()
    key = [(tuple(x.final_pose), x.icp_iterations, x.icp_state, x.sacia_best_iteration, x.fitness) for x in r]
    if ref is None:
        ref = key
    out["runs"].append({"lanes": lanes, "chunk": chunk, "frames_per_s": round(n / best, 1), "identical_to_first_config": key == ref})
print(json.dumps(out))
