#!/usr/bin/env python
"""Cost of the first ICP iterations in the frame-spanning launch: the ICP stage's device ms per frame with the iteration budget
capped at 1, 2, 3, 5, 10, 20, 100 (same frames, one lane so that the stage times are clean).   python tools/icp_iter_cost.py [--frames 296]"""
import os
os.environ["OPE_BATCH_LANES"] = "1"
import json, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench
n = int(sys.argv[sys.argv.index("--frames") + 1]) if "--frames" in sys.argv else 296
clusters = bench.make_clusters(range(n))
import ope_pkg
ope_pkg.load()
from ope_b200 import cuda_lib, abi_types as T
import ctypes
libc = ctypes.CDLL(None)
ctx = cuda_lib.Context(0)
from ope_b200 import synth
model = synth.bundled_model()
out = {}
for cap in (1, 2, 3, 5, 10, 20, 100):
    prm = cuda_lib.pose_params()
    prm.icp.max_iterations = cap
    libc.srand(5); ctx.pose_batch(model, clusters, prm=prm, workers=16)
    ctx.batch_stage_ms(1)
    libc.srand(5); r, st = ctx.pose_batch(model, clusters, prm=prm, workers=16)
    ms = ctx.batch_stage_ms(0)
    out[cap] = {"icp_ms_per_frame": round(ms["icp"] / n, 5), "mean_iterations": sum(x.icp_iterations for x in r) / n}
print(json.dumps(out))
