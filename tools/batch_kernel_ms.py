#!/usr/bin/env python
"""Pure kernel time of the batched ICP and SAC-IA scoring launches (CUDA events around the launch) next to the per-stage device
and host times of one chunk (OPE_BATCH_TRACE).   python tools/batch_kernel_ms.py [--frames 296]"""
import os, sys
os.environ["OPE_BATCH_LANES"] = "1"
n = int(sys.argv[sys.argv.index("--frames") + 1]) if "--frames" in sys.argv else 296
os.environ["OPE_BATCH_CHUNK"] = str(n)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench
clusters = bench.make_clusters(range(n))
import ope_pkg; ope_pkg.load()
from ope_b200 import cuda_lib, synth
import ctypes
libc = ctypes.CDLL(None)
ctx = cuda_lib.Context(0); model = synth.bundled_model()
for rep in range(2):
    libc.srand(5); ctx.pose_batch(model, clusters, workers=16)
    print("rep", rep, "icp kernel ms/frame", ctx.last_kernel_ms(0) / n, "sacia score kernel ms/frame", ctx.last_kernel_ms(1) / n)
os.environ["OPE_BATCH_TRACE"] = "1"
ctx.batch_stage_ms(1)
libc.srand(5); ctx.pose_batch(model, clusters, workers=16)
print({k: round(v / n, 5) for k, v in ctx.batch_stage_ms(0).items()})
