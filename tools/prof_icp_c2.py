#!/usr/bin/env python
"""C2 (50k x 50k point-to-point ICP, 50 forced iterations) with the kernel's own phase profile (OPE_PROFILE=1, OPE_PROFILE_ITER=1):
block 0's cycles per phase and iteration, far-queue statistics.   python tools/prof_icp_c2.py"""
import os, sys
os.environ["OPE_PROFILE"] = "1"; os.environ["OPE_PROFILE_ITER"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench
import ope_pkg; ope_pkg.load()
from ope_b200 import cuda_lib, synth
ctx = cuda_lib.Context(0)
model = synth.bundled_model()
src, tgt, _ = synth.icp_pair(bench.N_PTS, seed=0, model=model)
prm = cuda_lib.icp_params(**bench.icp_kwargs())
cs, ct = ctx.upload(src), ctx.upload(tgt)
for rep in range(2):
    ctx.invalidate(ct)
    r = ctx.icp(cs, ct, prm)
    print("rep", rep, "iterations", r.iterations, "kernel ms", ctx.last_kernel_ms(0), file=sys.stderr)
