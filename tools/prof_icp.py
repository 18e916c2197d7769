"""Small driver for profiling: C2 ICP (50k/50k, 50 forced iterations) a few times; prints kernel ms."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ope_pkg
ope_pkg.load()
from ope_b200 import cuda_lib, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = cuda_lib.Context(0)
src, tgt, _ = synth.icp_pair(n, seed=0)
prm = cuda_lib.icp_params(max_iterations=iters, max_correspondence_distance=0.05, transformation_epsilon=1e-8,
                          euclidean_fitness_epsilon=1e-8, force_all_iterations=1)
cs, ct = ctx.upload(src), ctx.upload(tgt)
for r in range(reps):
    res = ctx.icp(cs, ct, prm)
    print("rep", r, "iterations", res.iterations, "kernel_ms", ctx.last_kernel_ms(0), "n_corr", res.n_correspondences)
cs.free(); ct.free(); ctx.close()
