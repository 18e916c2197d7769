#!/usr/bin/env python
"""C2 ICP kernel time for every library build under build_variants/ (experiments with compile-time constants): each one is copied
over the package's libope_cuda.so in THIS checkout and timed in a fresh process; the final transform must not change.
   python tools/c2_variants.py"""
import glob, os, shutil, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "object-pose-estimation_b200", "libope_cuda.so")
CHILD = r"""
import sys, os, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "oracle"))
import bench, numpy as np
import ope_pkg; ope_pkg.load()
from ope_b200 import cuda_lib, synth
ctx = cuda_lib.Context(0)
model = synth.bundled_model()
src, tgt, _ = synth.icp_pair(bench.N_PTS, seed=0, model=model)
prm = cuda_lib.icp_params(**bench.icp_kwargs())
cs, ct = ctx.upload(src), ctx.upload(tgt)
ms = []
for rep in range(8):
    ctx.invalidate(ct)
    r = ctx.icp(cs, ct, prm)
    ms.append(ctx.last_kernel_ms(0))
print(json.dumps({"kernel_ms": float(np.median(ms[2:])), "T": [float(v) for v in r.T], "it": r.iterations, "n": r.n_correspondences}))
""" % (ROOT, ROOT)
keep = LIB + ".keep"
shutil.copy(LIB, keep)
ref = None
try:
    for so in sorted(glob.glob(os.path.join(ROOT, "build_variants", "lib_*.so"))):
        shutil.copy(so, LIB)
        out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=300)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(os.path.basename(so), "FAILED", out.stderr[-300:]); continue
        d = json.loads(line[-1])
        if ref is None:
            ref = (d["T"], d["it"], d["n"])
        print(os.path.basename(so), "kernel ms %.3f" % d["kernel_ms"], "it/s %.0f" % (50e3 / d["kernel_ms"]), "same result:", (d["T"], d["it"], d["n"]) == ref)
finally:
    shutil.copy(keep, LIB); os.remove(keep)
