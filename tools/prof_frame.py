"""Profile driver for the DetectAndLocalize frame path (C1): a few frames through the device-resident tracker, printing the
per-stage milliseconds; with OPE_PROFILE=1 the ICP kernel prints its per-phase cycles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes
import numpy as np
import ope_pkg
ope_pkg.load()
from ope_b200 import cuda_lib, synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = cuda_lib.Context(0)
model = synth.make_model()
mc = ctx.upload(model)
libc = ctypes.CDLL(None)
for f in range(reps):
    cl, _, _ = synth.make_frame(model, f)
    tc = ctx.upload(cl)
    tr = cuda_lib.PoseTracker(ctx)
    src = ctx.transform(mc, np.eye(4, dtype=np.float32))
    libc.srand(1)
    r = tr.estimate_final_device(src, tc)
    ms = tr.stage_ms()
    print("frame %d: cluster %d pts, coarse %d/%d fine %d/%d, icp iters %d state %d | ms: down %.3f normals %.3f fpfh %.3f sacia %.3f icp %.3f fit %.3f rest %.3f total %.3f"
          % (f, len(cl), r.n_src_coarse, r.n_tgt_coarse, r.n_src_fine, r.n_tgt_fine, r.icp_iterations, r.icp_state, *ms))
    tr.close(); src.free(); tc.free()
ctx.close()
