#!/usr/bin/env python
"""HBM roofline data point (only launches whose inputs exceed the 126 MB L2 say anything about HBM):
depth image -> cloud for a batch of 1 024 Kinect frames resident in HBM (629 MB of uint16 in, up to 5 GB of float4 out), one
launch per pass (count, scan, write). Reports the algorithmic GB/s (2 B per pixel in + 16 B per kept pixel out) and the
actual traffic model (the depth image is read twice) against the measured copy bandwidth in MEASURED_PEAKS.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ope_pkg
ope_pkg.load()
from ope_b200 import cuda_lib, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
model = synth.make_model(40000)
base = []
for f in range(8):
    _, cloud, _ = synth.make_frame(model, 700 + f)
    base.append(np.rint(np.nan_to_num(cloud[..., 2], nan=0.0).astype(np.float64) * 1000.0).astype(np.uint16))
base = np.stack(base)
depth = torch.from_numpy(base.astype(np.int16)).cuda().repeat(B // 8, 1, 1).contiguous()
Bf, R, Cc = depth.shape
out = torch.empty((Bf * R * Cc, 4), dtype=torch.float32, device="cuda")
col = torch.empty(Bf * Cc + 1, dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream()   # a real stream handle: the legacy default stream (handle 0) would make the ctx create its own
torch.cuda.synchronize()
ctx = cuda_lib.Context(0, stream.cuda_stream)
ms = []
for rep in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.depth_to_cloud_batch(depth.data_ptr(), Bf, R, Cc, out.data_ptr(), col.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
kept = int(col[-1].item())
t = float(np.median(ms[2:])) * 1e-3
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
npx = Bf * R * Cc
alg = (2 * npx + 16 * kept) / t / 1e9
two_pass = bool(os.environ.get("OPE_DEPTH_TWO_PASS"))
actual = ((4 if two_pass else 2) * npx + 16 * kept + (12 if two_pass else 4) * Bf * Cc) / t / 1e9
print(json.dumps({"workload": "depth->cloud, %d frames of %dx%d (inputs %.0f MB, outputs %.0f MB: far larger than L2)" % (Bf, Cc, R, 2 * npx / 1e6, 16 * kept / 1e6),
                  "frames_per_s": Bf / t, "ms": t * 1e3, "kept_fraction": kept / npx,
                  "roofline": {"bound": "hbm", "achieved": alg, "peak": peak, "unit": "GB/s", "frac": alg / peak,
                               "traffic_model_GBps": actual, "traffic_model_frac": actual / peak,
                               "kernel": "depth_count_kernel + depth_write_kernel" if two_pass else "depth_fused_kernel",
                               "note": "algorithmic = 2 B/pixel in + 16 B/kept pixel out; " + ("two passes: the depth is read twice" if two_pass else "single pass with decoupled look-back: traffic = algorithmic + 4 B per column")}}))
