#!/usr/bin/env python
"""C3 (BASELINE.json configs[2]): full-resolution FPFH + SAC-IA on a 307 200-point organised Kinect-like scene with a hand
occluder and 30 % outliers. Source = the model down-sampled at 5 mm; target = every valid scene point (no down-sampling).

  python tools/c3_fullres.py [--check]     --check also runs the CPU oracle on the same inputs (minutes) and compares

Prints one JSON object: per-stage device milliseconds (wall clock around synchronous C-ABI calls, data resident), sizes,
SAC-IA result, and with --check the parity verdicts (normals, FPFH within 1e-4 relative of the histogram scale, identical
per-hypothesis SAC-IA errors and winner)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ope_pkg
ope_pkg.load()
from ope_b200 import cuda_lib, synth

check = "--check" in sys.argv
model = synth.make_model()
_, cloud, pose = synth.make_frame(model, 33, outlier_frac=0.30, hand=True)
scene = cloud.reshape(-1, 3)
valid = np.isfinite(scene).all(1)
ctx = cuda_lib.Context(0)
T = {}
def timed(name, fn):
    ctx.synchronize(); t0 = time.perf_counter(); r = fn(); ctx.synchronize(); T[name] = (time.perf_counter() - t0) * 1e3; return r

cs = ctx.upload(scene)                                   # 307 200 points, NaN where the depth is invalid
cm = ctx.upload(model)
src_c = timed("model_uniform_5mm", lambda: ctx.uniform_sample_cloud(cm, 0.005))
timed("model_normals_k30", lambda: ctx.normals_knn(src_c, 30))
fsrc = timed("model_fpfh", lambda: ctx.fpfh(src_c, 0.03))
sn = timed("scene_normals_k30", lambda: ctx.normals_knn(cs, 30))
ftgt = timed("scene_fpfh_r3cm", lambda: ctx.fpfh(cs, 0.03))
sp = src_c.download()
H = 400
kw = dict(max_iterations=H, nr_samples=5, k_correspondences=5, min_sample_distance=0.01, max_correspondence_distance=0.05)
import ctypes
ctypes.CDLL(None).srand(1)
samples, picks = cuda_lib.sacia_draw(sp, H, 5, 5, 0.01)
table = cuda_lib.rng_table(samples, picks)
res, errs = timed("sacia_400x%d_vs_%d" % (len(sp), len(scene)), lambda: ctx.sacia(src_c, fsrc, cs, ftgt, cuda_lib.sacia_params(**kw), table, want_errors=True))
out = {"workload": "C3: full-resolution FPFH + SAC-IA, 640x480 organised scene (hand occluder, 30 % outliers)",
       "scene_points": int(len(scene)), "scene_valid": int(valid.sum()), "source_points": int(len(sp)),
       "stage_ms": {k: round(v, 3) for k, v in T.items()}, "sacia_kernel_ms": ctx.last_kernel_ms(1), "featgemm_kernel_ms": ctx.last_kernel_ms(2),
       "sacia_best_iteration": res.best_iteration, "sacia_best_error": res.best_error,
       "feature_gemm_queries_fallbacks": ctx.feature_knn_stats()}
rot, tr = synth.pose_error(cuda_lib.T.mat4(res.T), pose)
out["coarse_pose_error_vs_truth"] = {"rot_deg": float(np.rad2deg(rot)), "trans_m": tr}
if check:
    import orc_py as orc
    t0 = time.perf_counter()
    on = orc.normals_knn(scene, 30)
    out["cpu_scene_normals_s"] = time.perf_counter() - t0
    both = np.isfinite(on[:, 0]) & np.isfinite(sn[:, 0])
    out["normals_nan_pattern_equal"] = bool((np.isfinite(on[:, 0]) == np.isfinite(sn[:, 0])).all())
    out["normals_max_abs_diff"] = float(np.abs(on[both, :3] - sn[both, :3]).max())
    t0 = time.perf_counter()
    of = orc.fpfh(scene, on, 0.03)
    out["cpu_scene_fpfh_s"] = time.perf_counter() - t0
    gf = ctx.fpfh(ctx.upload(scene, normals=on), 0.03)    # same normals on both sides: isolates the FPFH stage
    fin = np.isfinite(of).all(1) & np.isfinite(gf).all(1)
    out["fpfh_nan_pattern_equal"] = bool((np.isfinite(of).all(1) == np.isfinite(gf).all(1)).all())
    out["fpfh_max_abs_diff_of_100"] = float(np.abs(of[fin] - gf[fin]).max())
    out["fpfh_frac_points_within_1e-4_rel"] = float((np.abs(of[fin] - gf[fin]).max(1) <= 1e-4 * 100).mean())
    osn = orc.normals_knn(sp, 30)
    osf = orc.fpfh(sp, osn, 0.03)
    t0 = time.perf_counter()
    o, oe = orc.sacia(sp, osf, scene, of, orc.sacia_params(**kw), orc.rng_table(samples, picks), want_errors=True)
    out["cpu_sacia_s"] = time.perf_counter() - t0
    g2, ge2 = ctx.sacia(src_c, osf, cs, of, cuda_lib.sacia_params(**kw), table, want_errors=True)   # oracle features on both sides
    out["sacia_errors_bit_identical"] = bool(np.array_equal(ge2, oe))
    out["sacia_same_winner"] = bool(g2.best_iteration == o.best_iteration and np.array_equal(np.array(list(g2.T)), np.array(list(o.T))))
print(json.dumps(out))
