#!/usr/bin/env python
"""C4 (BASELINE.json configs[3]): BuildModel multi-view reconstruction — 36 synthetic turntable views chained by pairwise
ICP-with-normals into one merged model cloud, exactly as RegMeshPcd::registerPointClouds does it
(BM/src/regmeshpcd.cpp:210-271): for every pair, normals k = 12 of the (growing) merged cloud and of the next view, normal
shooting k = 20, surface-normal rejector, TransformationEstimationPointToPlane (Levenberg-Marquardt), eps 1e-8; then
`*aligned += *target`. The chain does not shard (pair i consumes the merged output of pair i-1): one GPU, "replicas only".

  python tools/c4_chain.py [--views 36] [--check N] [--max-iter 60] [--te lm|svd]

Prints one JSON object: per-pair and total wall time around the synchronous C-ABI calls (host buffers in, 4x4 out: end to end),
the pose error of every pair against the ground-truth relative pose, and — with --check N — the agreement of the first N pairs
with the CPU oracle run on the same inputs (same normals handed to both sides), plus the oracle's time for them."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ope_pkg
ope_pkg.load()
from ope_b200 import cuda_lib, synth


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    n_views, n_check, max_iter, te = arg("--views", 36), arg("--check", 0), arg("--max-iter", 60), arg("--te", "lm")
    T = cuda_lib.T
    model = synth.make_model(40000)
    views = synth.turntable_views(model, n_views=36, first=n_views)
    kw = dict(max_iterations=max_iter, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8, estimator=T.EST_NORMAL_SHOOTING,
              k_search=20, rejectors=[(T.REJ_SURFACE_NORMAL, 0.7)], with_normals=1,
              transformation=T.TE_POINT_TO_PLANE if te == "lm" else T.TE_SVD)
    ctx = cuda_lib.Context(0)
    prm = cuda_lib.icp_params(**kw)
    merged = views[0][0]
    pairs = []
    total = 0.0
    orc = None
    if n_check:
        import orc_py as orc
    for i in range(len(views) - 1):
        target = views[i + 1][0]
        t0 = time.perf_counter()
        cs, ct = ctx.upload(merged), ctx.upload(target)
        sn = ctx.normals_knn(cs, 12)      # normEst.compute on the merged cloud and on the view (BM/src/regmeshpcd.cpp:74-90)
        tn = ctx.normals_knn(ct, 12)
        res = ctx.icp(cs, ct, prm)
        M = T.mat4(res.T)
        moved = ctx.transform(cs, M)      # pcl::transformPointCloud(*p_cloudSource, *cloudAligned, transformIcpNormal)
        new_merged = np.concatenate([moved.download(), target])   # *cloudAlignedIcp += *cloudTarget
        dt = time.perf_counter() - t0
        total += dt
        truth = views[i + 1][1] @ np.linalg.inv(views[i][1])
        r, t = synth.pose_error(M, truth)
        rec = {"pair": i, "source_points": len(merged), "target_points": len(target), "iterations": res.iterations, "state": res.state,
               "ms": dt * 1e3, "rot_err_deg_vs_truth": float(np.rad2deg(r)), "trans_err_mm_vs_truth": float(t * 1e3)}
        if i < n_check:
            t1 = time.perf_counter()
            o = orc.icp(merged, target, orc.icp_params(**kw), src_normals=sn, tgt_normals=tn)
            rec["oracle_ms"] = (time.perf_counter() - t1) * 1e3
            ro, to = synth.pose_error(M, T.mat4(o.T))
            rec.update({"rot_diff_vs_oracle_rad": float(ro), "trans_diff_vs_oracle_m": float(to),
                        "same_iterations_and_state": bool(o.iterations == res.iterations and o.state == res.state)})
        pairs.append(rec)
        for c in (cs, ct, moved):
            c.free()
        merged = new_merged
    out = {"workload": "C4: %d turntable views chained by pairwise ICP-with-normals (%s estimator), merged cloud grows to %d points"
                       % (len(views), "Levenberg-Marquardt point-to-plane" if te == "lm" else "SVD", len(merged)),
           "pairs": len(pairs), "total_s": total, "pairs_per_s": len(pairs) / total, "merged_points": int(len(merged)),
           "max_rot_err_deg_vs_truth": max(p["rot_err_deg_vs_truth"] for p in pairs),
           "max_trans_err_mm_vs_truth": max(p["trans_err_mm_vs_truth"] for p in pairs)}
    if n_check:
        chk = pairs[:n_check]
        out.update({"checked_pairs": len(chk), "max_rot_diff_vs_oracle_rad": max(p["rot_diff_vs_oracle_rad"] for p in chk),
                    "max_trans_diff_vs_oracle_m": max(p["trans_diff_vs_oracle_m"] for p in chk),
                    "all_same_iterations_and_state": all(p["same_iterations_and_state"] for p in chk),
                    "oracle_s_for_checked_pairs": sum(p["oracle_ms"] for p in chk) / 1e3,
                    "ours_s_for_checked_pairs": sum(p["ms"] for p in chk) / 1e3})
    out["per_pair"] = pairs
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
