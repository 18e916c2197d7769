#!/usr/bin/env python
"""Regenerates tests/golden/registration_v2.npz: seeded inputs and the CPU oracle's outputs for the stages added after v1 —
point-to-plane transformation estimation (LLS and Levenberg-Marquardt), the BuildModel ICP configuration that uses it, and the
depth image -> cloud conversion. Same caveat as make_golden.py: these freeze the ORACLE's behaviour (the reference ships no
vectors and cannot be built here); registration_v1.npz is left untouched.

    python tests/golden/make_golden_v2.py      # rewrites registration_v2.npz (commit the result)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import ope_pkg
    ope_pkg.load()
    from ope_b200 import synth
    import orc_py as orc
    T = orc.T

    model = synth.make_model(6000, seed=21)
    views = synth.turntable_views(model, n_views=36, first=2)
    (sp, A), (tp, B) = views
    sp, tp = sp[::3].copy(), tp[::3].copy()          # keep the fixture small
    sn, tn = orc.normals_knn(sp, 12), orc.normals_knn(tp, 12)
    g = {"view0": sp, "view1": tp, "view0_normals": sn, "view1_normals": tn, "rel_pose_true": (B @ np.linalg.inv(A)).astype(np.float64)}
    # explicit pairs: nearest neighbours of a subset
    rng = np.random.default_rng(3)
    isrc = rng.permutation(len(sp))[:700].astype(np.int32)
    itgt = np.asarray(orc.knn(tp, sp[isrc], 1)[0]).reshape(-1).astype(np.int32)
    g["pairs_src"], g["pairs_tgt"] = isrc, itgt
    lls = orc.point_to_plane(sp, tp, tn, isrc, itgt, kind=T.TE_POINT_TO_PLANE_LLS)
    lm, info = orc.point_to_plane(sp, tp, tn, isrc, itgt, kind=T.TE_POINT_TO_PLANE, want_info=True)
    g["p2p_lls_T"], g["p2p_lm_T"], g["p2p_lm_info"] = lls.astype(np.float32), lm.astype(np.float32), np.array(info, np.int64)
    # BuildModel's getIcpNormal configuration (BM/src/regmeshpcd.cpp:104-208)
    for name, te in (("lm", T.TE_POINT_TO_PLANE), ("lls", T.TE_POINT_TO_PLANE_LLS)):
        kw = dict(max_iterations=40, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8, estimator=T.EST_NORMAL_SHOOTING,
                  k_search=20, rejectors=[(T.REJ_SURFACE_NORMAL, 0.7)], with_normals=1, transformation=te)
        r = orc.icp(sp, tp, orc.icp_params(**kw), src_normals=sn, tgt_normals=tn)
        g["icp_%s_T" % name] = np.array(list(r.T), np.float32)
        g["icp_%s_meta" % name] = np.array([r.converged, r.state, r.iterations, r.n_correspondences], np.int64)
    # depth image -> cloud (D&L/src/datagrabber.cpp:9-62,121-174)
    depth = rng.integers(0, 2600, size=(96, 128)).astype(np.uint16)
    depth[10:20, 5:40] = 0
    g["depth"] = depth
    g["depth_cloud"] = orc.depth_to_cloud(depth)
    out = os.path.join(HERE, "registration_v2.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes;", len(g), "arrays")


if __name__ == "__main__":
    main()
