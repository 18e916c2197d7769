#!/usr/bin/env python
"""Regenerates tests/golden/registration_v1.npz: seeded inputs and the CPU oracle's outputs for every stage of the path.

The reference ships no golden vectors and cannot be built here (parity unpinned), so these fixtures do
NOT pin the oracle to PCL; they freeze the oracle's own behaviour so that (a) an accidental change of the oracle is caught
by tests/test_golden.py on the CPU and (b) the CUDA path is checked against committed numbers, not only against a
checker built in the same run. Inputs are stored with the outputs, so the fixtures do not depend on synth.py either.

    python tests/golden/make_golden.py      # rewrites registration_v1.npz (commit the result)
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

    model = synth.make_model(6000, seed=11)
    src, tgt, T_true = synth.icp_pair(1500, seed=12, model=model)
    g = {"model": model, "icp_src": src, "icp_tgt": tgt, "icp_T_true": T_true.astype(np.float64)}

    # search
    for k in (1, 5, 20):
        i, d = orc.knn(tgt, src[:400], k)
        g["knn%d_idx" % k], g["knn%d_d2" % k] = i, d
    off, ri, rd = orc.radius(tgt, src[:200], 0.01)
    g["radius_off"], g["radius_idx"], g["radius_d2"] = off, ri, rd
    # down-sampling
    g["uniform_idx_1cm"] = orc.uniform_sample(model, 0.01)
    g["uniform_idx_8mm"] = orc.uniform_sample(model, 0.008)
    vx, _ = orc.voxel_grid(model, 0.005)
    g["voxel_5mm_xyz"] = vx
    # features
    sp = model[g["uniform_idx_1cm"]]
    g["normals_k30"] = orc.normals_knn(sp, 30)
    g["normals_k12_model"] = orc.normals_knn(model[:2000], 12)
    g["fpfh_r3cm"] = orc.fpfh(sp, g["normals_k30"], 0.03)
    # rigid transform
    g["umeyama_T"] = np.array(orc.umeyama(src, synth.apply(T_true, src)), np.float32)
    # ICP (C2 in miniature) and fitness
    prm = dict(max_iterations=30, max_correspondence_distance=0.05, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8)
    r = orc.icp(src, tgt, orc.icp_params(**prm))
    g["icp_T"] = np.array(list(r.T), np.float32)
    g["icp_meta"] = np.array([r.converged, r.state, r.iterations, r.n_correspondences], np.int64)
    g["icp_fitness"] = np.array([orc.fitness(src, tgt, np.array(list(r.T), np.float32).reshape(4, 4).T)], np.float64)
    # SAC-IA with a replayed table
    cl, _, pose = synth.make_frame(model, 5)
    tp = cl[orc.uniform_sample(cl, 0.01)]
    tn = orc.normals_knn(tp, 30)
    tf = orc.fpfh(tp, tn, 0.03)
    orc.srand(1)
    samples, picks = orc.sacia_draw(sp, 100, 5, 5, 0.01)
    kw = dict(max_iterations=100, nr_samples=5, k_correspondences=5, min_sample_distance=0.01, max_correspondence_distance=0.05)
    s, errs = orc.sacia(sp, g["fpfh_r3cm"], tp, tf, orc.sacia_params(**kw), orc.rng_table(samples, picks), want_errors=True)
    g["sacia_src"], g["sacia_tgt"], g["sacia_ftgt"] = sp, tp, tf
    g["sacia_samples"], g["sacia_picks"], g["sacia_errors"] = samples, picks, errs
    g["sacia_T"] = np.array(list(s.T), np.float32)
    g["sacia_best"] = np.array([s.best_iteration], np.int64)
    # the frame path (estimateFinalPose), first frame
    g["frame_cluster"], g["frame_pose_true"] = cl, pose.astype(np.float64)
    pe = orc.PoseEstimator()
    msrc = model.copy()
    orc.srand(1)
    p = pe.estimate_final(msrc, cl)
    g["frame_final_pose"] = np.array(list(p.final_pose), np.float32)
    g["frame_coarse_pose"] = np.array(list(p.coarse_pose), np.float32)
    g["frame_fine_pose"] = np.array(list(p.fine_pose), np.float32)
    g["frame_meta"] = np.array([p.icp_iterations, p.icp_converged, p.icp_state, p.n_src_coarse, p.n_tgt_coarse, p.n_src_fine,
                                p.n_tgt_fine, p.sacia_best_iteration], np.int64)
    g["frame_fitness"] = np.array([p.fitness, p.align_strength], np.float64)
    out = os.path.join(HERE, "registration_v1.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes;", len(g), "arrays")


if __name__ == "__main__":
    main()
