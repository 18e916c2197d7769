"""Committed golden fixtures (tests/golden/registration_v1.npz, made by tests/golden/make_golden.py): the CPU oracle must
still reproduce them bit for bit (CPU test), and the CUDA path must match them to the north-star tolerances (GPU test).
They freeze the oracle's behaviour; they do not pin it to PCL (the reference ships no vectors — DESIGN.md section 2)."""
import os

import numpy as np
import pytest

G_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "registration_v1.npz")
ROT_TOL, TRANS_TOL, FIT_TOL, FPFH_RTOL = 1e-4, 1e-5, 1e-5, 1e-4
ICP = dict(max_iterations=30, max_correspondence_distance=0.05, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8)
SAC = dict(max_iterations=100, nr_samples=5, k_correspondences=5, min_sample_distance=0.01, max_correspondence_distance=0.05)


@pytest.fixture(scope="module")
def G():
    return dict(np.load(G_PATH))


def _m(v):
    return np.asarray(v, np.float64).reshape(4, 4).T


# ---------------------------------------------------------------------------------------------- oracle vs golden (CPU) ----
def test_oracle_reproduces_golden_search_and_sampling(G, orc):
    for k in (1, 5, 20):
        i, d = orc.knn(G["icp_tgt"], G["icp_src"][:400], k)
        assert np.array_equal(i, G["knn%d_idx" % k]) and np.array_equal(d, G["knn%d_d2" % k])
    off, ri, rd = orc.radius(G["icp_tgt"], G["icp_src"][:200], 0.01)
    assert np.array_equal(off, G["radius_off"]) and np.array_equal(ri, G["radius_idx"]) and np.array_equal(rd, G["radius_d2"])
    assert np.array_equal(orc.uniform_sample(G["model"], 0.01), G["uniform_idx_1cm"])
    assert np.array_equal(orc.uniform_sample(G["model"], 0.008), G["uniform_idx_8mm"])
    assert np.array_equal(orc.voxel_grid(G["model"], 0.005)[0], G["voxel_5mm_xyz"])


def test_oracle_reproduces_golden_features_and_registration(G, orc):
    sp = G["model"][G["uniform_idx_1cm"]]
    n = orc.normals_knn(sp, 30)
    assert np.array_equal(n, G["normals_k30"], equal_nan=True)
    assert np.array_equal(orc.fpfh(sp, n, 0.03), G["fpfh_r3cm"], equal_nan=True)
    r = orc.icp(G["icp_src"], G["icp_tgt"], orc.icp_params(**ICP))
    assert np.array_equal(np.array(list(r.T), np.float32), G["icp_T"])
    assert [r.converged, r.state, r.iterations, r.n_correspondences] == list(G["icp_meta"])
    s, errs = orc.sacia(G["sacia_src"], G["fpfh_r3cm"], G["sacia_tgt"], G["sacia_ftgt"], orc.sacia_params(**SAC),
                        orc.rng_table(G["sacia_samples"], G["sacia_picks"]), want_errors=True)
    assert np.array_equal(errs, G["sacia_errors"]) and s.best_iteration == int(G["sacia_best"][0])
    assert np.array_equal(np.array(list(s.T), np.float32), G["sacia_T"])
    # the golden ICP really recovers the synthetic pose (an independent anchor: the known transform)
    rot, trans = 0, 0
    T = _m(G["icp_T"]); Tt = G["icp_T_true"]
    R = T[:3, :3] @ Tt[:3, :3].T
    rot = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
    assert rot < np.deg2rad(1.5) and np.linalg.norm(T[:3, 3] - Tt[:3, 3]) < 5e-3   # 1 500 noisy points, 10 % outliers


# ------------------------------------------------------------------------------------------------ CUDA vs golden (GPU) ----
@pytest.mark.gpu
def test_cuda_matches_golden_indices(G, ctx):
    ct = ctx.upload(G["icp_tgt"])
    for k in (1, 5, 20):
        i, d = ctx.knn(ct, G["icp_src"][:400], k)
        assert np.array_equal(i, G["knn%d_idx" % k]) and np.array_equal(d, G["knn%d_d2" % k])
    q = ctx.upload(G["icp_src"][:200])
    off, ri, rd = ctx.radius(ct, q, 0.01)
    assert np.array_equal(off, G["radius_off"]) and np.array_equal(ri, G["radius_idx"]) and np.array_equal(rd, G["radius_d2"])
    cm = ctx.upload(G["model"])
    assert np.array_equal(ctx.uniform_sample(cm, 0.01), G["uniform_idx_1cm"])
    assert np.array_equal(ctx.uniform_sample(cm, 0.008), G["uniform_idx_8mm"])
    vx, _ = ctx.voxel_grid(cm, 0.005)
    assert vx.shape == G["voxel_5mm_xyz"].shape and np.abs(vx - G["voxel_5mm_xyz"]).max() < 1e-6


@pytest.mark.gpu
def test_cuda_matches_golden_features(G, ctx):
    sp = G["model"][G["uniform_idx_1cm"]]
    c = ctx.upload(sp)
    n = ctx.normals_knn(c, 30)
    ref = G["normals_k30"]
    assert np.abs(n[:, :3] - ref[:, :3]).max() < 1e-4 and np.abs(n[:, 3] - ref[:, 3]).max() < 1e-5
    c2 = ctx.upload(sp, normals=ref)
    f = ctx.fpfh(c2, 0.03)
    gold = G["fpfh_r3cm"]
    assert np.abs(f - gold).max() <= FPFH_RTOL * 100.0   # histograms are normalised to 100 per sub-histogram


@pytest.mark.gpu
def test_cuda_matches_golden_registration(G, ctx, cuda_lib, synth):
    T = cuda_lib.T
    r = ctx.icp(ctx.upload(G["icp_src"]), ctx.upload(G["icp_tgt"]), cuda_lib.icp_params(**ICP))
    rot, trans = synth.pose_error(T.mat4(r.T), _m(G["icp_T"]))
    assert rot < ROT_TOL and trans < TRANS_TOL
    assert [r.converged, r.state, r.iterations, r.n_correspondences] == list(G["icp_meta"])
    fit = ctx.fitness(ctx.upload(G["icp_src"]), ctx.upload(G["icp_tgt"]), T.mat4(r.T))
    assert abs(fit - float(G["icp_fitness"][0])) < FIT_TOL
    s, errs = ctx.sacia(ctx.upload(G["sacia_src"]), G["fpfh_r3cm"], ctx.upload(G["sacia_tgt"]), G["sacia_ftgt"],
                        cuda_lib.sacia_params(**SAC), cuda_lib.rng_table(G["sacia_samples"], G["sacia_picks"]), want_errors=True)
    assert np.array_equal(errs, G["sacia_errors"]) and s.best_iteration == int(G["sacia_best"][0])
    assert np.array_equal(np.array(list(s.T), np.float32), G["sacia_T"])


@pytest.mark.gpu
def test_cuda_matches_golden_frame(G, ctx, cuda_lib, synth, orc):
    T = cuda_lib.T
    tr = cuda_lib.PoseTracker(ctx)
    src = G["model"].copy()
    orc.srand(1)   # libc rand() drives the SAC-IA draw inside the library exactly as in the reference
    p = tr.estimate_final(src, G["frame_cluster"])
    for key in ("coarse_pose", "fine_pose"):
        rot, trans = synth.pose_error(T.mat4(getattr(p, key)), _m(G["frame_" + key]))
        assert rot < ROT_TOL and trans < TRANS_TOL, key
    rot, trans = synth.pose_error(T.mat4(p.final_pose), _m(G["frame_final_pose"]))
    assert rot < 2e-4 and trans < 2e-5
    meta = [p.icp_iterations, p.icp_converged, p.icp_state, p.n_src_coarse, p.n_tgt_coarse, p.n_src_fine, p.n_tgt_fine,
            p.sacia_best_iteration]
    assert meta == list(G["frame_meta"])
    assert abs(p.fitness - G["frame_fitness"][0]) < FIT_TOL and abs(p.align_strength - G["frame_fitness"][1]) < 1e-3
    tr.close()


# ------------------------------------------------------------------ v2: point-to-plane estimators, BuildModel ICP, depth ----
G2_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "registration_v2.npz")
BM = dict(max_iterations=40, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8, k_search=20, with_normals=1)


@pytest.fixture(scope="module")
def G2():
    return dict(np.load(G2_PATH))


def _bm_params(mod, te):
    T = mod.T
    return mod.icp_params(estimator=T.EST_NORMAL_SHOOTING, rejectors=[(T.REJ_SURFACE_NORMAL, 0.7)], transformation=te, **BM)


def test_oracle_reproduces_golden_v2(G2, orc):
    T = orc.T
    sp, tp, sn, tn = G2["view0"], G2["view1"], G2["view0_normals"], G2["view1_normals"]
    assert np.array_equal(orc.normals_knn(sp, 12), sn, equal_nan=True) and np.array_equal(orc.normals_knn(tp, 12), tn, equal_nan=True)
    lls = orc.point_to_plane(sp, tp, tn, G2["pairs_src"], G2["pairs_tgt"], kind=T.TE_POINT_TO_PLANE_LLS)
    lm, info = orc.point_to_plane(sp, tp, tn, G2["pairs_src"], G2["pairs_tgt"], kind=T.TE_POINT_TO_PLANE, want_info=True)
    assert np.array_equal(lls.astype(np.float32), G2["p2p_lls_T"]) and np.array_equal(lm.astype(np.float32), G2["p2p_lm_T"])
    assert list(info) == list(G2["p2p_lm_info"])
    for name, te in (("lm", T.TE_POINT_TO_PLANE), ("lls", T.TE_POINT_TO_PLANE_LLS)):
        r = orc.icp(sp, tp, _bm_params(orc, te), src_normals=sn, tgt_normals=tn)
        assert np.array_equal(np.array(list(r.T), np.float32), G2["icp_%s_T" % name])
        assert [r.converged, r.state, r.iterations, r.n_correspondences] == list(G2["icp_%s_meta" % name])
    assert np.array_equal(orc.depth_to_cloud(G2["depth"]), G2["depth_cloud"])


@pytest.mark.gpu
def test_cuda_matches_golden_v2(G2, ctx, cuda_lib, synth):
    T = cuda_lib.T
    sp, tp, sn, tn = G2["view0"], G2["view1"], G2["view0_normals"], G2["view1_normals"]
    cs, ct = ctx.upload(sp, sn), ctx.upload(tp, tn)
    for kind, key in ((T.TE_POINT_TO_PLANE_LLS, "p2p_lls_T"), (T.TE_POINT_TO_PLANE, "p2p_lm_T")):
        g, info = ctx.point_to_plane(cs, ct, G2["pairs_src"], G2["pairs_tgt"], kind=kind, want_info=True)
        rot, trans = synth.pose_error(g, G2[key].astype(np.float64))
        assert rot < ROT_TOL and trans < TRANS_TOL, (key, rot, trans)
        if kind == T.TE_POINT_TO_PLANE:
            assert list(info) == list(G2["p2p_lm_info"])
    for name, te in (("lm", T.TE_POINT_TO_PLANE), ("lls", T.TE_POINT_TO_PLANE_LLS)):
        r = ctx.icp(cs, ct, _bm_params(cuda_lib, te))
        rot, trans = synth.pose_error(T.mat4(r.T), _m(G2["icp_%s_T" % name]))
        assert rot < ROT_TOL and trans < TRANS_TOL, (name, rot, trans)
        assert [r.converged, r.state, r.iterations, r.n_correspondences] == list(G2["icp_%s_meta" % name])
    got = ctx.depth_to_cloud(G2["depth"])
    assert np.array_equal(got.download(), G2["depth_cloud"])
