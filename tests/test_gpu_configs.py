"""GPU parity at the sizes BASELINE.json's configs name (C1-C5), against the CPU oracle on the same seeded inputs, with the
bars of the north star: indices bit-exact, FPFH within 1e-4 relative, poses within 1e-4 rad / 1e-5 m with the same convergence
state and iteration count, fitness within 1e-5. The model is the cloud the reference ships (tests/golden/drill_model.npz =
D&L/3DModel/drillNewModelOrigin.pcd, 157 825 points)."""
import ctypes
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-4    # rad
TRANS_TOL = 1e-5  # m
FIT_TOL = 1e-5
FPFH_RTOL = 1e-4


@pytest.fixture(scope="module")
def drill(synth):
    m = synth.bundled_model()
    assert m.shape == (157825, 3)
    return m


# ------------------------------------------------------------------------------ a1-a5 on the bundled model (C1 inputs) ----
def test_bundled_model_downsampling_counts_and_indices(ctx, orc, drill):
    """drillNewModelOrigin.pcd -> 982 / 1 669 / 5 170 points at 1 cm / 8 mm / 5 mm; device indices == oracle."""
    c = ctx.upload(drill)
    for leaf, count in ((0.01, 982), (0.008, 1669), (0.005, 5170)):
        g = ctx.uniform_sample(c, leaf)
        o = orc.uniform_sample(drill, leaf)
        assert len(o) == count, (leaf, len(o))
        assert len(g) == len(o) and (g == o).all()


def test_bundled_model_features_and_sacia(ctx, orc, synth, cuda_lib, drill):
    cl, _, _ = synth.make_frame(drill, 1001)
    sp = drill[orc.uniform_sample(drill, 0.01)]
    tp = cl[orc.uniform_sample(cl, 0.01)]
    cs, ct = ctx.upload(sp), ctx.upload(tp)
    gsn, gtn = ctx.normals_knn(cs, 30), ctx.normals_knn(ct, 30)
    sn, tn = orc.normals_knn(sp, 30), orc.normals_knn(tp, 30)
    for g, o in ((gsn, sn), (gtn, tn)):
        assert np.array_equal(np.isfinite(g), np.isfinite(o))
        assert np.abs(g - o)[np.isfinite(o)].max() < 2e-6
    cs2, ct2 = ctx.upload(sp, sn), ctx.upload(tp, tn)
    gsf, gtf = ctx.fpfh(cs2, 0.03), ctx.fpfh(ct2, 0.03)
    sf, tf = orc.fpfh(sp, sn, 0.03), orc.fpfh(tp, tn, 0.03)
    for g, o in ((gsf, sf), (gtf, tf)):
        assert (np.abs(g - o) / np.maximum(np.abs(o), 1.0)).max() < FPFH_RTOL
    kw = dict(max_iterations=400, nr_samples=5, k_correspondences=5, min_sample_distance=0.01, max_correspondence_distance=0.05)
    orc.srand(1)
    samples, picks = orc.sacia_draw(sp, 400, 5, 5, 0.01)
    o, oe = orc.sacia(sp, sf, tp, tf, orc.sacia_params(**kw), orc.rng_table(samples, picks), want_errors=True)
    g, ge = ctx.sacia(cs, sf, ct, tf, cuda_lib.sacia_params(**kw), cuda_lib.rng_table(samples, picks), want_errors=True)
    assert np.array_equal(ge, oe)
    assert g.best_iteration == o.best_iteration and np.array_equal(np.array(list(g.T)), np.array(list(o.T)))


# ----------------------------------------------------------------------------------------------------------- C2 ----
def test_c2_icp_50k_vs_50k_50_iterations(ctx, orc, synth, cuda_lib, drill):
    """BASELINE.json configs[1] at full size: 50 000 vs 50 000 points, max-corr-distance 0.05, 50 iterations — once with the
    convergence criteria acting (the parity run), once forced to all 50 iterations (what bench.py times)."""
    T = cuda_lib.T
    src, tgt, _ = synth.icp_pair(50000, seed=0, model=drill)
    cs, ct = ctx.upload(src), ctx.upload(tgt)
    for force in (0, 1):
        kw = dict(max_iterations=50, max_correspondence_distance=0.05, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8,
                  force_all_iterations=force)
        g, gc = ctx.icp(cs, ct, cuda_lib.icp_params(**kw), want_corr=True)
        o, oc = orc.icp(src, tgt, orc.icp_params(**kw), want_corr=True)
        r, t = synth.pose_error(T.mat4(g.T), T.mat4(o.T))
        assert r < ROT_TOL and t < TRANS_TOL, (force, r, t)
        assert (g.converged, g.state, g.iterations, g.n_correspondences) == (o.converged, o.state, o.iterations, o.n_correspondences)
        assert (gc[0] == oc[0]).all() and (gc[1] == oc[1]).all() and (gc[2] == oc[2]).all()
        gf, of = ctx.fitness(cs, ct, T.mat4(g.T)), orc.fitness(src, tgt, T.mat4(o.T))
        assert abs(gf - of) < FIT_TOL
        if force:
            assert g.iterations == 50


# ----------------------------------------------------------------------------------------------------------- C3 ----
def test_c3_full_resolution_scene_normals_fpfh_sacia(ctx, orc, synth, cuda_lib, drill):
    """BASELINE.json configs[2]: 307 200-point organised scene (hand occluder, 30 % outliers, NaN for invalid pixels); source =
    the model at 5 mm. Normals k = 30 over the whole scene; FPFH r = 3 cm with a COUNTED bin-flip budget (a pair feature that
    lands within float rounding of a bin edge may fall into the neighbouring bin: at most 1 point in 10 000 may exceed 1e-4
    relative, and none by more than two flipped pairs); SAC-IA 400 x 5 against the full scene: every hypothesis error bit-equal."""
    _, cloud, _ = synth.make_frame(drill, 33, outlier_frac=0.30, hand=True)
    scene = cloud.reshape(-1, 3)
    assert len(scene) == 307200
    cs = ctx.upload(scene)
    gn = ctx.normals_knn(cs, 30)
    on = orc.normals_knn(scene, 30)
    assert np.array_equal(np.isfinite(gn), np.isfinite(on))
    fin = np.isfinite(on)
    assert np.abs(gn[fin] - on[fin]).max() < 2e-6
    gf = ctx.fpfh(ctx.upload(scene, normals=on), 0.03)        # same normals on both sides: isolates the FPFH stage
    of = orc.fpfh(scene, on, 0.03)
    assert np.array_equal(np.isfinite(gf).all(1), np.isfinite(of).all(1))
    ok = np.isfinite(of).all(1)
    rel = (np.abs(gf[ok] - of[ok]) / np.maximum(np.abs(of[ok]), 1.0)).max(1)
    flipped = int((rel >= FPFH_RTOL).sum())
    assert flipped <= ok.sum() // 10000, flipped
    assert np.abs(gf[ok] - of[ok]).max() < 0.5                 # a flipped pair moves 100/(m-1) weighted percent between two bins
    cm = ctx.upload(drill)
    sp = drill[orc.uniform_sample(drill, 0.005)]
    sn = orc.normals_knn(sp, 30)
    sf = orc.fpfh(sp, sn, 0.03)
    kw = dict(max_iterations=400, nr_samples=5, k_correspondences=5, min_sample_distance=0.01, max_correspondence_distance=0.05)
    orc.srand(1)
    samples, picks = orc.sacia_draw(sp, 400, 5, 5, 0.01)
    o, oe = orc.sacia(sp, sf, scene, of, orc.sacia_params(**kw), orc.rng_table(samples, picks), want_errors=True)
    g, ge = ctx.sacia(ctx.upload(sp), sf, cs, of, cuda_lib.sacia_params(**kw), cuda_lib.rng_table(samples, picks), want_errors=True)
    assert np.array_equal(ge, oe)
    assert g.best_iteration == o.best_iteration and g.best_error == o.best_error
    assert np.array_equal(np.array(list(g.T)), np.array(list(o.T)))
    cm.free()


# ----------------------------------------------------------------------------------------------------------- C4 ----
def test_c4_chain_first_pairs(ctx, orc, synth, cuda_lib, drill):
    """BASELINE.json configs[3], first 6 pairs of the 36-view chain exactly as RegMeshPcd::registerPointClouds runs it
    (BM/src/regmeshpcd.cpp:210-271): normals k = 12 on the growing merged cloud and the next view, normal shooting k = 20,
    surface-normal rejector, Levenberg-Marquardt point-to-plane, `*aligned += *target` — the merged cloud stays on the device
    (ope_cloud_append). The oracle runs the same chain from its own merged cloud: errors would accumulate, not cancel."""
    T = cuda_lib.T
    views = synth.turntable_views(drill[::4].copy(), n_views=36, first=7)
    kw = dict(max_iterations=60, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8, estimator=T.EST_NORMAL_SHOOTING,
              k_search=20, rejectors=[(T.REJ_SURFACE_NORMAL, 0.7)], with_normals=1, transformation=T.TE_POINT_TO_PLANE)
    g_merged = ctx.upload(views[0][0])
    o_merged = views[0][0]
    for i in range(6):
        target = views[i + 1][0]
        ct = ctx.upload(target)
        ctx.normals_knn(g_merged, 12)
        ctx.normals_knn(ct, 12)
        g = ctx.icp(g_merged, ct, cuda_lib.icp_params(**kw))
        sn, tn = orc.normals_knn(o_merged, 12), orc.normals_knn(target, 12)
        o = orc.icp(o_merged, target, orc.icp_params(**kw), src_normals=sn, tgt_normals=tn)
        r, t = synth.pose_error(T.mat4(g.T), T.mat4(o.T))
        assert r < ROT_TOL and t < TRANS_TOL, (i, r, t)
        assert (g.converged, g.state, g.iterations, g.n_correspondences) == (o.converged, o.state, o.iterations, o.n_correspondences), i
        moved = ctx.transform(g_merged, T.mat4(g.T))
        ctx.append(moved, ct)                                # *cloudAlignedIcp += *cloudTarget, on the device
        g_merged.free(); ct.free()
        g_merged = moved
        o_merged = np.concatenate([orc.transform(o_merged, T.mat4(o.T)), target])
        assert len(g_merged) == len(o_merged)
    assert np.abs(g_merged.download() - o_merged).max() < 2e-5
    # the same chain as ONE call (ope_register_point_clouds): bit-identical to the step-by-step device chain
    merged, pairs = ctx.register_point_clouds([v[0] for v in views], cuda_lib.icp_params(**kw), normal_k=12)
    assert len(pairs) == 6 and len(merged) == len(o_merged)
    assert np.array_equal(merged.download(), g_merged.download())


# ------------------------------------------------------------------------------------------------------ C1 / C5 ----
_S = {}


def _oracle_frame(f):
    import orc_py
    if "model" not in _S:
        import ope_pkg
        ope_pkg.load()
        from ope_b200 import synth
        _S["synth"], _S["model"] = synth, synth.bundled_model()
    synth, model = _S["synth"], _S["model"]
    cl = synth.make_frame(model, 1000 + f)[0]
    orc_py.srand(1)
    src = model.copy()
    p = orc_py.PoseEstimator().estimate_final(src, cl)
    return (np.array(list(p.final_pose)), np.array(list(p.coarse_pose)), np.array(list(p.fine_pose)), p.icp_state, p.icp_converged,
            p.icp_iterations, p.fitness, p.align_strength, p.sacia_best_iteration, src)


def test_c1_c5_scene_sample_single_and_batched(ctx, orc, synth, cuda_lib, drill):
    """64 of the 1 000 agreement scenes (frames 1000..1063; the bundled model under a random pose at 0.6-1.6 m, table, wall, depth
    noise), full estimateFinalPose: the oracle on the host cores, ope_pose_estimate_final one frame at a time, and ope_pose_batch —
    all replaying the decisions a fresh PoseEstimator draws from libc rand() after srand(1)."""
    T = cuda_lib.T
    n = 64
    with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 16)) as pool:
        oracle = pool.map(_oracle_frame, range(n))
    libc = ctypes.CDLL(None)
    clusters = [synth.make_frame(drill, 1000 + f)[0] for f in range(n)]
    orc.srand(1)
    sp = drill[orc.uniform_sample(drill, 0.01)]
    table = cuda_lib.rng_table(*orc.sacia_draw(sp, 400, 5, 5, 0.01))
    bres, bstatus = ctx.pose_batch(drill, clusters, tables=[table] * n, workers=8)
    assert (bstatus == 0).all()
    for f in range(n):
        o_final, o_coarse, o_fine, o_state, o_conv, o_it, o_fit, o_strength, o_best, o_src = oracle[f]
        tr = cuda_lib.PoseTracker(ctx)
        src = drill.copy()
        libc.srand(1)
        g = tr.estimate_final(src, clusters[f])
        tr.close()
        for name, p in (("single", g), ("batch", bres[f])):
            for mine, theirs in ((p.final_pose, o_final), (p.coarse_pose, o_coarse), (p.fine_pose, o_fine)):
                r, t = synth.pose_error(T.mat4(mine), theirs.reshape(4, 4).T)
                assert r < ROT_TOL and t < TRANS_TOL, (name, f, r, t)
            assert (p.icp_state, p.icp_converged, p.icp_iterations, p.sacia_best_iteration) == (o_state, o_conv, o_it, o_best), (name, f)
            assert abs(p.fitness - o_fit) < FIT_TOL and abs(p.align_strength - o_strength) < 1e-12, (name, f)
        assert np.abs(src - o_src).max() < 2e-5, f           # alignedSource handed back in place of the source


@pytest.mark.gpu
def test_pose_batch_lanes_and_chunks_do_not_change_results(ctx, synth, cuda_lib, drill, monkeypatch):
    """ope_pose_batch cuts a batch into chunks and runs them on concurrent lanes (helper threads on worker contexts); tables drawn
    inside the call from libc rand(). Whatever the cut — one lane and one chunk, many lanes and tiny chunks, sleeping or spinning
    waits — every frame's result is the same, bit for bit, and the decision tables come from the stream in frame order."""
    libc = ctypes.CDLL(None)
    n = 40
    clusters = [synth.make_frame(drill, 2000 + f)[0] for f in range(n)]

    def run(lanes, chunk, sync=None):
        monkeypatch.setenv("OPE_BATCH_LANES", str(lanes))
        monkeypatch.setenv("OPE_BATCH_CHUNK", str(chunk))
        if sync:
            monkeypatch.setenv("OPE_BATCH_LANE_SYNC", sync)
        else:
            monkeypatch.delenv("OPE_BATCH_LANE_SYNC", raising=False)
        libc.srand(11)
        res, status = ctx.pose_batch(drill, clusters, tables=None, workers=8)
        assert (status == 0).all()
        tail = libc.rand()   # where the stream stands after the call
        return [(tuple(r.final_pose), tuple(r.coarse_pose), tuple(r.fine_pose), r.icp_iterations, r.icp_state, r.icp_converged,
                 r.sacia_best_iteration, r.sacia_best_error, r.fitness, r.align_strength, r.n_src_fine, r.n_tgt_fine) for r in res], tail

    ref, tail = run(1, 64)
    for lanes, chunk, sync in ((2, 20, None), (4, 7, "block"), (8, 5, "yield"), (3, 13, "spin")):
        got, t = run(lanes, chunk, sync)
        assert got == ref, (lanes, chunk, sync)
        assert t == tail, (lanes, chunk, sync)
