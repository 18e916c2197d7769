"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> libope_cuda.so), against the CPU oracle
on the same seeded inputs. Bars (BASELINE.json north_star): neighbour / correspondence indices bit-exact, FPFH within
1e-4 relative, poses within 1e-4 rad / 1e-5 m with the same convergence flag and fitness within 1e-5."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-4    # rad
TRANS_TOL = 1e-5  # m
FIT_TOL = 1e-5
FPFH_RTOL = 1e-4


def _pose_close(synth, T, a, b, rot=ROT_TOL, trans=TRANS_TOL):
    r, t = synth.pose_error(T.mat4(a), T.mat4(b))
    assert r < rot and t < trans, (r, t)


# ------------------------------------------------------------------------------------------------ search ----
@pytest.mark.parametrize("k", [1, 5, 12, 20, 30])
def test_knn_indices_bit_exact_surface(ctx, orc, synth, small_model, k):
    src, tgt, _ = synth.icp_pair(6000, seed=3, model=small_model)
    ct = ctx.upload(tgt)
    gi, gd = ctx.knn(ct, src, k)
    oi, od = orc.knn(tgt, src, k)
    assert (gi == oi).all()
    assert (gd == od).all()


def test_knn_volume_far_queries_and_ties(ctx, orc):
    rng = np.random.default_rng(5)
    tgt = (np.round(rng.random((4000, 3)) * 64) / 64).astype(np.float32)      # many exact ties + duplicates
    qry = (rng.random((1500, 3)) * 3 - 1).astype(np.float32)                  # many queries far outside the box
    ct = ctx.upload(tgt)
    for k in (1, 8):
        gi, gd = ctx.knn(ct, qry, k)
        oi, od = orc.knn(tgt, qry, k, brute=True)
        assert (gi == oi).all() and (gd == od).all()


def test_knn_edge_cases(ctx, orc):
    tgt = np.array([[0, 0, 0], [1, 0, 0], [np.nan, 0, 0], [0, 2, 0]], np.float32)
    qry = np.array([[0.1, 0, 0], [5, 5, 5]], np.float32)
    ct = ctx.upload(tgt)
    gi, gd = ctx.knn(ct, qry, 5)          # k > finite points: padded with -1 / inf
    oi, od = orc.knn(tgt, qry, 5, brute=True)
    assert (gi == oi).all() and np.array_equal(gd, od)
    one = ctx.upload(tgt[:1])
    gi, _ = ctx.knn(one, qry, 1)
    assert (gi == 0).all()


def test_radius_search(ctx, orc, synth, small_model):
    pts = small_model[:5000]
    c = ctx.upload(pts)
    go, gi, gd = ctx.radius(c, c, 0.01)
    oo, oi, od = orc.radius(pts, pts, 0.01)
    assert (go == oo).all() and (gi == oi).all() and (gd == od).all()


# ---------------------------------------------------------------------------------------- down-sampling ----
@pytest.mark.parametrize("leaf", [0.01, 0.008, 0.005])
def test_uniform_sampling_indices(ctx, orc, model, leaf):
    c = ctx.upload(model)
    g = ctx.uniform_sample(c, leaf)
    o = orc.uniform_sample(model, leaf)
    assert len(g) == len(o) and (g == o).all()


def test_uniform_sampling_scene_with_nan(ctx, orc, synth, model):
    _, cloud, _ = synth.make_frame(model, 7)
    pts = cloud.reshape(-1, 3)
    assert np.isnan(pts).any() or True
    c = ctx.upload(pts)
    g = ctx.uniform_sample(c, 0.005)
    o = orc.uniform_sample(pts, 0.005)
    assert len(g) == len(o) and (g == o).all()


def test_voxel_grid_centroids(ctx, orc, synth, model):
    cl, _, _ = synth.make_frame(model, 3)
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 1 << 24, len(cl)).astype(np.uint32).view(np.float32)
    c = ctx.upload(cl)
    gx, gr = ctx.voxel_grid(c, 0.005, rgb)
    ox, orr = orc.voxel_grid(cl, 0.005, rgb)
    assert gx.shape == ox.shape
    assert np.array_equal(gx, ox)
    assert np.array_equal(gr.view(np.uint32), orr.view(np.uint32))


def test_empty_and_degenerate_clouds(ctx, cuda_lib):
    e = ctx.upload(np.zeros((0, 3), np.float32))
    assert len(ctx.uniform_sample(e, 0.01)) == 0
    same = ctx.upload(np.ones((10, 3), np.float32))
    assert len(ctx.uniform_sample(same, 0.01)) == 1
    prm = cuda_lib.icp_params()
    with pytest.raises(cuda_lib.OpeError):   # no target -> error code, transforms stay identity (no throw across the ABI)
        ctx.icp(same, e, prm)


# --------------------------------------------------------------------------------------------- features ----
@pytest.mark.parametrize("k", [12, 30])
@pytest.mark.parametrize("path", ["smem", "grid"])
def test_normals(ctx, orc, synth, model, k, path, monkeypatch):
    """small clouds are answered from shared memory without a spatial index, larger ones (or OPE_NORMALS_FORCE_GRID) through the
    Morton grid: both must give the oracle's normals, including NaN points and clouds with fewer than k points"""
    if path == "grid":
        monkeypatch.setenv("OPE_NORMALS_FORCE_GRID", "1")
    cl, _, _ = synth.make_frame(model, 11)
    idx = orc.uniform_sample(cl, 0.008)
    pts = cl[idx].copy()
    pts[5] = np.nan
    pts[77, 1] = np.inf
    for cloud in (pts, pts[:7], pts[:2], model[:6000]):
        c = ctx.upload(cloud)
        g = ctx.normals_knn(c, k)
        o = orc.normals_knn(cloud, k)
        assert np.array_equal(np.isfinite(g), np.isfinite(o))
        fin = np.isfinite(o)
        # same neighbour sets + same float order => identical up to transcendental rounding
        assert np.allclose(g[fin], o[fin], rtol=0, atol=2e-6), np.abs(g[fin] - o[fin]).max()
        assert len(cloud) < 100 or (g[fin] == o[fin]).mean() > 0.9
        c.free()


@pytest.mark.parametrize("path", ["smem", "grid"])
def test_fpfh_within_tolerance(ctx, orc, synth, model, path, monkeypatch):
    """small clouds are processed from shared memory without a spatial index (neighbours in ascending index order, like the oracle),
    larger ones (or OPE_FPFH_FORCE_GRID) through the Morton grid"""
    if path == "grid":
        monkeypatch.setenv("OPE_FPFH_FORCE_GRID", "1")
    cl, _, _ = synth.make_frame(model, 11)
    pts = cl[orc.uniform_sample(cl, 0.01)]
    assert len(pts) <= 2048
    nr = orc.normals_knn(pts, 30)
    pts = pts.copy(); pts[7] = np.nan     # a NaN point: its own histogram is NaN, nobody counts it as a neighbour
    c = ctx.upload(pts, nr)
    g, gs = ctx.fpfh(c, 0.03, want_spfh=True)
    o = orc.fpfh(pts, nr, 0.03)
    os_ = orc.spfh(pts, nr, 0.03)
    assert np.array_equal(np.isnan(g), np.isnan(o)) and np.array_equal(np.isnan(gs), np.isnan(os_))
    fin = ~np.isnan(o).any(1)
    assert np.array_equal(gs[fin], os_[fin]), np.abs(gs[fin] - os_[fin]).max()   # SPFH: integer counts + replayed float adds
    scale = np.maximum(np.abs(o[fin]), 1.0)                                      # histograms are percentages (0..100)
    assert (np.abs(g[fin] - o[fin]) / scale).max() < FPFH_RTOL
    sums = g[fin].reshape(fin.sum(), 3, 11).sum(-1)
    assert np.allclose(sums, 100.0, atol=1e-2)


def test_feature_knn_indices(ctx, orc):
    rng = np.random.default_rng(2)
    ft = (rng.random((3000, 33)) * 30).astype(np.float32)
    fq = (rng.random((700, 33)) * 30).astype(np.float32)
    ft[10] = ft[20]                                                   # exact tie -> smaller index first
    gi, gd = ctx.feature_knn(ft, fq, 5)
    oi, od = orc.feature_knn(ft, fq, 5)
    assert (gi == oi).all() and (gd == od).all()


# ------------------------------------------------------------------------------------------ registration ----
def test_umeyama_dense(ctx, orc, synth, model):
    rng = np.random.default_rng(4)
    T = synth.random_pose(rng)
    moved = synth.apply(T, model)
    a, b = ctx.upload(model), ctx.upload(moved)
    g = ctx.umeyama(a, b)
    r, t = synth.pose_error(g, T)
    assert r < 1e-5 and t < 1e-5
    o = orc.umeyama(model, moved)
    r, t = synth.pose_error(g, o)
    assert r < ROT_TOL and t < TRANS_TOL


@pytest.mark.parametrize("n,path", [(5000, "grid"), (3000, "smem"), (3000, "grid-forced")])
def test_fitness(ctx, orc, synth, small_model, n, path, monkeypatch):
    """getFitnessScore through the spatial index (large targets) and by brute force from shared memory (targets of a few thousand
    points), with a max_range, NaN points on both sides and a target without any finite point"""
    if path == "grid-forced":
        monkeypatch.setenv("OPE_FITNESS_FORCE_GRID", "1")
    src, tgt, T = synth.icp_pair(n, seed=9, model=small_model)
    src, tgt = src.copy(), tgt.copy()
    src[3] = np.nan
    tgt[11, 2] = np.nan
    a, b = ctx.upload(src), ctx.upload(tgt)
    for max_range in (np.finfo(np.float64).max, 1e-5):
        g = ctx.fitness(a, b, T, max_range)
        o = orc.fitness(src, tgt, T, max_range)
        assert abs(g - o) < FIT_TOL * max(1.0, abs(o)) and abs(g - o) / o < 1e-6, (max_range, g, o)
    none = ctx.upload(np.full((8, 3), np.nan, np.float32))
    assert ctx.fitness(a, none, T) == orc.fitness(src, np.full((8, 3), np.nan, np.float32), T)


def test_correspondences_nearest_and_normal_shooting(ctx, orc, synth, cuda_lib, model):
    T = cuda_lib.T
    cl, _, _ = synth.make_frame(model, 5)
    tp = cl[orc.uniform_sample(cl, 0.008)]
    tn = orc.normals_knn(tp, 30)
    rng = np.random.default_rng(1)
    sp = synth.apply(synth.small_pose(rng, 5, 0.01), tp[::2]).astype(np.float32)
    sn = orc.normals_knn(sp, 30)
    cs, ct = ctx.upload(sp, sn), ctx.upload(tp, tn)
    for kw in (dict(max_correspondence_distance=0.01),
               dict(estimator=T.EST_NORMAL_SHOOTING, k_search=20, rejectors=[(T.REJ_SURFACE_NORMAL, 0.7), (T.REJ_SELF_OCCLUDED_NORMAL, 0.6)])):
        g = ctx.correspondences(cs, ct, cuda_lib.icp_params(**kw))
        o = orc.correspondences(sp, tp, orc.icp_params(**kw), sn, tn)
        assert len(g[0]) == len(o[0]) and len(g[0]) > 10
        assert (g[0] == o[0]).all() and (g[1] == o[1]).all() and (g[2] == o[2]).all()


def test_icp_point_to_point_c2_small(ctx, orc, synth, cuda_lib, small_model):
    """C2 at a size the oracle finishes in seconds: 50 iterations, max-corr-distance rejection."""
    T = cuda_lib.T
    src, tgt, Tgt = synth.icp_pair(8000, seed=2, model=small_model)
    kw = dict(max_iterations=50, max_correspondence_distance=0.05, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8)
    cs, ct = ctx.upload(src), ctx.upload(tgt)
    g, gc = ctx.icp(cs, ct, cuda_lib.icp_params(**kw), want_corr=True)
    o, oc = orc.icp(src, tgt, orc.icp_params(**kw), want_corr=True)
    _pose_close(synth, T, g.T, o.T)
    assert g.converged == o.converged and g.state == o.state and g.iterations == o.iterations
    assert g.n_correspondences == o.n_correspondences
    assert (gc[0] == oc[0]).all() and (gc[1] == oc[1]).all()
    r, t = synth.pose_error(T.mat4(g.T), Tgt)          # and it actually registers the pair (to ICP's noisy local optimum)
    r0, t0 = synth.pose_error(np.eye(4), Tgt)
    assert r < 0.5 * r0 and t < 0.2 * t0
    gf, of = ctx.fitness(cs, ct, T.mat4(g.T)), orc.fitness(src, tgt, T.mat4(o.T))
    assert abs(gf - of) < FIT_TOL


def test_icp_guess_and_aligned_output(ctx, orc, synth, cuda_lib, small_model):
    T = cuda_lib.T
    src, tgt, Tgt = synth.icp_pair(3000, seed=4, model=small_model)
    rng = np.random.default_rng(3)
    guess = synth.small_pose(rng, 2, 0.005).astype(np.float32)
    kw = dict(max_iterations=15, max_correspondence_distance=0.05)
    cs, ct = ctx.upload(src), ctx.upload(tgt)
    g, aligned = ctx.icp(cs, ct, cuda_lib.icp_params(**kw), guess=guess, want_aligned=True)
    o = orc.icp(src, tgt, orc.icp_params(**kw), guess=guess)
    _pose_close(synth, T, g.T, o.T)
    assert g.iterations == o.iterations and g.state == o.state
    out = aligned.download()
    ref = orc.transform(src, T.mat4(g.T))
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("target_path", ["small", "smem", "grid"])
def test_icp_with_normals_d_and_l_configuration(ctx, orc, synth, cuda_lib, model, target_path, monkeypatch):
    """The reference's fine stage (D&L/src/poseestimator.cpp:310-341): normal shooting k=20, surface-normal 0.7 +
    self-occluded 0.6 rejectors, SVD, 100 iterations, eps 1e-8, no distance gating — through all three device paths: the
    thread-per-query kernel for small clouds (default), the warp-per-query kernel with the target in shared memory, and the
    warp-per-query kernel over the spatial index."""
    if target_path == "grid":
        monkeypatch.setenv("OPE_ICP_FORCE_GRID", "1")
    if target_path == "small":
        monkeypatch.setenv("OPE_ICP_SMALL", "1")
    T = cuda_lib.T
    cl, _, pose = synth.make_frame(model, 5)
    rng = np.random.default_rng(8)
    start = pose @ synth.small_pose(rng, 6, 0.01)
    sp_full = synth.apply(start, model)
    sp = sp_full[orc.uniform_sample(sp_full, 0.008)]
    tp = cl[orc.uniform_sample(cl, 0.008)]
    sn, tn = orc.normals_knn(sp, 30), orc.normals_knn(tp, 30)
    kw = dict(max_iterations=100, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8, estimator=T.EST_NORMAL_SHOOTING,
              k_search=20, rejectors=[(T.REJ_SURFACE_NORMAL, 0.7), (T.REJ_SELF_OCCLUDED_NORMAL, 0.6)], with_normals=1)
    for variant in (T.ICP_VARIANT_MOD, T.ICP_VARIANT_MODCORR):
        cs, ct = ctx.upload(sp, sn), ctx.upload(tp, tn)
        g = ctx.icp(cs, ct, cuda_lib.icp_params(variant=variant, **kw))
        o = orc.icp(sp, tp, orc.icp_params(variant=variant, **kw), src_normals=sn, tgt_normals=tn)
        _pose_close(synth, T, g.T, o.T)
        assert g.converged == o.converged and g.state == o.state
        assert g.iterations == o.iterations and g.n_correspondences == o.n_correspondences
        gf, of = ctx.fitness(cs, ct, T.mat4(g.T)), orc.fitness(sp, tp, T.mat4(o.T))
        assert abs(gf - of) < FIT_TOL


def test_icp_reciprocal_and_fixed_correspondences(ctx, orc, synth, cuda_lib, small_model):
    """The VP-held extras of a7/a8: determineReciprocalCorrespondences (VP/impl/correspondence_estimation_mod.hpp:216-303) alone
    and inside the loop, and the reference's fixed correspondences (VP/icp_mod.h:267-281, VP/impl/icp_mod.hpp:209-225) without
    and with a rejector chain — same list (order included), same rewritten distances, same transform as the oracle, which
    tests/test_ref.py pins bit-exactly to the reference's own compiled sources."""
    T = cuda_lib.T
    src, tgt, _ = synth.icp_pair(4000, seed=6, model=small_model)
    tgt = tgt[:2500].copy()
    cs, ct = ctx.upload(src), ctx.upload(tgt)
    for d in (0.05, 0.003):
        g = ctx.correspondences(cs, ct, cuda_lib.icp_params(max_correspondence_distance=d, use_reciprocal=1))
        o = orc.correspondences_fixed(src, tgt, orc.icp_params(max_correspondence_distance=d, use_reciprocal=1))
        assert len(o[0]) > 100 and all(np.array_equal(a, b) for a, b in zip(g, o))
    kw = dict(max_iterations=25, max_correspondence_distance=0.05, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8,
              use_reciprocal=1)
    g, gc = ctx.icp(cs, ct, cuda_lib.icp_params(**kw), want_corr=True)
    o, oc = orc.icp(src, tgt, orc.icp_params(**kw), want_corr=True)
    _pose_close(synth, T, g.T, o.T)
    assert (g.converged, g.state, g.iterations, g.n_correspondences) == (o.converged, o.state, o.iterations, o.n_correspondences)
    assert all(np.array_equal(a, b) for a, b in zip(gc, oc))
    # fixed correspondences: six source points pinned to their third-nearest target point
    rng = np.random.default_rng(11)
    fq = rng.choice(len(src), 6, replace=False)
    fm = orc.knn(tgt, src[fq], 3)[0][:, 2]
    kw = dict(max_iterations=20, max_correspondence_distance=0.05, transformation_epsilon=1e-10, euclidean_fitness_epsilon=1e-12)
    g, gc, gd = ctx.icp(cs, ct, cuda_lib.icp_params(**kw), want_corr=True, fixed=(fq, fm))
    o, oc, od = orc.icp(src, tgt, orc.icp_params(**kw), want_corr=True, fixed=(fq, fm))
    _pose_close(synth, T, g.T, o.T)
    assert (g.converged, g.state, g.iterations, g.n_correspondences) == (o.converged, o.state, o.iterations, o.n_correspondences)
    assert np.array_equal(gc[0], oc[0]) and np.array_equal(gc[1], oc[1]) and np.allclose(gc[2], oc[2], rtol=1e-5, atol=0)
    assert np.allclose(gd, od, rtol=1e-5, atol=0)
    sn, tn = orc.normals_knn(src, 12), orc.normals_knn(tgt, 12)
    cs2, ct2 = ctx.upload(src, sn), ctx.upload(tgt, tn)
    kw = dict(max_iterations=12, max_correspondence_distance=0.05, transformation_epsilon=1e-10, euclidean_fitness_epsilon=1e-12,
              rejectors=[(T.REJ_SURFACE_NORMAL, 0.2), (T.REJ_SELF_OCCLUDED_NORMAL, -2.0)])
    g, gc, gd = ctx.icp(cs2, ct2, cuda_lib.icp_params(**kw), want_corr=True, fixed=(fq, fm))
    o, oc, od = orc.icp(src, tgt, orc.icp_params(**kw), src_normals=sn, tgt_normals=tn, want_corr=True, fixed=(fq, fm))
    _pose_close(synth, T, g.T, o.T)
    assert (g.converged, g.state, g.iterations, g.n_correspondences) == (o.converged, o.state, o.iterations, o.n_correspondences)
    assert np.array_equal(gc[0], oc[0]) and np.array_equal(gc[1], oc[1])


def test_get_icp2_two_stage_point_to_point(ctx, orc, synth, cuda_lib, small_model):
    """RegMeshPcd::getIcp2 (BM/src/regmeshpcd.cpp:46-59): two plain ICPs in sequence, the second on the first's aligned output
    with a tighter distance (0.05 / 80 iterations, then 0.005 / 500; transformation epsilon 1e-16, BM/src/regmeshpcd.cpp:16-41,
    225-226)."""
    T = cuda_lib.T
    src, tgt, _ = synth.icp_pair(5000, seed=12, model=small_model)
    cs, ct = ctx.upload(src), ctx.upload(tgt)
    o_src, g_src = src, cs
    for dist, iters in ((0.05, 80), (0.005, 500)):
        kw = dict(max_iterations=iters, max_correspondence_distance=dist, transformation_epsilon=1e-16)
        g, aligned = ctx.icp(g_src, ct, cuda_lib.icp_params(**kw), want_aligned=True)
        o = orc.icp(o_src, tgt, orc.icp_params(**kw))
        _pose_close(synth, T, g.T, o.T)
        assert (g.converged, g.state, g.iterations, g.n_correspondences) == (o.converged, o.state, o.iterations, o.n_correspondences)
        o_src = orc.transform(o_src, T.mat4(o.T))
        assert np.array_equal(aligned.download(), o_src)
        g_src = aligned
        assert abs(ctx.fitness(g_src, ct, np.eye(4)) - orc.fitness(o_src, tgt, np.eye(4))) < FIT_TOL


def test_icp_no_correspondences(ctx, orc, cuda_lib):
    rng = np.random.default_rng(0)
    src = rng.random((200, 3)).astype(np.float32)
    tgt = src + 100.0
    kw = dict(max_iterations=5, max_correspondence_distance=0.01)
    g = ctx.icp(ctx.upload(src), ctx.upload(tgt), cuda_lib.icp_params(**kw))
    o = orc.icp(src, tgt, orc.icp_params(**kw))
    assert g.converged == o.converged == 0 and g.state == o.state == cuda_lib.T.CONV_NO_CORRESPONDENCES
    assert np.array_equal(np.array(list(g.T)), np.array(list(o.T)))


@pytest.mark.parametrize("path", ["smem", "grid"])
def test_sacia_replayed_rng(ctx, orc, synth, cuda_lib, model, path, monkeypatch):
    """SAC-IA with the same hypothesis sequence replayed: every hypothesis error bit-exact, same winner, same transform —
    through both scoring kernels (target resident in shared memory / spatial index)."""
    monkeypatch.setenv("OPE_SACIA_PATH", path)
    T = cuda_lib.T
    cl, _, pose = synth.make_frame(model, 2)
    sp = model[orc.uniform_sample(model, 0.01)]
    tp = cl[orc.uniform_sample(cl, 0.01)]
    sn, tn = orc.normals_knn(sp, 30), orc.normals_knn(tp, 30)
    sf, tf = orc.fpfh(sp, sn, 0.03), orc.fpfh(tp, tn, 0.03)
    kw = dict(max_iterations=400, nr_samples=5, k_correspondences=5, min_sample_distance=0.01, max_correspondence_distance=0.05)
    orc.srand(1)
    samples, picks = orc.sacia_draw(sp, 400, 5, 5, 0.01)
    o, oe = orc.sacia(sp, sf, tp, tf, orc.sacia_params(**kw), orc.rng_table(samples, picks), want_errors=True)
    g, ge = ctx.sacia(ctx.upload(sp), sf, ctx.upload(tp), tf, cuda_lib.sacia_params(**kw), cuda_lib.rng_table(samples, picks),
                      want_errors=True)
    assert np.array_equal(ge, oe), np.abs(ge - oe).max()
    assert g.best_iteration == o.best_iteration and g.best_error == o.best_error
    assert np.array_equal(np.array(list(g.T)), np.array(list(o.T)))
    # the product draws the same table from libc rand() as the oracle does
    orc.srand(1)
    s2, p2 = cuda_lib.sacia_draw(sp, 400, 5, 5, 0.01)
    assert np.array_equal(s2, samples) and np.array_equal(p2, picks)
    # sharded pool: the union of shards reproduces the pool (multi-GPU contract)
    halves = []
    for b, e in ((0, 200), (200, 400)):
        r = ctx.sacia(ctx.upload(sp), sf, ctx.upload(tp), tf, cuda_lib.sacia_params(hypothesis_begin=b, hypothesis_end=e, **kw),
                      cuda_lib.rng_table(samples, picks))
        halves.append((np.float32(r.best_error), r.best_iteration))
    best = min(halves)
    assert best[1] == o.best_iteration


def test_pose_pipeline_matches_oracle(ctx, orc, synth, cuda_lib, model):
    """DetectAndLocalize frame (C1): full estimateFinalPose, two consecutive frames (coarse+fine, then tracking)."""
    T = cuda_lib.T
    cl, _, pose = synth.make_frame(model, 0)
    otr, gtr = orc.PoseEstimator(), cuda_lib.PoseTracker(ctx)
    osrc, gsrc = model.copy(), model.copy()
    for frame in range(2):
        orc.srand(1)
        o = otr.estimate_final(osrc, cl)
        orc.srand(1)
        g = gtr.estimate_final(gsrc, cl)
        assert g.ran_coarse == o.ran_coarse
        assert (g.n_src_coarse, g.n_tgt_coarse, g.n_src_fine, g.n_tgt_fine) == (o.n_src_coarse, o.n_tgt_coarse, o.n_src_fine, o.n_tgt_fine)
        assert g.sacia_best_iteration == o.sacia_best_iteration
        _pose_close(synth, T, g.coarse_pose, o.coarse_pose)
        _pose_close(synth, T, g.fine_pose, o.fine_pose)
        _pose_close(synth, T, g.final_pose, o.final_pose)
        assert g.icp_converged == o.icp_converged and g.icp_state == o.icp_state
        assert abs(g.fitness - o.fitness) < FIT_TOL
        assert g.icp_iterations == o.icp_iterations and abs(g.align_strength - o.align_strength) < 1e-12
        assert np.abs(gsrc - osrc).max() < 1e-4


def test_model_side_cache_in_the_tracker_is_invisible(ctx, synth, cuda_lib, model, monkeypatch):
    """SURVEY 8f-3 inside ope_pose_tracker: fresh trackers handed the same model reuse its 1 cm sample, normals and FPFH (content
    hash of the source cloud); a changed source (the tracking frame's aligned cloud) misses. Cached and uncached runs must return
    the same bits."""
    import ctypes
    libc = ctypes.CDLL(None)
    frames = [synth.make_frame(model, 40 + f)[0] for f in range(3)]

    def run():
        out = []
        for cl in frames:
            tr = cuda_lib.PoseTracker(ctx)
            src = model.copy()
            libc.srand(3)
            a = tr.estimate_final(src, cl)
            libc.srand(3)
            b = tr.estimate_final(src, cl)           # tracking frame: the source is now the aligned cloud
            out.append((np.array(list(a.final_pose)), np.array(list(b.final_pose)), a.fitness, b.fitness, a.icp_iterations, src.copy()))
            tr.close()
        return out

    monkeypatch.setenv("OPE_MODEL_CACHE", "0")
    plain = run()
    monkeypatch.setenv("OPE_MODEL_CACHE", "1")
    h0, m0 = ctx.model_cache_stats()
    cached = run()
    h1, m1 = ctx.model_cache_stats()
    assert h1 - h0 >= 2 and m1 - m0 >= 1            # frames 2 and 3 hit the entry frame 1 made
    for p, c in zip(plain, cached):
        assert np.array_equal(p[0], c[0]) and np.array_equal(p[1], c[1]) and p[2] == c[2] and p[3] == c[3] and p[4] == c[4]
        assert np.array_equal(p[5], c[5])


# ------------------------------------------------------------------------------ tcgen05 feature-distance GEMM ----
@pytest.mark.parametrize("nq,nt", [(300, 700), (1000, 1500), (2500, 20000)])
def test_feature_knn_gemm_path_is_bit_identical_to_the_exact_kernel(ctx, orc, nq, nt, monkeypatch):
    """K6 through the tensor cores: the GEMM only nominates candidates, the ranking is the exact float32 one, so indices AND
    distances must equal the exact kernel's / the oracle's bit for bit — on FPFH-like histograms (many near ties), with NaN
    rows, and on clustered data where the completeness proof must sometimes fall back to the exact kernel."""
    rng = np.random.default_rng(nq + nt)
    ft = rng.gamma(0.6, 8.0, size=(nt, 33)).astype(np.float32)
    ft = (100.0 * ft / ft.reshape(nt, 3, 11).sum(2).repeat(11, 1)).astype(np.float32)   # three sub-histograms summing to 100
    fq = ft[rng.integers(0, nt, nq)] + rng.normal(0, 0.5, size=(nq, 33)).astype(np.float32)
    fq[::17] = ft[rng.integers(0, nt, len(fq[::17]))]            # exact duplicates of targets: zero distances, index ties
    ft[5] = np.nan
    fq[3] = np.nan
    ft[100:140] = ft[100]                                        # 40 identical targets: a plateau wider than the candidate list
    monkeypatch.setenv("OPE_FEATURE_KNN", "exact")
    ei, ed = ctx.feature_knn(ft, fq, 5)
    monkeypatch.setenv("OPE_FEATURE_KNN", "gemm")
    g0, f0 = ctx.feature_knn_stats()
    gi, gd = ctx.feature_knn(ft, fq, 5)
    g1, f1 = ctx.feature_knn_stats()
    assert g1 - g0 == nq                                          # the tensor-core path really ran
    assert np.array_equal(gi, ei) and np.array_equal(gd, ed)
    oi, od = orc.feature_knn(ft[:2000], fq[:200], 5)
    if nt <= 2000:
        assert np.array_equal(gi[:200], oi) and np.array_equal(gd[:200], od)
    assert f1 - f0 < nq // 2, "the completeness proof should rarely need the exact kernel"
    assert ctx.last_kernel_ms(2) > 0


# ------------------------------------------------------------------------------------ depth image -> cloud (8f-1) ----
def _depth_mm(cloud):
    return np.rint(np.nan_to_num(cloud[..., 2], nan=0.0).astype(np.float64) * 1000.0).astype(np.uint16)


def test_depth_to_cloud_matches_the_reference_traversal(ctx, orc, synth, small_model):
    """DataGrabber::rgbd2Pcl on the device: same points, bit for bit, in the reference's column-outer order, including the
    row/column swap quirk, the Z == 0 / Z > 2 m drop and an image whose size is not a multiple of the 32x32 tile."""
    _, cloud, _ = synth.make_frame(small_model, 9)
    d = _depth_mm(cloud)
    d[10:20, 30:50] = 2500          # beyond the 2 m limit: dropped
    d[100:130, 600:] = 0            # invalid: dropped
    for img in (d, d[:471, :613].copy(), np.zeros((480, 640), np.uint16)):
        ref = orc.depth_to_cloud(img)
        got = ctx.depth_to_cloud(img)
        assert len(got) == len(ref)
        if len(ref):
            assert np.array_equal(got.download(), ref)
        got.free()


def test_depth_to_cloud_batch_is_the_single_frame_result_per_frame(ctx, orc, synth, small_model):
    import torch
    frames = []
    for f in (3, 4, 5):
        _, cloud, _ = synth.make_frame(small_model, f)
        frames.append(_depth_mm(cloud))
    depth = torch.from_numpy(np.stack(frames).astype(np.int16)).cuda().contiguous()   # raw uint16 bits
    B, R, Cc = depth.shape
    out = torch.empty((B * R * Cc, 4), dtype=torch.float32, device="cuda")
    col = torch.empty(B * Cc + 1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.depth_to_cloud_batch(depth.data_ptr(), B, R, Cc, out.data_ptr(), col.data_ptr())
    ctx.synchronize()
    col = col.cpu().numpy()
    out = out.cpu().numpy()
    for f in range(B):
        ref = orc.depth_to_cloud(frames[f])
        b, e = col[f * Cc], col[(f + 1) * Cc]
        assert e - b == len(ref) and np.array_equal(out[b:e, :3], ref)


# --------------------------------------------------------------- point-to-plane estimators (8a16 / 8f-4) ----
def _plane_pairs(synth, orc, small_model, seed):
    rng = np.random.default_rng(seed)
    tgt = (small_model[rng.permutation(len(small_model))[:6000]] + np.array([0.02, -0.01, 0.9], np.float32)).astype(np.float32)
    tn = orc.normals_knn(tgt, 12)
    moved = synth.apply(synth.small_pose(rng, 4, 0.008), tgt) + rng.normal(0, 3e-4, size=tgt.shape)
    return moved.astype(np.float32), tgt, tn


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_point_to_plane_estimators_match_the_oracle(ctx, orc, synth, cuda_lib, small_model, seed):
    """TransformationEstimationPointToPlaneLLS and ...PointToPlane (Levenberg-Marquardt on the 6-parameter warp) over explicit
    pairs: same transform, and for LM the same Eigen status / evaluation count / iteration count as the oracle, which factors the
    full m x 6 Jacobian by Householder QR while the device works from the fused 6 x 6 normal equations."""
    T = cuda_lib.T
    src, tgt, tn = _plane_pairs(synth, orc, small_model, seed)
    cs, ct = ctx.upload(src), ctx.upload(tgt, tn)
    rng = np.random.default_rng(seed)
    cases = [(None, None), (rng.integers(0, len(src), 3000).astype(np.int32),) * 2, (np.arange(3, dtype=np.int32),) * 2]
    for isrc, itgt in cases:
        for kind in (T.TE_POINT_TO_PLANE_LLS, T.TE_POINT_TO_PLANE):
            g, gi = ctx.point_to_plane(cs, ct, isrc, itgt, kind=kind, want_info=True)
            o, oi = orc.point_to_plane(src, tgt, tn, isrc, itgt, kind=kind, want_info=True)
            if isrc is not None and len(isrc) == 3:
                if kind == T.TE_POINT_TO_PLANE:      # fewer than 4 pairs: PCL_ERROR + identity
                    assert np.array_equal(g, np.eye(4)) and np.array_equal(o, np.eye(4))
                continue                              # LLS on 3 pairs is singular on both sides; nothing to compare
            r, t = synth.pose_error(g, o)
            assert r < ROT_TOL and t < TRANS_TOL, (kind, r, t)
            if kind == T.TE_POINT_TO_PLANE:
                assert gi == oi, (gi, oi)


def test_icp_with_normals_build_model_configuration(ctx, orc, synth, cuda_lib, small_model):
    """BuildModel's getIcpNormal (BM/src/regmeshpcd.cpp:104-208): normals k = 12, normal shooting k = 20, surface-normal
    rejector, TransformationEstimationPointToPlane (LM), eps 1e-8 — on neighbouring turntable views; and the default estimator
    of IterativeClosestPointWithNormals (LLS) with the same chain."""
    T = cuda_lib.T
    views = synth.turntable_views(small_model, n_views=36, first=2)
    (sp, _), (tp, _) = views
    sn, tn = orc.normals_knn(sp, 12), orc.normals_knn(tp, 12)
    for te, iters in ((T.TE_POINT_TO_PLANE, 40), (T.TE_POINT_TO_PLANE_LLS, 40)):
        kw = dict(max_iterations=iters, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8, estimator=T.EST_NORMAL_SHOOTING,
                  k_search=20, rejectors=[(T.REJ_SURFACE_NORMAL, 0.7)], with_normals=1, transformation=te)
        cs, ct = ctx.upload(sp, sn), ctx.upload(tp, tn)
        g, aligned = ctx.icp(cs, ct, cuda_lib.icp_params(**kw), want_aligned=True)
        o = orc.icp(sp, tp, orc.icp_params(**kw), src_normals=sn, tgt_normals=tn)
        _pose_close(synth, T, g.T, o.T)
        assert g.converged == o.converged and g.state == o.state and g.iterations == o.iterations
        assert g.n_correspondences == o.n_correspondences
        assert np.allclose(aligned.download(), synth.apply(T.mat4(g.T), sp), atol=2e-6)
        # the estimate is a real registration: the relative pose of the two views, to the accuracy a 10 degree step allows
        truth = views[1][1] @ np.linalg.inv(views[0][1])
        r, t = synth.pose_error(T.mat4(g.T), truth)
        assert r < np.deg2rad(3.0) and t < 0.01, (te, r, t)


# ------------------------------------------------------------------------------ batched frames (C5, 8e row 1 + 8f-3) ----
def test_pose_batch_equals_the_serial_first_frame_path(ctx, orc, synth, cuda_lib, model):
    """ope_pose_batch (worker threads with their own streams, the model side cached, the ICP grid capped so that frames overlap)
    returns for every frame what a fresh PoseEstimator returns for it on its own — and what the oracle returns."""
    import ctypes
    T = cuda_lib.T
    libc = ctypes.CDLL(None)
    frames = [synth.make_frame(model, 900 + f)[0] for f in range(6)]
    frames.append(frames[0][:60].copy())                 # too few points for the fine stage (< 100) and for SAC-IA's features
    frames.append(frames[1][::40].copy())                # a sparse cluster
    frames.append(np.zeros((0, 3), np.float32))          # an empty cluster: no target, identity poses
    # serial reference on the device: one fresh tracker per frame, SAC-IA drawing from libc rand() in frame order
    libc.srand(7)
    serial = []
    for cl in frames:
        tr = cuda_lib.PoseTracker(ctx)
        src = model.copy()
        serial.append(tr.estimate_final(src, cl))
        tr.close()
    libc.srand(7)
    for workers, clouds in ((3, False), (8, True)):
        libc.srand(7)
        inputs = [ctx.upload(f) if (clouds and len(f)) else f for f in frames]
        got, status = ctx.pose_batch(model, inputs, workers=workers)
        assert (status == 0).all()
        for f, (g, s) in enumerate(zip(got, serial)):
            r, t = synth.pose_error(T.mat4(g.final_pose), T.mat4(s.final_pose))
            assert r < 1e-6 and t < 1e-6, (workers, f, r, t)
            assert (g.icp_iterations, g.icp_state, g.icp_converged, g.sacia_best_iteration) == \
                   (s.icp_iterations, s.icp_state, s.icp_converged, s.sacia_best_iteration)
            assert g.n_src_coarse == s.n_src_coarse and g.n_tgt_fine == s.n_tgt_fine
            assert abs(g.fitness - s.fitness) < 1e-9
    # and against the oracle for the first two frames
    orc.srand(7)
    for f in range(2):
        o = orc.PoseEstimator().estimate_final(model.copy(), frames[f])
        r, t = synth.pose_error(T.mat4(serial[f].final_pose), np.array(o.final_pose, np.float64).reshape(4, 4).T)
        assert r < ROT_TOL and t < TRANS_TOL, (f, r, t)


def test_depth_to_cloud_every_raw_value_and_other_intrinsics(ctx, orc):
    """The device replaces the three IEEE divisions per pixel by a reciprocal multiply + two exact-residual corrections and the
    validity tests by an integer interval: every raw value 0..65535, at rows/columns that exercise all signs, with the Kinect
    intrinsics and with awkward ones (non-representable reciprocals, another scale and range), must still match bit for bit."""
    raw = np.arange(65536, dtype=np.uint16)
    img = np.resize(raw, (480, 640)).copy()          # 307 200 pixels: every raw value at least four times, at different (i, j)
    img[::7, ::5] = raw[(np.arange(img[::7, ::5].size) * 8191) % 65536].reshape(img[::7, ::5].shape)
    for kw in (dict(), dict(fx=570.3422, fy=571.9631, cx=314.5, cy=235.5, scale=1000.0, z_max=2.0),
               dict(fx=365.7, fy=366.1, cx=255.2, cy=211.9, scale=5000.0, z_max=4.5), dict(scale=3.0, z_max=1e4)):
        ref = orc.depth_to_cloud(img, **kw)
        got = ctx.depth_to_cloud(img, **kw)
        assert len(got) == len(ref)
        assert np.array_equal(got.download().view(np.uint32), ref.view(np.uint32))
        got.free()


# ------------------------------------------------------------------------------ the library's own communicator (8e row 2) ----
def test_sharded_sacia_through_the_library_communicator(ctx, orc, synth, cuda_lib, model):
    """ope_comm + ope_sacia_align_sharded. With one rank no NCCL is involved and the call must equal ope_sacia_align; with two or
    more GPUs on the box, two processes (one per GPU, the unique id passed through a file) must both return the single-GPU
    winner (tools/multigpu_check.py covers torchrun)."""
    import subprocess, sys, os, json
    cl, _, _ = synth.make_frame(model, 2)
    sp = model[orc.uniform_sample(model, 0.01)]
    tp = cl[orc.uniform_sample(cl, 0.01)]
    sn, tn = orc.normals_knn(sp, 30), orc.normals_knn(tp, 30)
    sf, tf = orc.fpfh(sp, sn, 0.03), orc.fpfh(tp, tn, 0.03)
    kw = dict(max_iterations=400, nr_samples=5, k_correspondences=5, min_sample_distance=0.01, max_correspondence_distance=0.05)
    orc.srand(1)
    table = cuda_lib.rng_table(*orc.sacia_draw(sp, 400, 5, 5, 0.01))
    cs, ct = ctx.upload(sp), ctx.upload(tp)
    one = ctx.sacia(cs, sf, ct, tf, cuda_lib.sacia_params(**kw), table)
    comm = cuda_lib.Comm(ctx, None, 1, 0)
    sh = comm.sacia(cs, sf, ct, tf, cuda_lib.sacia_params(**kw), table)
    comm.close()
    assert sh.best_iteration == one.best_iteration and sh.best_error == one.best_error
    assert np.array_equal(np.array(list(sh.T)), np.array(list(one.T)))
    import torch
    if torch.cuda.device_count() < 2:
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "multigpu_check.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MISMATCH" not in r.stdout, r.stdout[-2000:]


# ------------------------------------------------------------------------------ scene preparation (8f-2) ----
def test_pass_through_and_euclidean_clusters(ctx, orc, synth, model):
    """ProcessingPcd::getPassThrough (three closed intervals, NaN removed, order kept) and EuclideanClusterExtraction (tolerance
    0.05, sizes 300 .. 1e5) on a full 640x480 frame: kept indices and cluster labels bit-exact against the oracle."""
    _, cloud, _ = synth.make_frame(model, 21)
    pts = cloud.reshape(-1, 3)
    c = ctx.upload(pts)
    limits = (-0.5, 0.5, -0.5, 0.3, 0.5, 1.6)                    # ObjectSegmentationPlane::getFiltered, D&L/src/objectsegmentationplane.cpp:17-18
    g, gi = ctx.pass_through(c, limits, want_idx=True)
    keep = np.arange(len(pts), dtype=np.int32)
    for field, lo, hi in ((2, limits[4], limits[5]), (1, limits[2], limits[3]), (0, limits[0], limits[1])):   # z, y, x
        keep = keep[orc.pass_through(pts[keep], field, lo, hi)]
    assert np.array_equal(gi, keep) and np.array_equal(g.download(), pts[keep])
    # clustering: the object pixels + scattered table / wall points; also NaN points, a tiny component and an empty cloud
    rng = np.random.default_rng(4)
    sub = pts[keep][::3].copy()
    sub[10] = np.nan
    blob = (rng.normal(0, 0.004, (40, 3)) + np.array([0.3, 0.2, 0.7])).astype(np.float32)    # 40 points: below min_size
    data = np.concatenate([sub, blob]).astype(np.float32)
    for tol, mn in ((0.05, 300), (0.012, 50)):
        gl, gk = ctx.euclidean_clusters(ctx.upload(data), tol, mn, 100000)
        ol, ok = orc.euclidean_clusters(data, tol, mn, 100000)
        assert gk == ok and np.array_equal(gl, ol), (tol, gk, ok)
        assert gk >= 1 and (gl[-40:] == -1).all()
    gl, gk = ctx.euclidean_clusters(ctx.upload(np.zeros((0, 3), np.float32)))
    assert gk == 0 and len(gl) == 0


def test_plane_ransac_prism_and_object_clusters(ctx, orc, synth, model):
    """The rest of getSegmentedObjectsOnPlane (D&L/src/objectsegmentationplane.cpp:122-282) on full synthetic frames: the table
    plane by SACSegmentation with PCL's own sampling sequence, the polygonal prism over the padded hull rectangle, the second
    plane, the clusters — per-point labels, both plane equations and both RANSAC iteration counts equal to the oracle's bit for
    bit; and the largest non-wall cluster is the object (it contains the analytically known object pixels)."""
    limits = (-0.5, 0.5, -0.5, 0.3, 0.5, 1.6)
    for frame in (21, 22, 23):
        cl, cloud, _ = synth.make_frame(model, frame)
        c = ctx.upload(cloud.reshape(-1, 3))
        filt = ctx.pass_through(c, limits)
        sub = filt.download()
        found, coeff, inl, it = ctx.plane_ransac(filt)
        gl, gk, gp1, gp2, git = ctx.segment_objects_on_plane(filt)
        ol, ok, op1, op2, oit = orc.segment_objects_on_plane(sub)
        assert found and it == oit[0] and np.array_equal(coeff, op1)
        assert gk == ok and gk >= 1 and git == oit
        assert np.array_equal(gp1, op1) and np.array_equal(gp2, op2)
        assert np.array_equal(gl, ol), (frame, int((gl != ol).sum()))
        # the object: some cluster holds (nearly) all of the analytically segmented object points that survived the crop
        keys = set(map(bytes, np.ascontiguousarray(cl)))
        best = max(sum(bytes(p) in keys for p in np.ascontiguousarray(sub[gl == k])) for k in range(gk))
        assert best > 0.5 * len(cl)


def test_feature_knn_gemm_near_ties_at_large_norms(ctx, monkeypatch):
    """the completeness proof of the tcgen05 path must hold where its error budget is tightest — descriptors of large
    norm (|q||t| ~ 3e4, the FPFH maximum) whose best candidates are separated by less than the tensor-core rounding: the result must
    still equal the exact kernel's, by proof or by fallback."""
    rng = np.random.default_rng(9)
    base = np.zeros((1, 33), np.float32)
    base[0, [3, 14, 25]] = 100.0                                  # all mass in one bin per sub-histogram: |x| = 173
    ft = np.repeat(base, 3000, 0) + rng.normal(0, 0.02, (3000, 33)).astype(np.float32)   # 3000 targets within ~0.1 of each other
    fq = np.repeat(base, 400, 0) + rng.normal(0, 0.02, (400, 33)).astype(np.float32)
    ft[7] = np.inf                                                # a non-finite target row: never a neighbour
    monkeypatch.setenv("OPE_FEATURE_KNN", "exact")
    ei, ed = ctx.feature_knn(ft, fq, 5)
    monkeypatch.setenv("OPE_FEATURE_KNN", "gemm")
    gi, gd = ctx.feature_knn(ft, fq, 5)
    assert np.array_equal(gi, ei) and np.array_equal(gd, ed)
    assert not (gi == 7).any()
