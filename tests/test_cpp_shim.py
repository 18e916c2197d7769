"""The drop-in boundary, exercised the way the reference applications exercise PCL: tests/cpp/pose_harness.cpp is C++ written
against <pcl/...> include paths and pcl:: class names (setInputSource, setInputTarget, setSearchMethod, align,
getFinalTransformation, hasConverged, getFitnessScore ...), compiled against include/ope_pcl_compat + include/ope_pcl and
linked to libope_cuda.so. Its results are compared with the CPU oracle running the same sequence on the same inputs."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
HARNESS = os.path.join(CPP, "pose_harness")

ROT_TOL, TRANS_TOL, FIT_TOL = 1e-4, 1e-5, 1e-5


@pytest.fixture(scope="module")
def harness(cuda_lib):
    cuda_lib.build()
    subprocess.run(["make", "-C", CPP, "-s"], check=True)
    return HARNESS


def _write(tmp_path, name, pts):
    p = os.path.join(str(tmp_path), name)
    np.ascontiguousarray(pts[:, :3], np.float32).tofile(p)
    return p


def _run(args, **kw):
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, **kw)
    return r.returncode, r.stdout, r.stderr


def _mat(v):
    return np.array(v, np.float64).reshape(4, 4).T  # printed column-major


def test_shim_headers_cover_the_reference_call_surface():
    """every PCL header the reference's hot-path translation units include for this path has a forwarder"""
    need = ["pcl/registration/icp.h", "pcl/registration/icp_mod.h", "pcl/registration/ia_ransac.h", "pcl/features/fpfh.h",
            "pcl/features/normal_3d.h", "pcl/filters/voxel_grid.h", "pcl/keypoints/uniform_sampling.h", "pcl/search/kdtree.h",
            "pcl/registration/correspondence_estimation_normal_shooting.h",
            "pcl/registration/correspondence_rejection_surface_normal.h",
            "pcl/registration/correspondence_rejection_self_occluded_normal.h",
            "pcl/registration/transformation_estimation_svd.h", "pcl/registration/transformation_estimation_point_to_plane.h",
            "pcl/registration/correspondence_estimation_mod.h", "pcl/common/transforms.h", "pcl/filters/filter.h"]
    for h in need:
        assert os.path.exists(os.path.join(ROOT, "include", "ope_pcl_compat", h)), h


def test_harness_without_a_gpu_fails_loudly_and_leaves_identity(harness, tmp_path, synth, small_model):
    """No CPU fallback behind the PCL-style classes: without a device align() reports it and leaves the transform at identity
    (the PCL_ERROR + return behaviour of VP/impl/registration_mod.hpp:73-77), it does not compute anything on the host."""
    src, tgt, _ = synth.icp_pair(500, seed=1, model=small_model)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    rc, out, err = _run([harness, "icp", _write(tmp_path, "s.bin", src), _write(tmp_path, "t.bin", tgt), "0.05", "5"], env=env)
    assert rc == 0, err
    res = json.loads(out)
    assert np.array_equal(_mat(res["T"]), np.eye(4)) and res["converged"] == 0
    assert "no usable CUDA device" in err


@pytest.mark.gpu
def test_plain_icp_through_the_pcl_api(harness, tmp_path, orc, synth, small_model):
    src, tgt, _ = synth.icp_pair(5000, seed=4, model=small_model)
    rc, out, err = _run([harness, "icp", _write(tmp_path, "s.bin", src), _write(tmp_path, "t.bin", tgt), "0.05", "40"])
    assert rc == 0, err
    g = json.loads(out)
    o = orc.icp(src, tgt, orc.icp_params(max_iterations=40, max_correspondence_distance=0.05, transformation_epsilon=1e-16))
    r, t = synth.pose_error(_mat(g["T"]), np.array(o.T, np.float64).reshape(4, 4).T)
    assert r < ROT_TOL and t < TRANS_TOL, (r, t)
    assert g["converged"] == o.converged and g["iterations"] == o.iterations and g["state"] == o.state
    of = orc.fitness(src, tgt, np.array(o.T, np.float32).reshape(4, 4).T)
    assert abs(g["fitness"] - of) < FIT_TOL
    # align(output): the input moved by the final transformation
    moved = synth.apply(_mat(g["T"]), src)
    assert g["n_out"] == len(src) and np.allclose(g["out_sum"], moved.astype(np.float64).sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("second", [1, 2])
def test_icp_realigns_a_cloud_modified_in_place(harness, tmp_path, orc, synth, small_model, second):
    """PCL keeps the caller's shared_ptr and re-reads the cloud on every align(): the shim must notice an in-place change
    (`*source = aligned`, no setInputSource) instead of aligning its stale device copy. Second stage = getIcp2's parameters
    (0.005 / 500, BM/src/regmeshpcd.cpp:46-59,225-226), once plain and once with reciprocal correspondences."""
    src, tgt, _ = synth.icp_pair(5000, seed=4, model=small_model)
    rc, out, err = _run([harness, "icp", _write(tmp_path, "s.bin", src), _write(tmp_path, "t.bin", tgt), "0.05", "80"],
                        env=dict(os.environ, OPE_ICP_SECOND_STAGE=str(second)))
    assert rc == 0, err
    dec, objs, rest = json.JSONDecoder(), [], out.strip()
    while rest:
        obj, end = dec.raw_decode(rest)
        objs.append(obj)
        rest = rest[end:].lstrip()
    first, g = objs
    o1 = orc.icp(src, tgt, orc.icp_params(max_iterations=80, max_correspondence_distance=0.05, transformation_epsilon=1e-16))
    moved = orc.transform(src, np.array(o1.T, np.float32).reshape(4, 4).T)
    o2 = orc.icp(moved, tgt, orc.icp_params(max_iterations=500, max_correspondence_distance=0.005, transformation_epsilon=1e-16,
                                            use_reciprocal=int(second == 2)))
    r, t = synth.pose_error(_mat(g["T"]), np.array(o2.T, np.float64).reshape(4, 4).T)
    assert r < ROT_TOL and t < TRANS_TOL, (r, t)
    assert (g["converged"], g["iterations"], g["state"]) == (o2.converged, o2.iterations, o2.state)
    assert not np.array_equal(_mat(g["T"]), _mat(first["T"]))


@pytest.mark.gpu
def test_detect_and_localize_sequence_through_the_pcl_api(harness, tmp_path, orc, synth, model):
    """two consecutive frames of estimateFinalPose (coarse + fine, then tracking) — libc rand() drives SAC-IA on both sides
    from its default seed, exactly as in the reference (PCL never seeds it)."""
    cl, _, _ = synth.make_frame(model, 0)
    rc, out, err = _run([harness, "pose", _write(tmp_path, "model.bin", model), _write(tmp_path, "f0.bin", cl),
                         _write(tmp_path, "f1.bin", cl)])
    assert rc == 0, err
    frames = json.loads(out)["frames"]
    otr = orc.PoseEstimator()
    osrc = model.copy()
    orc.srand(1)
    for f in range(2):
        o = otr.estimate_final(osrc, cl)
        g = frames[f]
        for key in ("coarse_pose", "fine_pose"):
            r, t = synth.pose_error(_mat(g[key]), np.array(getattr(o, key), np.float64).reshape(4, 4).T)
            assert r < ROT_TOL and t < TRANS_TOL, (f, key, r, t)
        r, t = synth.pose_error(_mat(g["final_pose"]), np.array(o.final_pose, np.float64).reshape(4, 4).T)
        assert r < ROT_TOL and t < TRANS_TOL, (f, r, t)
        assert g["icp_converged"] == o.icp_converged and g["icp_state"] == o.icp_state and g["icp_iterations"] == o.icp_iterations
        assert abs(g["fitness"] - o.fitness) < FIT_TOL
        assert abs(g["align_strength"] - o.align_strength) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("te", ["lm", "svd"])
def test_build_model_chain_through_the_pcl_api(harness, tmp_path, orc, synth, small_model, te):
    """BuildModel chain (C4, shortened; BM/src/regmeshpcd.cpp:210-271): getIcpNormal per pair with the estimator the reference
    sets (TransformationEstimationPointToPlane, Levenberg-Marquardt) and with SVD. Every pairwise alignment must equal the
    oracle running the same chain, converge to the ground-truth relative pose, and the merged cloud must grow by each view."""
    T = orc.T
    views = synth.turntable_views(small_model, n_views=36, first=4)
    files = [_write(tmp_path, "v%d.bin" % i, v) for i, (v, _) in enumerate(views)]
    rc, out, err = _run([harness, "chain", "0.7", "60"] + files, env=dict(os.environ, OPE_CHAIN_TE=te))
    assert rc == 0, err
    pairs = json.loads(out)["pairs"]
    assert len(pairs) == 3
    total = len(views[0][0])
    merged = views[0][0]
    kw = dict(max_iterations=60, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8, estimator=T.EST_NORMAL_SHOOTING,
              k_search=20, rejectors=[(T.REJ_SURFACE_NORMAL, 0.7)], with_normals=1,
              transformation=T.TE_POINT_TO_PLANE if te == "lm" else T.TE_SVD)
    for i, p in enumerate(pairs):
        target = views[i + 1][0]
        total += len(target)
        assert p["merged"] == total
        # source frame = view 0's merged cloud expressed in view i's camera; truth: T_{i+1} * inv(T_i)
        truth = views[i + 1][1] @ np.linalg.inv(views[i][1])
        r, t = synth.pose_error(_mat(p["T"]), truth)
        assert r < np.deg2rad(3.0) and t < 0.01, (i, r, t)
        assert p["fitness"] < 1e-4
        # the oracle on the same chain
        o = orc.icp(merged, target, orc.icp_params(**kw), src_normals=orc.normals_knn(merged, 12), tgt_normals=orc.normals_knn(target, 12))
        oT = T.mat4(o.T)
        r, t = synth.pose_error(_mat(p["T"]), oT)
        # the contract's bar for both estimators: the device's normals equal the oracle's (both evaluate the transcendentals of
        # eigen33 as correctly rounded floats), so even the Levenberg-Marquardt loop — discontinuous in its inputs at the 1e-5 m
        # level — takes the same path on both sides
        assert r < ROT_TOL and t < TRANS_TOL, (te, i, r, t)
        assert p["iterations"] == o.iterations and p["converged"] == o.converged
        # the next pair starts from the harness's own merged cloud, so that every pair is compared on identical inputs (an LM
        # chain amplifies the 1e-5 differences of one pair several times in the next)
        merged = np.concatenate([orc.transform(merged, _mat(p["T"]).astype(np.float32)), target])
