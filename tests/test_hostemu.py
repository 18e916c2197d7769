"""CPU tests of the product's DEVICE functions compiled by the host compiler (tests/hostemu): the per-query logic of
csrc/ope_grid.cuh and the scalar math of csrc/ope_device.cuh against the oracle. The kernels themselves run only on the GPU
(tests/test_gpu_parity.py); this keeps their logic honest on a box without one."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu")
f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return C.CDLL(os.path.join(HERE, "libope_hostemu.so"))


def _knn(emu, t, q, k, h, max_d2=3.0e38):
    t = np.ascontiguousarray(t[:, :3], np.float32); q = np.ascontiguousarray(q[:, :3], np.float32)
    idx = np.empty((len(q), k), np.int32); d2 = np.empty((len(q), k), np.float32)
    emu.emu_knn(t.ctypes.data_as(f32p), len(t), q.ctypes.data_as(f32p), len(q), k, C.c_float(h), C.c_float(max_d2),
                idx.ctypes.data_as(i32p), d2.ctypes.data_as(f32p))
    return idx, d2


@pytest.mark.parametrize("k,h", [(1, 0.02), (1, 0.3), (5, 0.05), (20, 0.08), (30, 0.01), (32, 0.7)])
def test_grid_knn_exact_volume(emu, orc, k, h):
    rng = np.random.default_rng(1)
    t = rng.random((4000, 3), dtype=np.float32)
    q = (rng.random((1500, 3), dtype=np.float32) * 1.6 - 0.3).astype(np.float32)   # inside and far outside the grid
    a, b = _knn(emu, t, q, k, h), orc.knn(t, q, k, brute=True)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


@pytest.mark.parametrize("k,h", [(1, 0.002), (20, 0.004), (12, 0.02)])
def test_grid_knn_exact_surface_and_ties(emu, orc, synth, small_model, k, h):
    src, tgt, _ = synth.icp_pair(4000, 0, small_model)
    a, b = _knn(emu, tgt, src, k, h), orc.knn(tgt, src, k, brute=True)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    tq = (np.round(small_model * 500) / 500).astype(np.float32)              # quantised: many exact ties and duplicates
    a, b = _knn(emu, tq, tq[:2000], k, h), orc.knn(tq, tq[:2000], k, brute=True)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_grid_nn1_bounded_search(emu, orc, synth, small_model):
    src, tgt, _ = synth.icp_pair(4000, 0, small_model)
    lim = np.float32(0.05) ** 2
    a, b = _knn(emu, tgt, src, 1, 0.002, max_d2=float(lim)), orc.knn(tgt, src, 1, brute=True)
    inside = b[1][:, 0] <= lim
    assert (a[0][inside] == b[0][inside]).all() and (a[1][inside] == b[1][inside]).all()
    assert ((a[0][~inside, 0] == -1) | (a[1][~inside, 0] > lim)).all()


@pytest.mark.parametrize("h", [0.0015, 0.004, 0.03])
def test_grid_nn1_seed_independent(emu, orc, synth, small_model, h):
    """the seeded search (previous ICP match as the initial bound) returns the same exact neighbour as the unseeded one"""
    src, tgt, _ = synth.icp_pair(4000, 2, small_model)
    rng = np.random.default_rng(3)
    ref_i, ref_d = orc.knn(tgt, src, 1, brute=True)
    t = np.ascontiguousarray(tgt, np.float32); q = np.ascontiguousarray(src, np.float32)
    for seeds in (ref_i[:, 0].copy(), rng.integers(0, len(tgt), len(src)).astype(np.int32),
                  np.where(rng.random(len(src)) < 0.5, -1, rng.integers(0, len(tgt), len(src))).astype(np.int32)):
        seeds = np.ascontiguousarray(seeds, np.int32)
        idx = np.empty(len(q), np.int32); d2 = np.empty(len(q), np.float32)
        emu.emu_nn1_seeded(t.ctypes.data_as(f32p), len(t), q.ctypes.data_as(f32p), len(q), seeds.ctypes.data_as(i32p),
                           C.c_float(h), C.c_float(3.0e38), idx.ctypes.data_as(i32p), d2.ctypes.data_as(f32p))
        assert (idx == ref_i[:, 0]).all() and (d2 == ref_d[:, 0]).all()


@pytest.mark.parametrize("h,gap", [(0.0015, 0.0008), (0.004, 0.002), (0.004, 0.0)])
def test_grid_nn1_certificate_radius(emu, orc, synth, small_model, h, gap):
    """the certificate of the ICP loop: after a search with a gap, no point other than the result lies within sqrt(r2), the
    result itself is the exact neighbour, and r2 reaches at least (d1 + gap)^2 unless a second point is closer than that"""
    src, tgt, _ = synth.icp_pair(3000, 5, small_model)
    t = np.ascontiguousarray(tgt, np.float32); q = np.ascontiguousarray(src, np.float32)
    lim = np.float32(0.05) ** 2
    idx = np.empty(len(q), np.int32); d2 = np.empty(len(q), np.float32); r2 = np.empty(len(q), np.float32)
    emu.emu_nn1_cert(t.ctypes.data_as(f32p), len(t), q.ctypes.data_as(f32p), len(q), None, C.c_float(h), C.c_float(float(lim)),
                     C.c_float(gap), idx.ctypes.data_as(i32p), d2.ctypes.data_as(f32p), r2.ctypes.data_as(f32p))
    ri, rd = orc.knn(tgt, src, 2, brute=True)
    inside = rd[:, 0] <= lim
    assert (idx[inside] == ri[inside, 0]).all() and (d2[inside] == rd[inside, 0]).all()
    # nothing else within the certified radius: the true second-nearest distance is >= r2
    assert (rd[inside, 1] >= r2[inside]).all()
    # and the radius is not vacuous: at least min(second distance, (min(d1, dmax) + gap)^2) (more when the seed was farther)
    want = np.minimum(rd[:, 1].astype(np.float64), (np.sqrt(np.minimum(rd[:, 0], lim).astype(np.float64)) + gap) ** 2)
    assert (r2[inside] >= want[inside] * (1 - 1e-4)).all()
    # queries with nothing inside the limit: everything is at least sqrt(r2) away and r2 covers the limit plus the gap
    out = ~inside
    # beyond the limit the returned index is only a candidate (the caller rejects it), but the radius must still hold for
    # every OTHER point: the nearest point that is not the returned one is at least sqrt(r2) away
    other = np.where(idx == ri[:, 0], rd[:, 1], rd[:, 0])
    assert (other[out] >= r2[out] * (1 - 1e-6)).all()


def test_grid_radius_counts(emu, orc, synth, small_model):
    src, tgt, _ = synth.icp_pair(3000, 0, small_model)
    cnt = np.empty(len(src), np.int32)
    t = np.ascontiguousarray(tgt, np.float32); q = np.ascontiguousarray(src, np.float32)
    emu.emu_radius_count(t.ctypes.data_as(f32p), len(t), q.ctypes.data_as(f32p), len(q), C.c_float(0.03), C.c_float(0.015),
                         cnt.ctypes.data_as(i32p))
    off, _, _ = orc.radius(tgt, src, 0.03)
    assert (np.diff(off) == cnt).all()


def test_device_normal_matches_oracle(emu, orc, synth, model):
    cl, _, _ = synth.make_frame(model, 11)
    pts = np.ascontiguousarray(cl[orc.uniform_sample(cl, 0.008)], np.float32)
    ref = orc.normals_knn(pts, 30)
    idx, _ = orc.knn(pts, pts, 30)
    out = np.empty((len(pts), 4), np.float32)
    vp = (C.c_float * 3)(0, 0, 0)
    for i in range(len(pts)):
        nn = np.ascontiguousarray(idx[i], np.int32)
        emu.emu_normal(pts.ctypes.data_as(f32p), nn.ctypes.data_as(i32p), 30, pts[i].ctypes.data_as(f32p), vp,
                       out[i].ctypes.data_as(f32p))
    assert np.abs(out - ref).max() < 2e-6
    assert (out == ref).mean() > 0.9                                  # the rest: last-bit differences of sin/cos/atan2


def test_device_pair_bins_match_oracle_spfh(emu, orc, small_model):
    pts = np.ascontiguousarray(small_model[orc.uniform_sample(small_model, 0.01)], np.float32)
    nr = orc.normals_knn(pts, 30)
    ref = orc.spfh(pts, nr, 0.03)
    off, idx, _ = orc.radius(pts, pts, 0.03)
    h = (C.c_int * 3)()
    bad = 0
    for p in range(0, len(pts), 7):
        nb = idx[off[p]:off[p + 1]]
        hist = np.zeros(33, np.int64)
        for q in nb:
            if q == p:
                continue
            n1 = np.ascontiguousarray(nr[p, :3]); n2 = np.ascontiguousarray(nr[q, :3])
            emu.emu_pair_bins(pts[p].ctypes.data_as(f32p), n1.ctypes.data_as(f32p), pts[q].ctypes.data_as(f32p),
                              n2.ctypes.data_as(f32p), h)
            hist[h[0]] += 1; hist[11 + h[1]] += 1; hist[22 + h[2]] += 1
        incr = np.float32(100.0) / np.float32(len(nb) - 1)
        val = np.zeros(33, np.float32)
        for b in range(33):
            v = np.float32(0)
            for _ in range(hist[b]):
                v = np.float32(v + incr)
            val[b] = v
        bad += int(not np.array_equal(val, ref[p]))
    assert bad == 0


def test_device_umeyama_variants(emu, orc, synth):
    rng = np.random.default_rng(7)
    for n in (5, 5, 5, 50, 3):
        s = np.ascontiguousarray(rng.normal(size=(n, 3)), np.float32)
        d = np.ascontiguousarray(rng.normal(size=(n, 3)), np.float32)
        out = (C.c_float * 16)()
        emu.emu_umeyama_moments(s.ctypes.data_as(f32p), d.ctypes.data_as(f32p), n, out)
        ref = orc.umeyama(s, d)
        assert np.array_equal(orc.T.mat4(out), ref)                      # double moments + double SVD on both sides
        out2 = (C.c_float * 16)()
        emu.emu_umeyama_small(s.ctypes.data_as(f32p), d.ctypes.data_as(f32p), n, out2)   # float variant: close, not equal
        assert np.abs(orc.T.mat4(out2) - ref).max() < 1e-4
    T = synth.random_pose(rng)
    s = np.ascontiguousarray(rng.normal(size=(2000, 3)) + [0, 0, 1], np.float32)
    d = np.ascontiguousarray(synth.apply(T, s))
    out = (C.c_float * 16)()
    emu.emu_umeyama_moments(s.ctypes.data_as(f32p), d.ctypes.data_as(f32p), len(s), out)
    r, t = synth.pose_error(orc.T.mat4(out), T)
    assert r < 1e-6 and t < 2e-6


@pytest.mark.parametrize("scale,f,c,dim", [(1000.0, 525.0, 319.5, 640), (1000.0, 525.0, 239.5, 480), (5000.0, 365.7, 255.2, 512),
                                           (1000.0, 570.3422, 314.5, 640), (3.0, 571.9631, 235.5, 480)])
def test_division_by_launch_constants_is_correctly_rounded(emu, scale, f, c, dim):
    """depth -> cloud replaces the IEEE divisions by a reciprocal multiply and two exact-residual FMA corrections
    (ope::div_by_const): for every raw depth value and every pixel index of an axis the result must equal the division's."""
    emu.emu_div_by_const_mismatches.restype = C.c_longlong
    assert emu.emu_div_by_const_mismatches(C.c_float(scale), C.c_float(f), C.c_float(c), dim) == 0
