"""TEST INFRASTRUCTURE: ctypes binding of the CPU oracle (oracle/libope_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import ope_pkg  # noqa: E402

T = ope_pkg.load().abi_types
_LIB = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libope_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_radius.restype = C.c_int64
        L.orc_pose_create.restype = C.c_void_p
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _pts(a):
    a = _f32(a)
    assert a.ndim == 2 and a.shape[1] >= 3
    return a, a.ctypes.data_as(f32p), C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1])


def _chk(rc):
    if rc != 0:
        raise RuntimeError("oracle call failed: rc=%d" % rc)


def icp_params(**kw):
    p = T.IcpParams()
    lib().orc_icp_params_default(C.byref(p))
    _set(p, kw)
    return p


def sacia_params(**kw):
    p = T.SaciaParams()
    lib().orc_sacia_params_default(C.byref(p))
    _set(p, kw)
    return p


def pose_params(**kw):
    p = T.PoseParams()
    lib().orc_pose_params_default(C.byref(p))
    _set(p, kw)
    return p


def _set(p, kw):
    for k, v in kw.items():
        if k == "rejectors":
            p.n_rejectors = len(v)
            for i, (kind, thr) in enumerate(v):
                p.rejector_kind[i] = kind
                p.rejector_threshold[i] = thr
        else:
            setattr(p, k, v)


def knn(tgt, qry, k, brute=False):
    t, tp, nt, ts = _pts(tgt)
    q, qp, nq, qs = _pts(qry)
    idx = np.empty((q.shape[0], k), np.int32)
    d2 = np.empty((q.shape[0], k), np.float32)
    _chk(lib().orc_knn(tp, nt, ts, qp, nq, qs, k, int(brute), idx.ctypes.data_as(i32p), d2.ctypes.data_as(f32p)))
    return idx, d2


def radius(tgt, qry, r):
    t, tp, nt, ts = _pts(tgt)
    q, qp, nq, qs = _pts(qry)
    off = np.empty(q.shape[0] + 1, np.int64)
    total = lib().orc_radius(tp, nt, ts, qp, nq, qs, C.c_float(r), C.c_int64(0), off.ctypes.data_as(i64p), None, None)
    idx = np.empty(max(total, 1), np.int32)
    d2 = np.empty(max(total, 1), np.float32)
    lib().orc_radius(tp, nt, ts, qp, nq, qs, C.c_float(r), C.c_int64(total), off.ctypes.data_as(i64p),
                     idx.ctypes.data_as(i32p), d2.ctypes.data_as(f32p))
    return off, idx[:total], d2[:total]


def feature_knn(ftgt, fqry, k):
    ft, fq = _f32(ftgt), _f32(fqry)
    idx = np.empty((fq.shape[0], k), np.int32)
    d2 = np.empty((fq.shape[0], k), np.float32)
    _chk(lib().orc_feature_knn(ft.ctypes.data_as(f32p), C.c_size_t(ft.shape[0]), fq.ctypes.data_as(f32p),
                               C.c_size_t(fq.shape[0]), ft.shape[1], k, idx.ctypes.data_as(i32p),
                               d2.ctypes.data_as(f32p)))
    return idx, d2


def uniform_sample(pts, leaf):
    p, pp, n, s = _pts(pts)
    out = np.empty(max(p.shape[0], 1), np.int32)
    m = C.c_size_t(0)
    _chk(lib().orc_uniform_sample(pp, n, s, C.c_float(leaf), out.ctypes.data_as(i32p), C.byref(m)))
    return out[:m.value].copy()


def voxel_grid(pts, leaf, rgb=None):
    p, pp, n, s = _pts(pts)
    lx, ly, lz = (leaf, leaf, leaf) if np.isscalar(leaf) else leaf
    oxyz = np.empty((max(p.shape[0], 1), 3), np.float32)
    orgb = np.empty(max(p.shape[0], 1), np.float32)
    m = C.c_size_t(0)
    rgbp = None
    if rgb is not None:
        rgb = _f32(rgb)
        rgbp = rgb.ctypes.data_as(f32p)
    rc = lib().orc_voxel_grid(pp, n, s, rgbp, C.c_float(lx), C.c_float(ly), C.c_float(lz), oxyz.ctypes.data_as(f32p),
                              orgb.ctypes.data_as(f32p), C.byref(m))
    _chk(rc)
    return oxyz[:m.value].copy(), (orgb[:m.value].copy() if rgb is not None else None)


def normals_knn(pts, k, vp=(0, 0, 0)):
    p, pp, n, s = _pts(pts)
    out = np.empty((p.shape[0], 4), np.float32)
    v = (C.c_float * 3)(*vp)
    _chk(lib().orc_normals_knn(pp, n, s, k, v, out.ctypes.data_as(f32p)))
    return out


def fpfh(pts, normals, r):
    p, pp, n, s = _pts(pts)
    nr = _f32(normals)
    out = np.empty((p.shape[0], 33), np.float32)
    _chk(lib().orc_fpfh(pp, n, s, nr.ctypes.data_as(f32p), C.c_float(r), out.ctypes.data_as(f32p)))
    return out


def spfh(pts, normals, r):
    p, pp, n, s = _pts(pts)
    nr = _f32(normals)
    out = np.empty((p.shape[0], 33), np.float32)
    _chk(lib().orc_spfh(pp, n, s, nr.ctypes.data_as(f32p), C.c_float(r), out.ctypes.data_as(f32p)))
    return out


def umeyama(src, tgt, isrc=None, itgt=None):
    s, sp, ns, ss = _pts(src)
    t, tp, nt, ts = _pts(tgt)
    n = s.shape[0] if isrc is None else len(isrc)
    a = None if isrc is None else np.ascontiguousarray(isrc, np.int32)
    b = None if itgt is None else np.ascontiguousarray(itgt, np.int32)
    Tm = (C.c_float * 16)()
    _chk(lib().orc_umeyama(sp, ss, tp, ts, None if a is None else a.ctypes.data_as(i32p),
                           None if b is None else b.ctypes.data_as(i32p), C.c_size_t(n), Tm))
    return T.mat4(Tm)


def point_to_plane(src, tgt, tgt_normals, isrc=None, itgt=None, kind=None, want_info=False):
    """TransformationEstimationPointToPlane (kind = T.TE_POINT_TO_PLANE, Levenberg-Marquardt) or ...LLS"""
    s, sp, ns, ss = _pts(src)
    t, tp, nt, ts = _pts(tgt)
    nr = _f32(tgt_normals)
    assert nr.shape == (t.shape[0], 4)
    n = min(s.shape[0], t.shape[0]) if isrc is None else len(isrc)
    a = None if isrc is None else np.ascontiguousarray(isrc, np.int32)
    b = None if itgt is None else np.ascontiguousarray(itgt, np.int32)
    Tm = (C.c_float * 16)()
    info = (C.c_int32 * 3)()
    _chk(lib().orc_point_to_plane(sp, ns, ss, tp, nt, ts, nr.ctypes.data_as(f32p), None if a is None else a.ctypes.data_as(i32p),
                                  None if b is None else b.ctypes.data_as(i32p), C.c_size_t(n),
                                  int(T.TE_POINT_TO_PLANE if kind is None else kind), Tm, info))
    return (T.mat4(Tm), tuple(info)) if want_info else T.mat4(Tm)


def pass_through(pts, field, lo, hi):
    """SURVEY 8f-2 groundwork (oracle only): kept original indices of pcl::PassThrough on field 0/1/2"""
    p, pp, n, s = _pts(pts)
    out = np.empty(max(p.shape[0], 1), np.int32)
    lib().orc_pass_through.restype = C.c_int64
    m = lib().orc_pass_through(pp, n, s, int(field), C.c_float(lo), C.c_float(hi), out.ctypes.data_as(i32p))
    if m < 0:
        raise RuntimeError("orc_pass_through failed")
    return out[:m].copy()


def euclidean_clusters(pts, tolerance=0.05, min_size=300, max_size=100000):
    """SURVEY 8f-2 groundwork (oracle only): labels per point (0 = largest cluster, -1 = none) and the number of clusters"""
    p, pp, n, s = _pts(pts)
    labels = np.empty(max(p.shape[0], 1), np.int32)
    k = lib().orc_euclidean_clusters(pp, n, s, C.c_float(tolerance), int(min_size), int(max_size), labels.ctypes.data_as(i32p))
    if k < 0:
        raise RuntimeError("orc_euclidean_clusters failed")
    return labels[:p.shape[0]].copy(), k


def segment_objects_on_plane(pts, **kw):
    """ObjectSegmentationPlane::getSegmentedObjectsOnPlane on a filtered cloud: (labels, n_clusters or -1, plane1, plane2, iterations)"""
    p, pp, n, s = _pts(pts)
    prm = T.SegmentParams()
    lib().orc_segment_params_default(C.byref(prm))
    _set(prm, kw)
    labels = np.empty(max(p.shape[0], 1), np.int32)
    p1, p2 = (C.c_float * 4)(), (C.c_float * 4)()
    it = (C.c_int32 * 2)()
    k = lib().orc_segment_objects_on_plane(pp, n, s, C.byref(prm), labels.ctypes.data_as(i32p), p1, p2, it)
    return labels[:p.shape[0]].copy(), k, np.array(list(p1), np.float32), np.array(list(p2), np.float32), (it[0], it[1])


def lm_set_route(householder):
    lib().orc_lm_set_route(int(bool(householder)))


def depth_to_cloud(depth, fx=525.0, fy=525.0, cx=319.5, cy=239.5, scale=1000.0, z_max=2.0):
    d = np.ascontiguousarray(depth, np.uint16)
    out = np.empty((d.shape[0] * d.shape[1], 3), np.float32)
    lib().orc_depth_to_cloud.restype = C.c_int64
    n = lib().orc_depth_to_cloud(d.ctypes.data_as(C.POINTER(C.c_uint16)), d.shape[0], d.shape[1], C.c_float(fx), C.c_float(fy),
                                 C.c_float(cx), C.c_float(cy), C.c_float(scale), C.c_float(z_max), out.ctypes.data_as(f32p))
    return out[:n].copy()


def transform(pts, M, normals=None):
    p, pp, n, s = _pts(pts)
    out = np.empty((p.shape[0], 3), np.float32)
    on = None
    nrp = None
    onp = None
    if normals is not None:
        nr = _f32(normals)
        nrp = nr.ctypes.data_as(f32p)
        on = np.empty((p.shape[0], 4), np.float32)
        onp = on.ctypes.data_as(f32p)
    _chk(lib().orc_transform(pp, n, s, nrp, T.mat4_to_c(M), out.ctypes.data_as(f32p), onp))
    return (out, on) if normals is not None else out


def fitness(src, tgt, M, max_range=np.finfo(np.float64).max):
    s, sp, ns, ss = _pts(src)
    t, tp, nt, ts = _pts(tgt)
    out = C.c_double(0)
    _chk(lib().orc_fitness(sp, ns, ss, tp, nt, ts, T.mat4_to_c(M), C.c_double(max_range), C.byref(out)))
    return out.value


def _corr_to_np(buf, n):
    a = np.ctypeslib.as_array(buf)[:n]
    return a["index_query"].copy(), a["index_match"].copy(), a["distance"].copy()


def correspondences(src, tgt, prm, src_normals=None, tgt_normals=None):
    s, sp, ns, ss = _pts(src)
    t, tp, nt, ts = _pts(tgt)
    sn = None if src_normals is None else _f32(src_normals)
    tn = None if tgt_normals is None else _f32(tgt_normals)
    buf = (T.Correspondence * max(s.shape[0], 1))()
    m = C.c_size_t(0)
    _chk(lib().orc_correspondences(sp, ns, ss, None if sn is None else sn.ctypes.data_as(f32p), tp, nt, ts,
                                   None if tn is None else tn.ctypes.data_as(f32p), C.byref(prm), buf, C.byref(m)))
    return _corr_to_np(buf, m.value)


def fixed_buf(fixed):
    """(query indices, match indices[, distances]) -> ctypes array of pcl::Correspondence (in/out for the fixed-list calls)"""
    n = 0 if fixed is None else len(fixed[0])
    buf = (T.Correspondence * max(n, 1))()
    for i in range(n):
        buf[i] = T.Correspondence(int(fixed[0][i]), int(fixed[1][i]), float(fixed[2][i]) if len(fixed) > 2 else 0.0)
    return buf, n


def icp(src, tgt, prm, guess=None, src_normals=None, tgt_normals=None, want_corr=False, fixed=None):
    """fixed: (query indices, match indices) pinned by the caller (setFixedCorrespondences); with want_corr the rewritten
    fixed distances are returned as a third element"""
    s, sp, ns, ss = _pts(src)
    t, tp, nt, ts = _pts(tgt)
    sn = None if src_normals is None else _f32(src_normals)
    tn = None if tgt_normals is None else _f32(tgt_normals)
    res = T.RegResult()
    fb, nf = fixed_buf(fixed)
    buf = (T.Correspondence * max(s.shape[0] + 2 * nf, 1))() if want_corr else None
    g = None if guess is None else T.mat4_to_c(guess)
    rc = lib().orc_icp_fixed(sp, ns, ss, None if sn is None else sn.ctypes.data_as(f32p), tp, nt, ts,
                             None if tn is None else tn.ctypes.data_as(f32p), C.byref(prm), g, fb if nf else None, C.c_size_t(nf),
                             C.byref(res), buf)
    _chk(rc)
    if want_corr:
        out = (res, _corr_to_np(buf, res.n_correspondences))
        return out + (_corr_to_np(fb, nf)[2],) if nf else out
    return res


def correspondences_fixed(src, tgt, prm, fixed=None):
    """nearest-neighbour estimator with a fixed list in front / the reciprocal variant (prm.use_reciprocal), no rejectors"""
    s, sp, ns, ss = _pts(src)
    t, tp, nt, ts = _pts(tgt)
    fb, nf = fixed_buf(fixed)
    buf = (T.Correspondence * max(s.shape[0] + nf, 1))()
    m = C.c_size_t(0)
    _chk(lib().orc_correspondences_fixed(sp, ns, ss, tp, nt, ts, C.byref(prm), fb if nf else None, C.c_size_t(nf), buf, C.byref(m)))
    return _corr_to_np(buf, m.value)


def set_eigen_model(redux_level=3, cast_vectorized=False):
    """which Eigen build the oracle mimics for 4-float reductions (default: SSE3+ hadd order, Eigen 3.2 casts)"""
    lib().orc_set_eigen_model(int(redux_level), int(bool(cast_vectorized)))


def srand(seed):
    lib().orc_srand(C.c_uint(seed))


def sacia_draw(src, iterations, nr_samples, k_corr, min_sample_distance):
    s, sp, ns, ss = _pts(src)
    samples = np.empty((iterations, nr_samples), np.int32)
    picks = np.empty((iterations, nr_samples), np.int32)
    msd = C.c_float(min_sample_distance)
    _chk(lib().orc_sacia_draw(sp, ns, ss, iterations, nr_samples, k_corr, C.byref(msd), samples.ctypes.data_as(i32p),
                              picks.ctypes.data_as(i32p)))
    return samples, picks


def rng_table(samples, picks):
    samples = np.ascontiguousarray(samples, np.int32)
    picks = np.ascontiguousarray(picks, np.int32)
    t = T.RngTable(samples.shape[0], samples.shape[1], samples.ctypes.data_as(i32p), picks.ctypes.data_as(i32p))
    t._keep = (samples, picks)
    return t


def sacia(src, fsrc, tgt, ftgt, prm, table=None, want_errors=False):
    s, sp, ns, ss = _pts(src)
    t, tp, nt, ts = _pts(tgt)
    fs, ft = _f32(fsrc), _f32(ftgt)
    res = T.RegResult()
    errs = np.full(prm.max_iterations, np.nan, np.float32) if want_errors else None
    _chk(lib().orc_sacia(sp, ns, ss, fs.ctypes.data_as(f32p), tp, nt, ts, ft.ctypes.data_as(f32p), C.byref(prm),
                         None if table is None else C.byref(table), C.byref(res),
                         None if errs is None else errs.ctypes.data_as(f32p)))
    return (res, errs) if want_errors else res


class PoseEstimator:
    """D&L/src/poseestimator.cpp PoseEstimator, CPU oracle."""

    def __init__(self, prm=None):
        self._h = C.c_void_p(lib().orc_pose_create(None if prm is None else C.byref(prm)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_pose_destroy(self._h)
            self._h = None

    def estimate_final(self, source, target, table=None):
        """source: (N,3) float32 array, modified IN PLACE (-> alignedSource); target (M,>=3)."""
        assert source.dtype == np.float32 and source.flags.c_contiguous and source.shape[1] == 3
        t, tp, nt, ts = _pts(target) if len(target) else (None, None, C.c_size_t(0), C.c_size_t(3))
        res = T.PoseResult()
        _chk(lib().orc_pose_estimate_final(self._h, source.ctypes.data_as(f32p), C.c_size_t(source.shape[0]), tp, nt,
                                           ts, None if table is None else C.byref(table), C.byref(res)))
        return res

    def stage_seconds(self):
        out = (C.c_double * 8)()
        lib().orc_pose_stage_seconds(self._h, out)
        return list(out)
