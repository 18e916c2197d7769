"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libope_ref_{mod,modcorr}.so — the reference's own vendored registration
sources (VP) compiled against the mock PCL of oracle/refstub (see oracle/ref_harness.cpp). Used by tests/test_ref.py to pin
the oracle's restatement of the VP-held half of the path. Never imported by the product."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "object-pose-estimation_b200"))
import abi_types as T  # noqa: E402

_LIBS = {}
f32p = C.POINTER(C.c_float)


def available():
    return all(os.path.exists(os.path.join(_HERE, "_ref", "libope_ref_%s.so" % v)) for v in ("mod", "modcorr"))


def build():
    """needs /root/reference (only the build container has it); the built .so files travel to the GPU box"""
    subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=True)


def lib(variant="mod"):
    if variant not in _LIBS:
        # the oracle first (RTLD_GLOBAL): the harness resolves orc_point_to_plane from it
        C.CDLL(os.path.join(_HERE, "libope_oracle.so"), mode=C.RTLD_GLOBAL)
        L = C.CDLL(os.path.join(_HERE, "_ref", "libope_ref_%s.so" % variant), mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L.ref_variant.restype = C.c_char_p
        _LIBS[variant] = L
    return _LIBS[variant]


def _xyz(a):
    return np.ascontiguousarray(np.asarray(a, np.float32)[:, :3])


def _corr_buf(fixed):
    n = 0 if fixed is None else len(fixed[0])
    buf = (T.Correspondence * max(n, 1))()
    for i in range(n):
        buf[i] = T.Correspondence(int(fixed[0][i]), int(fixed[1][i]), float(fixed[2][i]) if len(fixed) > 2 else 0.0)
    return buf, n


def icp(src, tgt, prm, guess=None, src_normals=None, tgt_normals=None, fixed=None, variant="mod", want_corr=False):
    """returns dict(res=RegResult, corr=(q, m, d), aligned=(n,3), fitness, align_strength)"""
    s, t = _xyz(src), _xyz(tgt)
    sn = None if src_normals is None else np.ascontiguousarray(src_normals, np.float32)
    tn = None if tgt_normals is None else np.ascontiguousarray(tgt_normals, np.float32)
    res = T.RegResult()
    cbuf = (T.Correspondence * (len(s) + (0 if fixed is None else 2 * len(fixed[0])) + 1))()
    aligned = np.zeros((len(s), 3), np.float32)
    fit, strength = C.c_double(0), C.c_double(0)
    fbuf, nf = _corr_buf(fixed)
    g = None if guess is None else T.mat4_to_c(guess)
    rc = lib(variant).ref_icp(s.ctypes.data_as(f32p), C.c_size_t(len(s)), None if sn is None else sn.ctypes.data_as(f32p),
                              t.ctypes.data_as(f32p), C.c_size_t(len(t)), None if tn is None else tn.ctypes.data_as(f32p),
                              C.byref(prm), g, fbuf, C.c_size_t(nf), C.byref(res), cbuf, aligned.ctypes.data_as(f32p),
                              C.byref(fit), C.byref(strength))
    if rc != 0:
        raise RuntimeError("ref_icp rc=%d" % rc)
    a = np.ctypeslib.as_array(cbuf)[:res.n_correspondences]
    return {"res": res, "corr": (a["index_query"].copy(), a["index_match"].copy(), a["distance"].copy()), "aligned": aligned,
            "fitness": fit.value, "align_strength": strength.value}


def correspondences(src, tgt, max_distance, reciprocal=False, fixed=None, variant="mod"):
    s, t = _xyz(src), _xyz(tgt)
    fbuf, nf = _corr_buf(fixed)
    out = (T.Correspondence * (len(s) + nf + 1))()
    n = C.c_size_t(0)
    rc = lib(variant).ref_correspondences(s.ctypes.data_as(f32p), C.c_size_t(len(s)), t.ctypes.data_as(f32p), C.c_size_t(len(t)),
                                          C.c_double(max_distance), int(bool(reciprocal)), fbuf, C.c_size_t(nf), out, C.byref(n))
    if rc != 0:
        raise RuntimeError("ref_correspondences rc=%d" % rc)
    a = np.ctypeslib.as_array(out)[:n.value]
    return a["index_query"].copy(), a["index_match"].copy(), a["distance"].copy()
