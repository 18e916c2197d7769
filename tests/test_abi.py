"""The C-ABI library loads without a GPU and exports every symbol include/ope_cuda.h declares (no compute calls here);
its entry points refuse to work without a device instead of falling back to the CPU."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "ope_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ope_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(cuda_lib):
    cuda_lib.build()
    lib = cuda_lib.lib()
    names = _declared()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(cuda_lib.EXPORTS) == names, set(names) ^ set(cuda_lib.EXPORTS)
    assert b"sm_100a" in lib.ope_version()


def test_no_cpu_fallback_without_a_device(cuda_lib):
    """on a box without a GPU context creation must fail with OPE_ERR_NO_DEVICE; on a GPU box it must succeed"""
    import torch
    lib = cuda_lib.lib()
    h = C.c_void_p()
    rc = lib.ope_ctx_create(0, None, C.byref(h))
    if torch.cuda.is_available():
        assert rc == 0
        lib.ope_ctx_destroy(h)
    else:
        assert rc == cuda_lib.T.OPE_ERR_NO_DEVICE and not h.value


def test_product_never_touches_the_oracle():
    """nothing under the product package, include/ or the C++ sources may reference oracle/ (it is test infrastructure)"""
    bad = []
    for base in ("object-pose-estimation_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    s = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"orc_py|ope_oracle|libope_oracle|oracle/", s):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
