"""ope_sacia_draw consumes the process's libc rand() stream in bulk (it borrows glibc's state through initstate/setstate instead
of taking rand()'s lock 4 000 times per alignment): the numbers, and the stream afterwards, must be exactly what calls of rand()
give — for every generator glibc can be switched to, and with the fast path turned off. Host-only: no device needed."""
import ctypes as C
import subprocess
import sys
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import ctypes as C, sys, os
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np
import ope_pkg; ope_pkg.load()
from ope_b200 import cuda_lib
import orc_py as orc
libc = C.CDLL(None)
libc.initstate.restype = C.c_void_p
libc.initstate.argtypes = [C.c_uint, C.c_void_p, C.c_size_t]
size = int(sys.argv[1])
keep = C.create_string_buffer(256)
if size:
    libc.initstate(1, keep, size)          # glibc: 8 -> LCG, 32 / 64 / 128 / 256 -> additive generators of 7 / 15 / 31 / 63 words
rng = np.random.default_rng(3)
sp = (rng.random((300, 3)) * 0.2).astype(np.float32)
out = []
for seed in (1, 12345):
    libc.srand(seed)
    head = [libc.rand() for _ in range(5)]                 # the stream is mid-way, not freshly seeded
    s_ref, p_ref = orc.sacia_draw(sp, 40, 5, 5, 0.01)      # the oracle: plain rand() calls
    tail_ref = [libc.rand() for _ in range(7)]
    libc.srand(seed)
    assert [libc.rand() for _ in range(5)] == head
    s, p = cuda_lib.sacia_draw(sp, 40, 5, 5, 0.01)
    s2, p2 = cuda_lib.sacia_draw(sp, 3, 5, 5, 0.01)        # a second session continues where the first stopped
    libc.srand(seed)
    [libc.rand() for _ in range(5)]
    orc.sacia_draw(sp, 40, 5, 5, 0.01)
    s2_ref, p2_ref = orc.sacia_draw(sp, 3, 5, 5, 0.01)
    after_ref = [libc.rand() for _ in range(7)]
    libc.srand(seed)
    [libc.rand() for _ in range(5)]
    cuda_lib.sacia_draw(sp, 40, 5, 5, 0.01)
    tail = [libc.rand() for _ in range(7)]                 # rand() itself resumes right after the session
    ok = (np.array_equal(s, s_ref) and np.array_equal(p, p_ref) and np.array_equal(s2, s2_ref) and np.array_equal(p2, p2_ref)
          and tail == tail_ref)
    out.append(ok)
    del after_ref
print("OK" if all(out) else "MISMATCH", out)
"""


@pytest.mark.parametrize("state_bytes", [0, 8, 32, 64, 128, 256])
@pytest.mark.parametrize("fast", ["1", "0"])
def test_bulk_draws_equal_rand_calls(state_bytes, fast):
    env = dict(os.environ, OPE_LIBC_RAND_FAST=fast)
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}, str(state_bytes)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().startswith("OK"), r.stdout + r.stderr[-1000:]
