"""K6 at C3 scale: feature-space 5-NN of 5 000 source descriptors against 307 200 target descriptors (FPFH-like), tensor-core
path vs exact float32 kernel; prints the featgemm_kernel time and its algorithmic TFLOP/s (2 * 33 * nq * nt)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ope_pkg
ope_pkg.load()
from ope_b200 import cuda_lib

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 307200
rng = np.random.default_rng(0)
ft = rng.gamma(0.6, 8.0, size=(nt, 33)).astype(np.float32)
ft = (100.0 * ft / ft.reshape(nt, 3, 11).sum(2).repeat(11, 1)).astype(np.float32)
fq = ft[rng.integers(0, nt, nq)] + rng.normal(0, 0.5, size=(nq, 33)).astype(np.float32)
ctx = cuda_lib.Context(0)
out = {}
for mode in ("gemm", "exact"):
    os.environ["OPE_FEATURE_KNN"] = mode
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter(); idx, d2 = ctx.feature_knn(ft, fq, 5); best = min(best, time.perf_counter() - t0)
    out[mode] = (idx, d2, best)
same = bool(np.array_equal(out["gemm"][0], out["exact"][0]) and np.array_equal(out["gemm"][1], out["exact"][1]))
kms = ctx.last_kernel_ms(2)
g, f = ctx.feature_knn_stats()
print(json.dumps({"nq": nq, "nt": nt, "identical": same, "featgemm_kernel_ms": kms,
                  "algorithmic_tflops": 2 * 33 * nq * nt / (kms * 1e-3) / 1e12,
                  "padded_tflops": 2 * 128 * nq * nt / (kms * 1e-3) / 1e12,
                  "call_ms_gemm_incl_h2d": out["gemm"][2] * 1e3, "call_ms_exact_incl_h2d": out["exact"][2] * 1e3,
                  "gemm_queries": g, "exact_fallbacks": f}))
