"""B200-native registration hot path of gopi-erabati/Object-Pose-Estimation.

Python here is plumbing only: ctypes bindings of the C-ABI library (libope_cuda.so, built from csrc/),
a synthetic-scene generator for tests and benchmarks, and the frame/hypothesis sharding helpers that run
one process per GPU. The product is csrc/ (CUDA, sm_100a) behind include/ope_cuda.h and the PCL-style
C++ shim in include/ope_pcl/.

Import as `ope_b200` through the repo-root helper `ope_pkg.load()`.
"""
from . import abi_types  # noqa: F401

__all__ = ["abi_types"]
