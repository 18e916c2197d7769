"""Turns the model cloud the reference ships into the committed fixture tests/golden/drill_model.npz.

Run in the build container (the only place /root/reference exists):  python tests/golden/make_model_fixture.py
D&L/3DModel/drillNewModelOrigin.pcd: 157 825 points, fields x y z rgb (binary); it is the cloud BASELINE.json configs[0]
names ("the bundled 3DModel object") and what the ROS node hands PoseEstimator as p_sourceCloud. The fixture stores the
float32 coordinates and the packed colours bit for bit; nothing under -m gpu, smoke() or bench.py reads /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import pcd_io  # noqa: E402

SRC = "/root/reference/DetectAndLocalize/3DModel/drillNewModelOrigin.pcd"

if __name__ == "__main__":
    c = pcd_io.read_pcd(SRC)
    xyz = np.stack([c["x"], c["y"], c["z"]], 1).astype(np.float32)
    rgb = c["rgb"].view(np.uint32)
    assert len(xyz) == 157825 and np.isfinite(xyz).all()
    out = os.path.join(ROOT, "tests", "golden", "drill_model.npz")
    np.savez_compressed(out, xyz=xyz, rgb=rgb, source=np.array(os.path.relpath(SRC, "/root/reference")))
    print(out, os.path.getsize(out), xyz.min(0), xyz.max(0))
