"""Minimal PCD v0.7 reader/writer (ascii + binary, float32/uint32 fields) for tests and tools only.

The product never reads files: clouds cross the C ABI as host arrays. This exists so that the model clouds the
reference ships (D&L/3DModel/*.pcd, what PoseEstimator is handed at D&L/src/rosinterface.cpp:170-186) can be turned
into the committed fixture tests/golden/drill_model.npz (see tests/golden/make_model_fixture.py) and so that a cloud
can be dumped for inspection.
"""
import numpy as np

_NP = {("F", 4): np.float32, ("F", 8): np.float64, ("U", 1): np.uint8, ("U", 2): np.uint16, ("U", 4): np.uint32,
       ("I", 1): np.int8, ("I", 2): np.int16, ("I", 4): np.int32}


def read_pcd(path):
    """Returns a dict field -> array (n,) in file order, plus '_width'/'_height'."""
    with open(path, "rb") as f:
        hdr = {}
        while True:
            line = f.readline()
            if not line:
                raise ValueError("no DATA line in %s" % path)
            s = line.decode("ascii", "replace").strip()
            if not s or s.startswith("#"):
                continue
            key, _, rest = s.partition(" ")
            hdr[key] = rest.split()
            if key == "DATA":
                break
        fields = hdr["FIELDS"]
        sizes = [int(x) for x in hdr["SIZE"]]
        types = hdr["TYPE"]
        counts = [int(x) for x in hdr.get("COUNT", ["1"] * len(fields))]
        n = int(hdr["POINTS"][0])
        if any(c != 1 for c in counts):
            raise ValueError("COUNT != 1 not supported")
        dt = np.dtype([(fn, _NP[(t, s)]) for fn, t, s in zip(fields, types, sizes)])
        mode = hdr["DATA"][0]
        if mode == "binary":
            rec = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
        elif mode == "ascii":
            rec = np.loadtxt(f, dtype=dt, ndmin=1)
        else:
            raise ValueError("DATA %s not supported" % mode)
    out = {fn: np.ascontiguousarray(rec[fn]) for fn in fields}
    out["_width"], out["_height"] = int(hdr["WIDTH"][0]), int(hdr["HEIGHT"][0])
    return out


def write_pcd(path, xyz, rgb=None):
    """binary PCD with fields x y z [rgb] (rgb as the packed float PCL uses)."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    n = len(xyz)
    cols = [xyz]
    fields, sizes, types = "x y z", "4 4 4", "F F F"
    if rgb is not None:
        cols.append(np.ascontiguousarray(rgb).view(np.float32).reshape(n, 1))
        fields, sizes, types = fields + " rgb", sizes + " 4", types + " F"
    data = np.ascontiguousarray(np.concatenate(cols, 1), np.float32)
    with open(path, "wb") as f:
        f.write(("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\n"
                 "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n"
                 % (fields, sizes, types, " ".join(["1"] * data.shape[1]), n, n)).encode())
        f.write(data.tobytes())
