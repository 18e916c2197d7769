#!/usr/bin/env python
"""Scene preparation on full 640x480 frames: device ms per frame, the oracle on one core, parity — bench.py's
`scene_preparation` key on its own.   python tools/bench_scene.py"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench, torch
import ope_pkg; ope_pkg.load()
from ope_b200 import cuda_lib, synth
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = cuda_lib.Context(0, stream.cuda_stream)
model = synth.bundled_model()
print(json.dumps(bench.scene_numbers(ctx, cuda_lib, synth, model, stream, torch)))
