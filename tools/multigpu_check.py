#!/usr/bin/env python
"""Multi-GPU check, run under torchrun with the NCCL backend (one process per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py

1. SAC-IA hypothesis pool of ONE alignment sharded over the ranks (pre-drawn libc rand() table, ope_sacia_align with
   hypothesis_begin/end, all_reduce(MIN) of the packed (error, index) key + broadcast of the winner's 4x4): every rank must
   end with exactly the single-GPU winner.
2. A batch of independent frames sharded frame f -> rank f mod N, no data-path collective; the poses are all-gathered only
   to print one line.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import torch.distributed as dist
    import ope_pkg
    ope_pkg.load()
    from ope_b200 import cuda_lib, parallel, synth
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = cuda_lib.Context(local, torch.cuda.current_stream().cuda_stream)
    model = synth.make_model(40000, seed=3)
    cl, _, _ = synth.make_frame(model, 1)
    cm, cc = ctx.upload(model), ctx.upload(cl)
    sp_c, tp_c = ctx.uniform_sample_cloud(cm, 0.01), ctx.uniform_sample_cloud(cc, 0.01)
    ctx.normals_knn(sp_c, 30); ctx.normals_knn(tp_c, 30)
    sf, tf = ctx.fpfh(sp_c, 0.03), ctx.fpfh(tp_c, 0.03)
    sp = sp_c.download()
    H = 400
    kw = dict(max_iterations=H, nr_samples=5, k_correspondences=5, min_sample_distance=0.01, max_correspondence_distance=0.05)
    import ctypes
    ctypes.CDLL(None).srand(1)      # every rank draws the same table from libc rand()
    samples, picks = cuda_lib.sacia_draw(sp, H, 5, 5, 0.01)
    table = cuda_lib.rng_table(samples, picks)
    full = ctx.sacia(sp_c, sf, tp_c, tf, cuda_lib.sacia_params(**kw), table)
    err, hyp, T = parallel.sharded_sacia(ctx, cuda_lib, sp_c, sf, tp_c, tf, kw, table)
    ok = (hyp == full.best_iteration and np.float32(err) == np.float32(full.best_error)
          and np.array_equal(T.T.reshape(16), np.array(list(full.T), np.float32)))
    print("rank %d/%d sharded SAC-IA: winner %d error %.6f %s" % (rank, world, hyp, err, "OK" if ok else "MISMATCH"), flush=True)
    # frame batch
    n_frames = 8
    mine = parallel.shard_units(n_frames, rank, world)
    poses = torch.zeros(n_frames, 16, device="cuda")
    for f in mine:
        fr, _, _ = synth.make_frame(model, 100 + f)
        tr = cuda_lib.PoseTracker(ctx)
        src = model.copy()
        ctypes.CDLL(None).srand(1)
        p = tr.estimate_final(src, fr)
        poses[f] = torch.tensor(list(p.final_pose), device="cuda")
        tr.close()
    if world > 1:
        dist.all_reduce(poses)   # disjoint rows: a sum is a gather
    if rank == 0:
        print("frame batch: %d frames over %d ranks, pose checksum %.6f" % (n_frames, world, float(poses.abs().sum().item())), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
