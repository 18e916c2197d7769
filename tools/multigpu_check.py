#!/usr/bin/env python
"""Multi-GPU check, run under torchrun with the NCCL backend (one process per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py

1. SAC-IA hypothesis pool of ONE alignment sharded over the ranks, twice: through the LIBRARY's own communicator
   (ope_comm_create + ope_sacia_align_sharded: ncclAllReduce(MIN) of the packed (error, index) key + ncclBroadcast of the
   winner's 4x4 on the context's stream, C ABI only; torch.distributed merely ships the 128-byte unique id) with a pool of 400
   and of 131 072 hypotheses, and through the Python helper (parallel.sharded_sacia). Every rank must end with exactly the
   single-GPU winner.
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
    print("rank %d/%d sharded SAC-IA (torch.distributed helper): winner %d error %.6f %s" % (rank, world, hyp, err, "OK" if ok else "MISMATCH"), flush=True)
    # the library's own communicator: NCCL inside libope_cuda.so
    import time
    uid = [cuda_lib.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    comm = cuda_lib.Comm(ctx, uid[0], world, rank)
    for pool in (H, 131072):
        kw2 = dict(kw, max_iterations=pool)
        ctypes.CDLL(None).srand(1)
        s2, p2 = cuda_lib.sacia_draw(sp, pool, 5, 5, 0.01)
        tb = cuda_lib.rng_table(s2, p2)
        ctx.synchronize()
        t0 = time.perf_counter()
        one = ctx.sacia(sp_c, sf, tp_c, tf, cuda_lib.sacia_params(**kw2), tb)
        t1 = time.perf_counter()
        if world > 1:
            dist.barrier()
        t2 = time.perf_counter()
        sh = comm.sacia(sp_c, sf, tp_c, tf, cuda_lib.sacia_params(**kw2), tb)
        t3 = time.perf_counter()
        same = (sh.best_iteration == one.best_iteration and np.float32(sh.best_error) == np.float32(one.best_error)
                and np.array_equal(np.array(list(sh.T), np.float32), np.array(list(one.T), np.float32)))
        ok = ok and same
        print("rank %d/%d ope_sacia_align_sharded pool %d: winner %d error %.6f %s | one GPU %.2f ms, sharded %.2f ms"
              % (rank, world, pool, sh.best_iteration, sh.best_error, "OK" if same else "MISMATCH", (t1 - t0) * 1e3, (t3 - t2) * 1e3), flush=True)
    comm.close()
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
