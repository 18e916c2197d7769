"""Host-side logic of the multi-GPU path, on CPU with the gloo backend and world_size 2: the sharded SAC-IA hypothesis
pool reduces to the SAME winner (error bits, hypothesis index, transform) as the unsharded pool. The per-shard results come
from the CPU oracle here (the GPU evaluates shards through ope_sacia_align's hypothesis_begin/end; its equality with the
oracle per hypothesis is what tests/test_gpu_parity.py::test_sacia_replayed_rng checks)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, errors, transforms, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    import ope_pkg
    ope_pkg.load()
    from ope_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = parallel.shard_pool(len(errors), rank, world)
    # shard-local "first strictly lower error wins"
    best = -1
    for h in range(b, e):
        if best < 0 or errors[h] < errors[best]:
            best = h
    local = (errors[best], best, transforms[best]) if best >= 0 else (np.float32(np.inf), -1, np.eye(4, dtype=np.float32).reshape(16))
    err, hyp, T = parallel.reduce_best(*local)
    out.put((rank, float(err), int(hyp), np.asarray(T).tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pool_reduces_to_the_serial_winner(orc, synth, small_model, world):
    import torch.multiprocessing as mp
    cl, _, _ = synth.make_frame(small_model, 3)
    sp = small_model[orc.uniform_sample(small_model, 0.01)]
    tp = cl[orc.uniform_sample(cl, 0.01)]
    sf, tf = orc.fpfh(sp, orc.normals_knn(sp, 30), 0.03), orc.fpfh(tp, orc.normals_knn(tp, 30), 0.03)
    H = 64
    kw = dict(max_iterations=H, nr_samples=5, k_correspondences=5, min_sample_distance=0.01, max_correspondence_distance=0.05)
    orc.srand(7)
    samples, picks = orc.sacia_draw(sp, H, 5, 5, 0.01)
    table = orc.rng_table(samples, picks)
    serial, errors = orc.sacia(sp, sf, tp, tf, orc.sacia_params(**kw), table, want_errors=True)
    errors = np.asarray(errors, np.float32)
    errors[10] = errors[serial.best_iteration]   # an exact tie later in the pool must not steal the win
    if serial.best_iteration > 10:
        errors[10] = np.float32(errors[serial.best_iteration] * 2)
    # per-hypothesis transforms: re-run each hypothesis alone (shard of one)
    transforms = []
    for h in range(H):
        r = orc.sacia(sp, sf, tp, tf, orc.sacia_params(hypothesis_begin=h, hypothesis_end=h + 1, **kw), table)
        transforms.append(np.array(list(r.T), np.float32))
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, errors, transforms, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_h = int(np.argmin(errors))  # argmin returns the first minimum == "first strictly lower wins"
    for rank, err, hyp, T in got:
        assert hyp == want_h, (rank, hyp, want_h)
        assert np.float32(err) == errors[want_h]
        assert np.array_equal(np.array(T, np.float32), transforms[want_h])


def test_key_order_is_error_then_index(ope):
    from ope_b200 import parallel
    rng = np.random.default_rng(0)
    e = np.abs(rng.normal(size=200)).astype(np.float32)
    e[50] = e[3]
    keys = [parallel.pack_key(e[i], i) for i in range(len(e))]
    order = np.argsort(np.array(keys, np.int64), kind="stable")
    want = sorted(range(len(e)), key=lambda i: (e[i], i))
    assert list(order) == want
    assert parallel.unpack_key(keys[7]) == (float(e[7]), 7)
    assert parallel.pack_key(np.nan, 1) == parallel.EMPTY_KEY
    assert parallel.shard_pool(400, 7, 8) == (350, 400) and parallel.shard_pool(5, 7, 8) == (5, 5)
    assert parallel.shard_units(10, 1, 4) == [1, 5, 9]
