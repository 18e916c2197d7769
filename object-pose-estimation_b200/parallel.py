"""One process per GPU: how frame batches and SAC-IA hypothesis pools are partitioned, and the only collective on the path.

* Frame batches / BuildModel view pairs / independent ICP pairs are independent units: unit `u` belongs to rank
  `u mod world` (`shard_units`). No data-path collective; results are gathered by the host (optionally `all_gather`).
* The hypothesis pool of ONE alignment is split into contiguous ranges of the PRE-DRAWN libc rand() decision table
  (`shard_pool`), every rank evaluates its range on its GPU (ope_sacia_align with hypothesis_begin/end), and the ranks
  exchange exactly one 8-byte key and one 4x4:
      key = float_bits(error) << 32 | hypothesis_index      (error >= 0, so integer order == numeric order)
      all_reduce(MIN) on the key  ->  "first strictly lower error wins"  independent of the sharding,
      broadcast of the winner's 16 floats from the rank that owns it.
  72 bytes per alignment: NCCL over NVLink on GPUs (latency only), gloo in the CPU tests.

PyTorch is plumbing here (process group, tensors for the collective); nothing is computed with it.
"""
import numpy as np


def shard_units(n_units, rank, world):
    """indices of the independent units (frames, view pairs, ICP pairs) that belong to `rank`"""
    return list(range(rank, n_units, world))


def shard_pool(n_hypotheses, rank, world):
    """contiguous [begin, end) range of the hypothesis pool evaluated by `rank`"""
    per = (n_hypotheses + world - 1) // world
    b = min(n_hypotheses, rank * per)
    return b, min(n_hypotheses, b + per)


def pack_key(error, hypothesis):
    """(float32 error >= 0, hypothesis index) -> int64 key whose integer order is (error, index) lexicographic order"""
    e = np.float32(error)
    if not (e >= 0):  # NaN or negative: never wins
        return np.int64(np.iinfo(np.int64).max)
    bits = int(np.array([e], np.float32).view(np.uint32)[0])
    return np.int64((bits << 32) | (int(hypothesis) & 0xFFFFFFFF))


def unpack_key(key):
    key = int(key)
    bits = (key >> 32) & 0xFFFFFFFF
    return float(np.array([bits], np.uint32).view(np.float32)[0]), key & 0xFFFFFFFF


EMPTY_KEY = np.int64(np.iinfo(np.int64).max)


def reduce_best(local_error, local_hypothesis, local_T, group=None, device=None):
    """All ranks call this with their shard's best (error, hypothesis index, 4x4 as 16 floats; hypothesis < 0 = empty shard).
    Returns (error, hypothesis, T[16]) of the pool's winner on every rank. Without an initialised process group it
    returns the local triple (single-GPU run)."""
    import torch
    import torch.distributed as dist
    key = EMPTY_KEY if local_hypothesis is None or local_hypothesis < 0 else pack_key(local_error, local_hypothesis)
    T = np.asarray(local_T, np.float32).reshape(16)
    if not (dist.is_available() and dist.is_initialized()):
        e, h = unpack_key(key) if key != EMPTY_KEY else (float("inf"), -1)
        return e, (h if key != EMPTY_KEY else -1), T.copy()
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    k = torch.tensor([int(key)], dtype=torch.int64, device=dev)
    mine = int(key)
    dist.all_reduce(k, op=dist.ReduceOp.MIN, group=group)
    best = int(k.item())
    if best == int(EMPTY_KEY):
        return float("inf"), -1, np.eye(4, dtype=np.float32).T.reshape(16).copy()
    # the owner of the winning key is unique (hypothesis indices are disjoint across shards): find its rank, broadcast T
    rank = dist.get_rank(group)
    owner = torch.tensor([rank if mine == best else -1], dtype=torch.int64, device=dev)
    dist.all_reduce(owner, op=dist.ReduceOp.MAX, group=group)
    t = torch.from_numpy(T.copy()).to(dev)
    src = int(owner.item())
    dist.broadcast(t, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    e, h = unpack_key(best)
    return e, h, t.cpu().numpy()


def sharded_sacia(ctx, cuda_lib, src_cloud, fsrc, tgt_cloud, ftgt, prm_kwargs, table, group=None):
    """SampleConsensusInitialAlignment over a hypothesis pool sharded across the ranks of `group`.
    `table` is the pre-drawn decision table (cuda_lib.rng_table): every rank must hold the SAME table so that the set of
    hypotheses equals the serial run's. Returns (error, hypothesis, T 4x4 row-major numpy) on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    b, e = shard_pool(int(prm_kwargs["max_iterations"]), rank, world)
    if e > b:
        prm = cuda_lib.sacia_params(hypothesis_begin=b, hypothesis_end=e, **prm_kwargs)
        r = ctx.sacia(src_cloud, fsrc, tgt_cloud, ftgt, prm, table)
        local = (np.float32(r.best_error), r.best_iteration, np.array(list(r.T), np.float32))
    else:
        local = (np.float32(np.inf), -1, np.eye(4, dtype=np.float32).reshape(16))
    err, hyp, T = reduce_best(*local, group=group)
    return err, hyp, np.asarray(T, np.float32).reshape(4, 4).T
