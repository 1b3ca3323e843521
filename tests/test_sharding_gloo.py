"""CPU: N>1 host logic -- keyframe-affinity sharding and the result all-gather -- on world_size-2 gloo."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from egomotion_with_local_loop_closures_b200 import capi
from egomotion_with_local_loop_closures_b200.sharding import gather_results, shard_pairs, shard_pairs_by_keyframe


def test_shard_by_keyframe_properties():
    rng = np.random.default_rng(0)
    kf = rng.integers(0, 37, 1000)
    for world in (1, 2, 4, 8):
        shards = shard_pairs_by_keyframe(kf, world)
        allidx = np.concatenate(shards)
        assert sorted(allidx.tolist()) == list(range(1000))                  # a partition
        owners = {}
        for r, s in enumerate(shards):
            assert np.all(np.diff(s) > 0)                                    # original order kept
            for k in np.unique(kf[s]):
                assert owners.setdefault(int(k), r) == r                     # keyframe affinity
        sizes = np.array([len(s) for s in shards])
        assert sizes.max() - sizes.min() <= np.bincount(kf).max()            # balanced up to one keyframe
    assert [len(s) for s in shard_pairs_by_keyframe([], 2)] == [0, 0]


def test_shard_pairs_keeps_sequence_segments_together():
    # 4 independent segments (keyframes 8g..8g+7, frames 100g..100g+63, every frame vs 3 keyframes of its segment)
    rng = np.random.default_rng(1)
    kf, fr = [], []
    for g in range(4):
        for f in range(64):
            for k in rng.choice(8, 3, replace=False):
                kf.append(8 * g + k); fr.append(100 * g + f)
    perm = rng.permutation(len(kf))
    kf, fr = np.array(kf)[perm], np.array(fr)[perm]
    for world in (1, 2, 4):
        shards = shard_pairs(kf, fr, world)
        assert sorted(np.concatenate(shards).tolist()) == list(range(len(kf)))
        assert [len(s) for s in shards] == [len(kf) // world] * world
        for s in shards:
            assert np.all(np.diff(s) > 0)
            segs = set((kf[s] // 8).tolist())
            assert segs == set((fr[s] // 100).tolist()) and len(segs) == 4 // world      # whole segments, no frame shared
    # one giant component falls back to keyframe affinity
    shards = shard_pairs(np.arange(40) % 5, np.zeros(40, int), 2)
    assert sorted(np.concatenate(shards).tolist()) == list(range(40)) and min(len(s) for s in shards) >= 16


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kf = np.arange(n_total) % 5
    mine = shard_pairs_by_keyframe(kf, world)[rank]
    rec = np.zeros(len(mine), capi.RESULT_DTYPE)
    rec["pose"][:, 0] = mine                                                 # fabricate: pose[0] = global pair index
    rec["n_iters"][:, 0] = rank
    out = gather_results(rec, mine, n_total)
    got = out.numpy().view(capi.RESULT_DTYPE).reshape(-1)
    ok = np.array_equal(got["pose"][:, 0], np.arange(n_total, dtype=np.float32))
    owners = np.array([int(np.nonzero([i in s for s in shard_pairs_by_keyframe(kf, world)])[0][0]) for i in range(n_total)])
    ok = ok and np.array_equal(got["n_iters"][:, 0], owners)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_results_world2_gloo():
    world, n_total = 2, 23
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_total, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_gather_results_single_process():
    rec = np.zeros(4, capi.RESULT_DTYPE)
    rec["pose"][:, 1] = [3, 1, 0, 2]
    out = gather_results(rec, [3, 1, 0, 2], 4).numpy().view(capi.RESULT_DTYPE).reshape(-1)
    assert np.array_equal(out["pose"][:, 1], [0, 1, 2, 3])
