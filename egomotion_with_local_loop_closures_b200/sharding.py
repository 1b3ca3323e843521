"""Multi-GPU plumbing: shard independent frame-keyframe pairs across ranks and gather the fixed-size result records.

The path has no data-path collective: a track reads only its own pair.  Pairs are partitioned by
keyframe affinity -- every pair of a keyframe goes to the same rank, so its pyramid / selection lists are built
once -- with keyframes dealt to the least-loaded rank (by pair count).  The only exchange is one all-gather of
256-byte ellc_result records at the end of a batch (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_pairs_by_keyframe(kf_ids, world_size):
    """Return a list (one entry per rank) of index arrays into the pair list.

    Deterministic: keyframes are visited by decreasing pair count (ties by id) and given to the rank with the fewest
    pairs so far (ties by rank).  Within a rank, pairs keep their original order."""
    kf_ids = np.asarray(kf_ids)
    uniq, counts = np.unique(kf_ids, return_counts=True)
    order = np.lexsort((uniq, -counts))
    load = np.zeros(world_size, np.int64)
    owner = {}
    for i in order:
        r = int(np.lexsort((np.arange(world_size), load))[0])
        owner[int(uniq[i])] = r
        load[r] += counts[i]
    ranks = np.array([owner[int(k)] for k in kf_ids], np.int64) if len(kf_ids) else np.zeros(0, np.int64)
    return [np.nonzero(ranks == r)[0] for r in range(world_size)]


def shard_pairs(kf_ids, frame_ids, world_size):
    """Shard a pair list so that ranks share as little as possible.

    Pairs form a bipartite graph keyframes <-> frames.  Its connected components (a stretch of the sequence: frames and
    the keyframes they are matched against) are independent in DATA as well as in compute, so whole components are dealt
    to the least-loaded rank, largest first; a component bigger than 1.5x the fair share is split by keyframe affinity
    (its frames are then prepared on more than one rank).  Deterministic; returns one index array per rank, each in the
    original pair order."""
    kf_ids = np.asarray(kf_ids, np.int64)
    frame_ids = np.asarray(frame_ids, np.int64)
    n = len(kf_ids)
    if n == 0:
        return [np.zeros(0, np.int64) for _ in range(world_size)]
    ku, kinv = np.unique(kf_ids, return_inverse=True)
    fu, finv = np.unique(frame_ids, return_inverse=True)
    parent = np.arange(len(ku) + len(fu))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for a, b in zip(kinv, finv + len(ku)):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    comp = np.array([find(a) for a in kinv])
    cu, cinv, ccount = np.unique(comp, return_inverse=True, return_counts=True)
    fair = -(-n // world_size)
    load = np.zeros(world_size, np.int64)
    rank_of = np.full(n, -1, np.int64)
    for c in np.lexsort((cu, -ccount)):
        idx = np.nonzero(cinv == c)[0]
        if len(idx) > 1.5 * fair and world_size > 1:
            for r, sub in enumerate(shard_pairs_by_keyframe(kf_ids[idx], world_size)):
                tgt = int(np.lexsort((np.arange(world_size), load))[0]) if len(sub) else 0
                rank_of[idx[sub]] = tgt
                load[tgt] += len(sub)
        else:
            tgt = int(np.lexsort((np.arange(world_size), load))[0])
            rank_of[idx] = tgt
            load[tgt] += len(idx)
    return [np.nonzero(rank_of == r)[0] for r in range(world_size)]


def gather_results(local_records, local_indices, n_total, group=None, device=None, counts=None):
    """All-gather variable-length shards of fixed-size records and restore the global pair order.

    local_records: (n_local, record_bytes) uint8 torch tensor (CUDA for NCCL, CPU for gloo) or numpy structured array.
    local_indices: global pair indices of the local records.  counts: optional per-rank record counts (skips their exchange and
    its host synchronisation).  Returns a (n_total, record_bytes) uint8 tensor on every rank.
    """
    import torch
    import torch.distributed as dist

    if isinstance(local_records, np.ndarray):
        local_records = torch.from_numpy(local_records.view(np.uint8).reshape(len(local_records), -1))
    if device is not None:
        local_records = local_records.to(device)
    dev = local_records.device
    rec_bytes = local_records.shape[1] if local_records.ndim == 2 else 256
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    idx = torch.as_tensor(np.asarray(local_indices, np.int64), device=dev)
    out = torch.zeros((n_total, rec_bytes), dtype=torch.uint8, device=dev)
    if world == 1:
        out[idx] = local_records
        return out
    if counts is None:                                    # shard sizes: exchanged unless the caller already knows them (fixed sharding)
        n_local = torch.tensor([local_records.shape[0]], dtype=torch.int64, device=dev)
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(counts, n_local, group=group)
        counts = [int(c.item()) for c in counts]
    cmax = int(max(counts))
    pad_rec = torch.zeros((cmax, rec_bytes), dtype=torch.uint8, device=dev)
    pad_idx = torch.full((cmax,), -1, dtype=torch.int64, device=dev)
    pad_rec[: local_records.shape[0]] = local_records
    pad_idx[: idx.shape[0]] = idx
    all_rec = torch.empty((world * cmax, rec_bytes), dtype=torch.uint8, device=dev)
    all_idx = torch.empty((world * cmax,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_rec, pad_rec, group=group)
    dist.all_gather_into_tensor(all_idx, pad_idx, group=group)
    keep = all_idx >= 0
    out[all_idx[keep]] = all_rec[keep]
    return out
