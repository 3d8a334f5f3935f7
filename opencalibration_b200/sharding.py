"""Multi-GPU sharding of the pair list: one process per GPU, no data-path collective.

Directed image pairs are independent units (reference src/pipeline/link_stage.cpp:75-112). Images are ordered along a
Hilbert curve over their positions and cut into `world` contiguous chunks (overlap-graph locality: most pairs have both
images in one chunk); a pair belongs to the rank that owns its SOURCE image; every rank uploads its own images plus the
"halo" images its pairs reference. The only cross-rank step is a host gather of the variable-length match lists back
into the serial order the reference restores in LinkStage::finalize (link_stage.cpp:119-131).
"""
from dataclasses import dataclass, field

import numpy as np


def hilbert_index(order, x, y):
    """Position of integer cell (x, y) along a Hilbert curve over an order x order grid (order = power of two).
    Same curve as the reference's xy2d helper (include/opencalibration/types/hilbert.hpp:8-27)."""
    d = 0
    s = order // 2
    while s > 0:
        rx = 1 if (x & s) else 0
        ry = 1 if (y & s) else 0
        d += s * s * ((3 * rx) ^ ry)
        if ry == 0:
            if rx == 1:
                x, y = s - 1 - x, s - 1 - y
            x, y = y, x
        s //= 2
    return d


def hilbert_order(positions, order=1024):
    """Permutation of image ids along the curve."""
    pos = np.asarray(positions, np.float64).reshape(-1, 2)
    if len(pos) == 0:
        return np.zeros(0, np.int64)
    lo, hi = pos.min(0), pos.max(0)
    span = np.maximum(hi - lo, 1e-12)
    cells = np.minimum(((pos - lo) / span * order).astype(np.int64), order - 1)
    keys = np.array([hilbert_index(order, int(cx), int(cy)) for cx, cy in cells], np.int64)
    return np.lexsort((np.arange(len(pos)), keys))


@dataclass
class Shard:
    rank: int
    owned_images: np.ndarray  # image ids whose outgoing pairs this rank matches
    halo_images: np.ndarray   # other images those pairs reference (uploaded too)
    pair_ids: np.ndarray      # indices into the global pair list, ascending (= serial order)
    pairs: list = field(default_factory=list)

    @property
    def resident_images(self):
        return np.concatenate([self.owned_images, self.halo_images])


def partition(positions, pairs, world):
    """-> list of Shard, one per rank. Chunks are balanced by the number of pairs each image sources."""
    n_img = len(positions)
    order = hilbert_order(positions)
    load = np.zeros(n_img, np.int64)
    for a, _ in pairs:
        load[a] += 1
    cum = np.cumsum(load[order])
    total = int(cum[-1]) if n_img else 0
    owner = np.zeros(n_img, np.int64)
    for pos_in_curve, img in enumerate(order):
        before = int(cum[pos_in_curve] - load[img])
        owner[img] = min(world - 1, before * world // max(total, 1))
    shards = []
    for r in range(world):
        ids = np.array([i for i, (a, _) in enumerate(pairs) if owner[a] == r], np.int64)
        owned = np.array(sorted(int(i) for i in np.nonzero(owner == r)[0]), np.int64)
        halo = np.array(sorted({int(pairs[i][1]) for i in ids} - set(owned.tolist())), np.int64)
        shards.append(Shard(r, owned, halo, ids, [pairs[i] for i in ids]))
    return shards


def cut_statistics(shards, n_pairs):
    """How good the locality is: fraction of pairs whose candidate image is a halo image; replication factor."""
    cut = 0
    for s in shards:
        halo = set(s.halo_images.tolist())
        cut += sum(1 for (_, b) in s.pairs if b in halo)
    resident = sum(len(s.owned_images) + len(s.halo_images) for s in shards)
    owned = sum(len(s.owned_images) for s in shards)
    return {"cut_pair_fraction": cut / max(n_pairs, 1), "replication": resident / max(owned, 1)}


def gather_results(local_pair_ids, local_results, n_pairs, dist=None, dst=0):
    """Host gather of per-pair results (any picklable objects, e.g. match arrays) to rank `dst`, restored to the
    global serial pair order. Without a process group (world 1) it just reorders."""
    payload = (np.asarray(local_pair_ids, np.int64), list(local_results))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        gathered = [payload]
        me = dst
    else:
        me = dist.get_rank()
        gathered = [None] * dist.get_world_size() if me == dst else None
        dist.gather_object(payload, gathered, dst=dst)
    if me != dst:
        return None
    out = [None] * n_pairs
    seen = 0
    for ids, res in gathered:
        for i, r in zip(ids.tolist(), res):
            assert out[i] is None, f"pair {i} matched twice"
            out[i] = r
            seen += 1
    assert seen == n_pairs, f"{n_pairs - seen} pairs were never matched"
    return out


def run_sharded(positions, pairs, rank, world, upload, match_batch, dist=None):
    """Drives one rank: upload(image_ids) makes the images resident; match_batch(pairs) -> list of per-pair results.
    Returns the full serial-order result list on rank 0, None elsewhere."""
    shard = partition(positions, pairs, world)[rank]
    upload(shard.resident_images)
    results = match_batch(shard.pairs) if len(shard.pairs) else []
    return gather_results(shard.pair_ids, results, len(pairs), dist)
