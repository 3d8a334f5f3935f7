"""Host logic of the multi-GPU path on CPU: the C++ partitioner (host/partition.cpp through the flat shim: Hilbert order
of the overlap graph, pair ownership, halo images) against a Python restatement, and the host gather of the match lists
over a real world_size-2 (and 3) gloo process group."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from opencalibration_b200 import sharding, synthetic


def grid(rows, cols, k=10):
    pos = np.stack(np.meshgrid(np.arange(cols, dtype=float), np.arange(rows, dtype=float)), -1).reshape(-1, 2)
    pairs = []
    for i in range(len(pos)):
        d2 = ((pos - pos[i]) ** 2).sum(1)
        for j in np.argsort(d2, kind="stable")[:k]:
            if j != i:
                pairs.append((i, int(j)))
    return pos, pairs


# ---- restatement of the partition (the checker for the C++ code) ----
def ref_hilbert_index(order, x, y):
    # include/opencalibration/types/hilbert.hpp:8-27
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


def ref_partition(pos, pairs, world, order=1024):
    pos = np.asarray(pos, np.float64).reshape(-1, 2)
    n = len(pos)
    lo, hi = pos.min(0), pos.max(0)
    span = np.maximum(hi - lo, 1e-12)
    cells = np.minimum(((pos - lo) / span * order).astype(np.int64), order - 1)
    keys = np.array([ref_hilbert_index(order, int(cx), int(cy)) for cx, cy in cells], np.int64)
    curve = np.lexsort((np.arange(n), keys))
    load = np.zeros(n, np.int64)
    for a, _ in pairs:
        load[a] += 1
    cum = np.cumsum(load[curve])
    total = int(cum[-1])
    owner = np.zeros(n, np.int64)
    for k, img in enumerate(curve):
        owner[img] = min(world - 1, int(cum[k] - load[img]) * world // max(total, 1))
    return curve, owner


def test_hilbert_index_is_a_bijection_and_local(hostlib):
    n = 16
    idx = np.array([[sharding.hilbert_index(n, x, y) for y in range(n)] for x in range(n)])
    assert sorted(idx.ravel().tolist()) == list(range(n * n))
    where = {int(idx[x, y]): (x, y) for x in range(n) for y in range(n)}
    for d in range(n * n - 1):  # consecutive curve positions are grid neighbours
        (x0, y0), (x1, y1) = where[d], where[d + 1]
        assert abs(x0 - x1) + abs(y0 - y1) == 1
    # order-2 curve written out by hand from include/opencalibration/types/hilbert.hpp:8-27
    assert [sharding.hilbert_index(2, x, y) for x, y in [(0, 0), (0, 1), (1, 1), (1, 0)]] == [0, 1, 2, 3]
    for order, x, y in [(1024, 0, 0), (1024, 1023, 1023), (1024, 517, 3), (64, 63, 0), (8, 5, 6)]:
        assert sharding.hilbert_index(order, x, y) == ref_hilbert_index(order, x, y)


def test_hilbert_order_equals_the_restatement(hostlib):
    rng = np.random.default_rng(3)
    for pos in (grid(25, 40)[0], rng.uniform(-500, 2000, (777, 2)), np.zeros((5, 2)), rng.uniform(0, 1, (1, 2))):
        curve, _ = ref_partition(pos, [(0, 0)], 1)
        assert np.array_equal(sharding.hilbert_order(pos), curve)
    assert len(sharding.hilbert_order(np.zeros((0, 2)))) == 0


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_partition_covers_every_pair_once(hostlib, world):
    pos, pairs = grid(25, 40)
    shards = sharding.partition(pos, pairs, world)
    ids = np.concatenate([s.pair_ids for s in shards])
    assert sorted(ids.tolist()) == list(range(len(pairs)))
    owned = np.concatenate([s.owned_images for s in shards])
    assert sorted(owned.tolist()) == list(range(len(pos)))
    _, owner = ref_partition(pos, pairs, world)
    for s in shards:
        assert np.array_equal(s.owned_images, np.nonzero(owner == s.rank)[0])  # same cut as the restatement
        res = set(s.resident_images.tolist())
        assert all(a in res and b in res for a, b in s.pairs)
        assert np.all(np.diff(s.pair_ids) > 0)
        assert set(s.halo_images.tolist()) == {b for _, b in s.pairs} - set(s.owned_images.tolist())
    sizes = [len(s.pairs) for s in shards]
    assert max(sizes) - min(sizes) <= 0.05 * len(pairs) / world + 10  # balanced
    st = sharding.cut_statistics(shards, len(pairs))
    if world > 1:
        assert st["cut_pair_fraction"] < 0.25 and st["replication"] < 1.6  # overlap-graph locality
    else:
        assert st["cut_pair_fraction"] == 0 and st["replication"] == 1.0


def test_partition_edge_cases(hostlib):
    pos, pairs = grid(3, 3)
    shards = sharding.partition(pos, pairs, 16)  # more parts than images: some parts stay empty
    assert sum(len(s.pairs) for s in shards) == len(pairs)
    shards = sharding.partition(pos, [], 2)
    assert all(len(s.pairs) == 0 for s in shards)
    with pytest.raises(hostlib.OcbError):
        sharding.partition(pos, [(0, 99)], 2)


def _expected(a, b):
    n = (a * 31 + b * 17) % 13
    return np.stack([np.arange(n) + a * 1000 + b, np.arange(n) * 7 + b, (np.arange(n) * 5 + a) % 487], 1).astype(np.uint32)


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pos, pairs = grid(6, 7)
        shard = sharding.partition(pos, pairs, world)[rank]
        resident = set(shard.resident_images.tolist())
        assert all(a in resident and b in resident for a, b in shard.pairs)
        lists = [_expected(a, b) for a, b in shard.pairs]  # a stand-in matcher: variable-length, deterministic per pair
        shards = sharding.partition(pos, pairs, world)
        g = sharding.MatchGather(capacity_records=13 * len(pairs), max_pairs=len(pairs), dist=dist)
        rec, off, cnt = g.buffers()
        # written the way the runner's tail workers write them: in some order of completion, not in pair order
        cursor = 0
        for k in np.random.default_rng(rank).permutation(len(lists)):
            n = len(lists[k])
            rec[3 * cursor:3 * (cursor + n)] = lists[k].reshape(-1)
            off[k], cnt[k] = cursor, n
            cursor += n
        out = g.gather([s.pair_ids for s in shards], len(pairs))
        if rank == 0:
            ok = len(out) == len(pairs)
            for p, (a, b) in enumerate(pairs):
                i1, i2, d = out.pair(p)
                w = _expected(a, b)
                ok = ok and np.array_equal(i1, w[:, 0]) and np.array_equal(i2, w[:, 1]) and \
                    np.array_equal(d, w[:, 2] * (1.0 / 486))
            ret.put(("ok" if ok else "mismatch", len(out), out.total()))
        else:
            assert out is None
        dist.barrier()
        g.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_over_gloo(hostlib, world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(180) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    status, n, total = ret.get(timeout=5)
    pairs = grid(6, 7)[1]
    assert status == "ok" and n == len(pairs) and total == sum(len(_expected(a, b)) for a, b in pairs)


def test_gather_without_process_group():
    g = sharding.MatchGather(capacity_records=40, max_pairs=3)
    lists = [_expected(2, 1), _expected(0, 5), _expected(1, 1)]
    rec, off, cnt = g.buffers()
    cursor = 0
    for k, w in enumerate(lists):
        rec[3 * cursor:3 * (cursor + len(w))] = w.reshape(-1)
        off[k], cnt[k] = cursor, len(w)
        cursor += len(w)
    out = g.gather([[2, 0, 1]], 3)
    for p, w in zip([2, 0, 1], lists):
        assert np.array_equal(out.pair(p)[0], w[:, 0]) and np.array_equal(out.pair(p)[2], w[:, 2] * (1.0 / 486))
        assert np.array_equal(out.records(p), w)
    assert out.total() == sum(len(w) for w in lists)
    with pytest.raises(AssertionError):
        g.gather([[0, 0, 1]], 2)
    with pytest.raises(AssertionError):
        g.gather([[0]], 2)
    del out
    g.close()


def test_grid_survey_neighbours_share_descriptors():
    imgs, pos, pairs = synthetic.grid_survey(3, 3, 300, seed=2)
    # adjacent images share world points, so some rows are within noise distance of each other
    a, b = imgs[4], imgs[5]
    x = np.bitwise_xor(a[:, None, :], b[None, :64, :])
    d = np.unpackbits(x.view(np.uint8), axis=-1).sum(-1)
    assert d.min() < 120
