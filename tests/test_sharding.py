"""Host logic of the multi-GPU path on CPU: Hilbert partition of the overlap graph, pair ownership, and the host
gather over a real world_size-2 (and 3) gloo process group."""
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


def test_hilbert_index_is_a_bijection_and_local():
    n = 16
    idx = np.array([[sharding.hilbert_index(n, x, y) for y in range(n)] for x in range(n)])
    assert sorted(idx.ravel().tolist()) == list(range(n * n))
    where = {int(idx[x, y]): (x, y) for x in range(n) for y in range(n)}
    for d in range(n * n - 1):  # consecutive curve positions are grid neighbours
        (x0, y0), (x1, y1) = where[d], where[d + 1]
        assert abs(x0 - x1) + abs(y0 - y1) == 1
    # order-2 curve written out by hand from include/opencalibration/types/hilbert.hpp:8-27
    assert [sharding.hilbert_index(2, x, y) for x, y in [(0, 0), (0, 1), (1, 1), (1, 0)]] == [0, 1, 2, 3]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_partition_covers_every_pair_once(world):
    pos, pairs = grid(25, 40)
    shards = sharding.partition(pos, pairs, world)
    ids = np.concatenate([s.pair_ids for s in shards])
    assert sorted(ids.tolist()) == list(range(len(pairs)))
    owned = np.concatenate([s.owned_images for s in shards])
    assert sorted(owned.tolist()) == list(range(len(pos)))
    for s in shards:
        res = set(s.resident_images.tolist())
        assert all(a in res and b in res for a, b in s.pairs)
        assert np.all(np.diff(s.pair_ids) > 0)
    sizes = [len(s.pairs) for s in shards]
    assert max(sizes) - min(sizes) <= 0.05 * len(pairs) / world + 10  # balanced
    st = sharding.cut_statistics(shards, len(pairs))
    if world > 1:
        assert st["cut_pair_fraction"] < 0.25 and st["replication"] < 1.6  # overlap-graph locality
    else:
        assert st["cut_pair_fraction"] == 0 and st["replication"] == 1.0


def _expected(a, b):
    return np.arange((a * 31 + b * 17) % 13, dtype=np.int64) + a * 1000 + b


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pos, pairs = grid(6, 7)
        uploaded = []

        def upload(ids):
            uploaded.extend(int(i) for i in ids)

        def match_batch(ps):  # a stand-in matcher: variable-length, deterministic per pair
            assert all(a in uploaded and b in uploaded for a, b in ps)
            return [_expected(a, b) for a, b in ps]

        out = sharding.run_sharded(pos, pairs, rank, world, upload, match_batch, dist)
        if rank == 0:
            want = [_expected(a, b) for a, b in pairs]
            ok = len(out) == len(want) and all(np.array_equal(x, y) for x, y in zip(out, want))
            ret.put(("ok" if ok else "mismatch", len(out)))
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_over_gloo(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(180) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    status, n = ret.get(timeout=5)
    assert status == "ok" and n == len(grid(6, 7)[1])


def test_gather_without_process_group():
    out = sharding.gather_results([2, 0, 1], ["c", "a", "b"], 3)
    assert out == ["a", "b", "c"]
    with pytest.raises(AssertionError):
        sharding.gather_results([0, 0], ["a", "b"], 2)
    with pytest.raises(AssertionError):
        sharding.gather_results([0], ["a"], 2)


def test_grid_survey_neighbours_share_descriptors():
    imgs, pos, pairs = synthetic.grid_survey(3, 3, 300, seed=2)
    # adjacent images share world points, so some rows are within noise distance of each other
    a, b = imgs[4], imgs[5]
    x = np.bitwise_xor(a[:, None, :], b[None, :64, :])
    d = np.unpackbits(x.view(np.uint8), axis=-1).sum(-1)
    assert d.min() < 120
