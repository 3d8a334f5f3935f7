"""Multi-GPU sharding of the pair list for the one-process-per-GPU harness.

The partition itself is product code in C++ (opencalibration_b200/host/partition.cpp: Hilbert order of the image
positions, same curve as the reference's include/opencalibration/types/hilbert.hpp:8-27, cut into `world` contiguous
runs balanced by sourced pairs; a pair belongs to the part that owns its SOURCE image; halo images are resident too);
this module only wraps it for bench.py / the tests, and implements the one cross-rank step of the path: the HOST GATHER
of the variable-length match lists back into the serial pair order, which is what LinkStage::finalize does with its
runners' results (reference src/pipeline/link_stage.cpp:119-131).

Gather. All ranks of the job run on one box (north_star: "the 8 GPUs of one box"), so the gather goes through POSIX
shared memory: rank 0 creates one segment, every rank packs its match lists (12-byte records: feature_index_1,
feature_index_2, integer Hamming distance) straight into its own region of it, and a barrier later rank 0 holds every
list without another copy; the serial order is restored as an index (pair -> offset, count), the way the reference
moves vector handles rather than their contents. torch.distributed carries the region sizes (all_gather) and the
barrier.
"""
from dataclasses import dataclass, field
from multiprocessing import shared_memory

import numpy as np


def hilbert_index(order, x, y):
    from . import host
    return host.hilbert_index(order, x, y)


def hilbert_order(positions):
    from . import host
    return host.hilbert_order(positions)


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
    """-> list of Shard, one per rank (host/partition.cpp: partition_pairs)."""
    from . import host
    owner, part, halo = host.partition_pairs(positions, pairs, world)
    shards = []
    for r in range(world):
        ids = np.nonzero(part == r)[0].astype(np.int64)
        shards.append(Shard(r, np.nonzero(owner == r)[0].astype(np.int64), np.nonzero(halo[r])[0].astype(np.int64), ids,
                            [pairs[i] for i in ids]))
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


RECORD_WORDS = 3  # uint32 words per match record: feature_index_1, feature_index_2, integer Hamming distance


class GatheredMatches:
    """What rank 0 holds after the gather: every pair's match list, addressable in serial pair order."""

    def __init__(self, records, offsets, counts):
        self.records, self.offsets, self.counts = records, offsets, counts  # records [total][3] uint32

    def __len__(self):
        return len(self.counts)

    def pair(self, p):
        """-> (feature_index_1, feature_index_2, distance) of pair p; distance = d * (1.0 / 486) like the reference."""
        r = self.records[int(self.offsets[p]):int(self.offsets[p]) + int(self.counts[p])]
        return r[:, 0].astype(np.uintp), r[:, 1].astype(np.uintp), r[:, 2] * (1.0 / 486)

    def total(self):
        return int(self.counts.sum())


class MatchGather:
    """Host gather of per-pair match lists to rank 0 through one shared-memory segment (see the module docstring).

    capacity_records = upper bound of the records any ONE rank contributes (e.g. its query rows); the segment holds
    world * capacity_records records of 12 bytes, of which only the pages actually written are ever touched."""

    def __init__(self, capacity_records, dist=None, name=None):
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        self.capacity = int(capacity_records)
        nbytes = max(1, self.world * self.capacity * RECORD_WORDS * 4)
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes, name=name)
            names = [self.shm.name]
        else:
            names = [None]
        if self.dist:
            self.dist.broadcast_object_list(names, src=0)
            if self.rank != 0:
                self.shm = shared_memory.SharedMemory(name=names[0])
                try:  # the creator unlinks it; an attaching process must not (Python < 3.13 registers it anyway)
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(self.shm._name, "shared_memory")
                except Exception:  # noqa: BLE001
                    pass
        self.words = np.ndarray((self.world, self.capacity * RECORD_WORDS), np.uint32, buffer=self.shm.buf)

    def region(self):
        """This rank's region of the segment (flat uint32): pack the match records straight into it."""
        return self.words[self.rank]

    def gather(self, pair_ids, counts, n_pairs):
        """Collective. pair_ids / counts: this rank's pairs (global ids, ascending) and their match counts, the records
        already packed into region() pair after pair. Returns GatheredMatches on rank 0, None elsewhere."""
        import torch
        pair_ids = np.asarray(pair_ids, np.int64)
        counts = np.asarray(counts, np.int64)
        assert len(pair_ids) == len(counts) and int(counts.sum()) <= self.capacity
        if self.dist:
            # sizes first (so that every rank can post a matching receive), then ids and counts padded to the longest
            dev = torch.device("cuda", torch.cuda.current_device()) if self.dist.get_backend() == "nccl" else "cpu"
            sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
            self.dist.all_gather(sizes, torch.tensor([len(pair_ids)], dtype=torch.int64, device=dev))
            longest = max(int(s.item()) for s in sizes)
            mine = torch.zeros(2, max(longest, 1), dtype=torch.int64)
            mine[0, :len(pair_ids)] = torch.from_numpy(pair_ids)
            mine[1, :len(counts)] = torch.from_numpy(counts)
            mine = mine.to(dev)
            every = [torch.zeros_like(mine) for _ in range(self.world)]
            self.dist.all_gather(every, mine)  # also orders every rank's writes to the segment before rank 0's reads
            self.dist.barrier()
            if self.rank != 0:
                return None
            parts = [(e[0, :int(s.item())].cpu().numpy(), e[1, :int(s.item())].cpu().numpy()) for e, s in zip(every, sizes)]
        else:
            parts = [(pair_ids, counts)]
        offsets = np.full(n_pairs, -1, np.int64)
        all_counts = np.zeros(n_pairs, np.int64)
        for r, (ids, cnt) in enumerate(parts):
            assert np.all(offsets[ids] == -1), "a pair was matched twice"
            start = np.concatenate([[0], np.cumsum(cnt)[:-1]]) if len(cnt) else np.zeros(0, np.int64)
            offsets[ids] = r * self.capacity + start
            all_counts[ids] = cnt
        assert np.all(offsets >= 0), f"{int((offsets < 0).sum())} pairs were never matched"
        return GatheredMatches(self.words.reshape(-1, RECORD_WORDS), offsets, all_counts)

    def close(self):
        self.words = None
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:  # noqa: BLE001
            pass
