"""Multi-GPU sharding of the pair list for the one-process-per-GPU harness.

The partition itself is product code in C++ (opencalibration_b200/host/partition.cpp: Hilbert order of the image
positions, same curve as the reference's include/opencalibration/types/hilbert.hpp:8-27, cut into `world` contiguous
runs balanced by sourced pairs; a pair belongs to the part that owns its SOURCE image; halo images are resident too);
this module only wraps it for bench.py / the tests, and implements the one cross-rank step of the path: the HOST GATHER
of the variable-length match lists back into the serial pair order, which is what LinkStage::finalize does with its
runners' results (reference src/pipeline/link_stage.cpp:119-131).

Gather. All ranks of the job run on one box (north_star: "the 8 GPUs of one box"), so the gather goes through POSIX
shared memory: rank 0 creates one segment, every rank's tail workers write its match lists (12-byte records:
feature_index_1, feature_index_2, integer Hamming distance) and their index straight into its own region of it while
the rank is still matching, and a barrier later rank 0 holds every list without another copy; the serial order is
restored as an index (pair -> offset, count), the way the reference moves vector handles rather than their contents.
torch.distributed carries the segment's name and the barrier.
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

    def __init__(self, words, offsets, counts):
        # words: the whole segment as uint32; offsets[p]: index of pair p's first word; counts[p]: its records
        self.words, self.offsets, self.counts = words, offsets, counts

    def __len__(self):
        return len(self.counts)

    def records(self, p):
        """-> [n][3] uint32 view: feature_index_1, feature_index_2, integer Hamming distance."""
        o, n = int(self.offsets[p]), int(self.counts[p])
        return self.words[o:o + n * RECORD_WORDS].reshape(n, RECORD_WORDS)

    def pair(self, p):
        """-> (feature_index_1, feature_index_2, distance) of pair p; distance = d * (1.0 / 486) like the reference."""
        r = self.records(p)
        return r[:, 0].astype(np.uintp), r[:, 1].astype(np.uintp), r[:, 2] * (1.0 / 486)

    def total(self):
        return int(self.counts.sum())


class MatchGather:
    """Host gather of per-pair match lists to rank 0 through one shared-memory segment (see the module docstring).

    capacity_records = upper bound of the records any ONE rank contributes (e.g. its query rows), max_pairs = upper bound
    of the pairs of any one rank. Per rank the segment holds an index (offset and count of every pair of that rank, in
    the rank's pair order) followed by the records; only the pages actually written are ever touched. A rank hands
    buffers() to the runner (host.link_pairs(packed=...)), which writes both while it works; gather() is then a barrier
    plus, on rank 0, building the serial-order index."""

    def __init__(self, capacity_records, max_pairs, dist=None, name=None):
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        self.capacity, self.max_pairs = int(capacity_records), max(1, int(max_pairs))
        self.stride = 2 * 8 * self.max_pairs + self.capacity * RECORD_WORDS * 4  # bytes per rank
        self.stride = (self.stride + 4095) // 4096 * 4096
        nbytes = max(1, self.world * self.stride)
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

    def _views(self, r):
        base = r * self.stride
        index = np.ndarray((2, self.max_pairs), np.uint64, buffer=self.shm.buf, offset=base)
        records = np.ndarray((self.capacity * RECORD_WORDS,), np.uint32, buffer=self.shm.buf,
                             offset=base + 2 * 8 * self.max_pairs)
        return records, index[0], index[1]

    def buffers(self):
        """This rank's (records flat uint32, offsets uint64 [max_pairs], counts uint64 [max_pairs])."""
        return self._views(self.rank)

    def gather(self, pair_ids_by_rank, n_pairs):
        """Collective. pair_ids_by_rank[r] = the global ids of rank r's pairs in that rank's pair order (every rank can
        compute all of them: the partition is deterministic). Returns GatheredMatches on rank 0, None elsewhere."""
        if self.dist:
            self.dist.barrier()  # every rank's records and index are in the segment
            if self.rank != 0:
                return None
        offsets = np.full(n_pairs, -1, np.int64)
        counts = np.zeros(n_pairs, np.int64)
        rec_base = 2 * 8 * self.max_pairs // (RECORD_WORDS * 4)  # not a whole record in general: index in words instead
        for r, ids in enumerate(pair_ids_by_rank):
            ids = np.asarray(ids, np.int64)
            assert len(ids) <= self.max_pairs
            _, off, cnt = self._views(r)
            seen = offsets[ids] != -1
            assert not seen.any() and len(np.unique(ids)) == len(ids), "a pair was matched twice"
            assert int((off[:len(ids)] + cnt[:len(ids)]).max(initial=0)) <= self.capacity
            # word offset of the pair's first record in the whole segment
            offsets[ids] = (r * self.stride + 2 * 8 * self.max_pairs) // 4 + off[:len(ids)].astype(np.int64) * RECORD_WORDS
            counts[ids] = cnt[:len(ids)].astype(np.int64)
        del rec_base
        assert np.all(offsets >= 0), f"{int((offsets < 0).sum())} pairs were never matched"
        words = np.ndarray((len(self.shm.buf) // 4,), np.uint32, buffer=self.shm.buf)
        return GatheredMatches(words, offsets, counts)

    def close(self):
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:  # noqa: BLE001
            pass
