"""Guided matcher of the dense stage (reference src/dense/dense_stereo.cpp:244-281): K4 through the C ABI and through
the C++ mirror, against the CPU oracle; the synthetic candidate lists against the reference's own jk-tree.
Bar: bit-exact list positions, integer distances, accepted set and double distances."""
import numpy as np
import pytest

from opencalibration_b200 import synthetic

BITS = 486


def brute_force_lists(q, c, list_query, begin, nearby):
    """numpy restatement of dense_stereo.cpp:251-273 on integer distances (checks the C++ oracle itself)."""
    q8, c8 = q.view(np.uint8).reshape(len(q), 64), c.view(np.uint8).reshape(len(c), 64)
    n = len(list_query)
    bp, bd, sd = np.zeros(n, np.uint32), np.full(n, np.inf), np.full(n, np.inf)
    for l in range(n):
        idx = nearby[int(begin[l]):int(begin[l + 1])]
        if len(idx) == 0:
            continue
        d = np.unpackbits(q8[list_query[l]][None, :] ^ c8[idx], axis=1).sum(axis=1)
        order = np.lexsort((np.arange(len(d)), d))  # (distance, position): first minimum, second with multiplicity
        bp[l], bd[l] = order[0], d[order[0]] * (1.0 / BITS)
        if len(d) > 1:
            sd[l] = d[order[1]] * (1.0 / BITS)
    return bp, bd, sd


def as_top2(bp, bd, sd):
    to_int = lambda x: np.where(np.isinf(x), 0xFFFF, np.rint(x * BITS)).astype(np.uint16)
    return bp, to_int(bd), to_int(sd)


def edge_case_lists(rng, n_q, n_c):
    """Lists that hit every boundary of the kernel's 8-lane groups and 2-deep unrolling, plus ties."""
    lengths = [0, 1, 2, 7, 8, 9, 15, 16, 17, 23, 24, 25, 31, 32, 33, 100, 0, 257, 1, 5000 if n_c >= 64 else 3]
    lists = []
    for k, ln in enumerate(lengths):
        idx = rng.integers(0, n_c, ln).astype(np.uint32)  # with replacement: duplicate rows = exact ties
        if ln >= 4 and k % 2 == 0:
            idx[ln // 2] = idx[0]
            idx[ln - 1] = idx[0]
        lists.append(idx)
    begin = np.zeros(len(lists) + 1, np.uint64)
    begin[1:] = np.cumsum([len(x) for x in lists])
    nearby = np.concatenate(lists).astype(np.uint32)
    list_query = rng.integers(0, n_q, len(lists)).astype(np.uint32)
    list_query[3] = list_query[2]  # the same source feature visited twice
    return list_query, begin, nearby


# ---------------------------------------------------------------------------------------------------------------
# CPU: the oracle and the workload generator
# ---------------------------------------------------------------------------------------------------------------
def test_oracle_lists_against_numpy(oracle):
    rng = np.random.default_rng(2)
    a, b = synthetic.config2_pair(60, 300, seed=5)
    lq, begin, nearby = edge_case_lists(rng, 60, 300)
    bp, bd, sd, good = oracle.match_lists(a, b, lq, begin, nearby)
    ebp, ebd, esd = brute_force_lists(a, b, lq, begin, nearby)
    assert np.array_equal(bp, ebp) and np.array_equal(bd, ebd) and np.array_equal(sd, esd)
    ln = np.diff(begin.astype(np.int64))
    expect = np.where(ln >= 2, bd < 0.85 * sd, (ln == 1) & (bd < 0.35))  # dense_stereo.cpp:275-276
    assert np.array_equal(good, expect)


def test_oracle_full_list_equals_dense_match(oracle):
    """A list holding every candidate in position order is match_features_subset's inner loop (same update rule)."""
    a, b = synthetic.config2_pair(50, 333, seed=9)
    b[100] = b[7]
    begin = (np.arange(51) * 333).astype(np.uint64)
    nearby = np.tile(np.arange(333, dtype=np.uint32), 50)
    bp, bd, sd, _ = oracle.match_lists(a, b, np.arange(50, dtype=np.uint32), begin, nearby)
    bk, ibd, isd = oracle.match_top2(a, b)
    assert np.array_equal(bp, bk)
    assert np.array_equal(bd, ibd * (1.0 / BITS)) and np.array_equal(sd, isd * (1.0 / BITS))


def test_generated_lists_equal_the_reference_kdtree(reference):
    """synthetic.guided_visits builds its candidate lists with scipy; the reference builds them with its vendored
    jk-tree (dense_stereo.cpp:127-131,244-246). Same members, same order."""
    for n_q, n_c, seed in ((500, 800, 1), (300, 5000, 2), (50, 3, 3)):
        w = synthetic.guided_visits(n_q, n_c, seed=seed)
        begin, nearby = reference.radius_lists(w["cand_xy"], w["pred_xy"], 150.0)
        assert np.array_equal(begin, w["begin"]) and np.array_equal(nearby, w["nearby"])


def test_lists_entry_points_fail_without_device(built):
    """No CPU fallback: in the build container (no GPU) the call must fail loudly, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from opencalibration_b200 import capi
    w = synthetic.guided_visits(20, 50, seed=1)
    with pytest.raises(capi.OcbError):
        capi.match_lists(w["q"], w["c"], np.arange(20, dtype=np.uint32), w["begin"], w["nearby"])


# ---------------------------------------------------------------------------------------------------------------
# GPU: K4 parity
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_k4_edge_cases_match_oracle(gpu, oracle):
    rng = np.random.default_rng(4)
    a, b = synthetic.config2_pair(70, 900, seed=12)
    lq, begin, nearby = edge_case_lists(rng, 70, 900)
    r = gpu.match_lists(a, b, lq, begin, nearby)
    bp, bd, sd = as_top2(*oracle.match_lists(a, b, lq, begin, nearby)[:3])
    assert np.array_equal(r["best_k"], bp) and np.array_equal(r["best_d"], bd) and np.array_equal(r["second_d"], sd)
    empty = np.diff(begin.astype(np.int64)) == 0
    assert np.all(r["best_k"][empty] == 0) and np.all(r["best_d"][empty] == 0xFFFF)


@pytest.mark.gpu
@pytest.mark.parametrize("n_q,n_c,seed", [(1, 1, 1), (33, 40, 2), (1000, 3000, 3), (4097, 20000, 4)])
def test_k4_dense_workload_matches_oracle(gpu, oracle, n_q, n_c, seed):
    w = synthetic.guided_visits(n_q, n_c, seed=seed)
    lq = np.arange(n_q, dtype=np.uint32)
    r = gpu.match_lists(w["q"], w["c"], lq, w["begin"], w["nearby"])
    bp, bd, sd = as_top2(*oracle.match_lists(w["q"], w["c"], lq, w["begin"], w["nearby"])[:3])
    assert np.array_equal(r["best_k"], bp) and np.array_equal(r["best_d"], bd) and np.array_equal(r["second_d"], sd)


@pytest.mark.gpu
def test_k4_full_lists_equal_k1(gpu):
    """Size-independent property: with every candidate listed in position order K4 must reproduce K1."""
    a, b = synthetic.config2_pair(300, 2500, seed=21)
    begin = (np.arange(301) * 2500).astype(np.uint64)
    nearby = np.tile(np.arange(2500, dtype=np.uint32), 300)
    r4 = gpu.match_lists(a, b, np.arange(300, dtype=np.uint32), begin, nearby)
    r1 = gpu.match_top2(a, b)
    assert np.array_equal(r4, r1)


@pytest.mark.gpu
def test_k4_rejects_bad_indices(gpu):
    a, b = synthetic.config2_pair(10, 10, seed=1)
    begin = np.array([0, 2], np.uint64)
    with pytest.raises(gpu.OcbError):
        gpu.match_lists(a, b, np.array([0], np.uint32), begin, np.array([1, 10], np.uint32))
    with pytest.raises(gpu.OcbError):
        gpu.match_lists(a, b, np.array([10], np.uint32), begin, np.array([1, 2], np.uint32))
    with pytest.raises(gpu.OcbError):
        gpu.match_lists(a, b, np.array([0], np.uint32), np.array([1, 2], np.uint64), np.array([1, 2], np.uint32))
    assert len(gpu.match_lists(a, b, np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.uint32))) == 0


@pytest.mark.gpu
def test_guided_mirror_matches_oracle(gpu, hostlib, oracle):
    """match_features_guided (C++ mirror): accepted visits, candidate feature indices and double distances."""
    w = synthetic.guided_visits(2000, 6000, seed=8)
    n_sparse_src, n_sparse_cand = 37, 91  # dense features are a tail of the feature vector (dense_stereo.cpp:127-131)
    rng = np.random.default_rng(0)
    src = np.concatenate([synthetic.random_descriptors(n_sparse_src, rng), w["q"]])
    cand = np.concatenate([synthetic.random_descriptors(n_sparse_cand, rng), w["c"]])
    query_feature = np.arange(2000, dtype=np.uintp) + n_sparse_src
    nearby = w["nearby"].astype(np.uintp) + n_sparse_cand
    ol, oq, oc, ob, os_ = hostlib.match_features_guided(src, cand, query_feature, w["begin"].astype(np.uintp), nearby)
    bp, bd, sd, good = oracle.match_lists(w["q"], w["c"], np.arange(2000, dtype=np.uint32), w["begin"], w["nearby"])
    lists = np.flatnonzero(good)
    assert len(lists) > 800
    assert np.array_equal(ol, lists) and np.array_equal(oq, query_feature[lists])
    assert np.array_equal(oc, nearby[(w["begin"][lists] + bp[lists]).astype(np.int64)])
    assert np.array_equal(ob, bd[lists]) and np.array_equal(os_, sd[lists])
