"""GPU parity tests for K1 (Hamming top-2) through the C ABI and through the C++ mirror of match_features_subset.
Bar: bit-exact indices, distances and output order versus the CPU oracle / the reference's own object code."""
import numpy as np
import pytest

from opencalibration_b200 import synthetic

pytestmark = pytest.mark.gpu

N_VARIANTS = 10


def k1t_launches(gpu):
    """Kernels of one tensor-core search: expansion, search, finish."""
    return 3


def assert_top2_equal(r, oracle_out):
    bk, bd, sd = oracle_out
    assert np.array_equal(r["best_k"], bk)
    assert np.array_equal(r["best_d"], bd)
    assert np.array_equal(r["second_d"], sd)


@pytest.fixture(autouse=True, params=[(1, 0), (2, 1), (2, 2), (2, 3), (2, 4)],
                ids=["int-pipes", "tensor-cores-form1", "tensor-cores-form2-m128", "tensor-cores-form2-m256",
                     "tensor-cores-form3-tmem"])
def engine(request, gpu):
    """Every test of this file runs on both engines of the single-pair search and on every form of the tensor-core
    search kernel (the batched entry points are K1 only and simply run again)."""
    gpu.set_option("k1_engine", request.param[0])
    gpu.set_option("k1t_variant", request.param[1])
    yield request.param[0]
    gpu.set_option("k1_engine", 0)
    gpu.set_option("k1t_variant", 0)


@pytest.fixture()
def default_options(gpu):
    yield
    gpu.set_option("k1_variant", 0)
    gpu.set_option("k1_items_per_sm", 32)
    gpu.set_option("k1_update", 0)


SHAPES = [(1, 1), (1, 2), (2, 1), (3, 63), (5, 64), (7, 65), (31, 129), (128, 1), (129, 200), (511, 513),
          (512, 4097), (513, 100), (1025, 1000), (2000, 37)]


@pytest.mark.parametrize("n1,n2", SHAPES)
def test_top2_matches_oracle(gpu, oracle, n1, n2):
    a, b = synthetic.config2_pair(n1, n2, seed=n1 * 131 + n2)
    if n2 > 8:  # planted duplicates: ties between best / second best and across tile or split boundaries
        b[n2 // 2] = b[1]
        b[n2 - 1] = b[1]
        b[3] = a[0]
        b[n2 - 2] = a[0]
    r, col = gpu.match_top2(a, b, cross_check=True)
    assert_top2_equal(r, oracle.match_top2(a, b))
    assert np.array_equal(col, oracle.match_col_best(a, b))


@pytest.mark.parametrize("n1,n2", [(40000, 100), (5000, 700), (130, 20000), (300, 33), (2049, 4097)])
def test_tensor_forms_on_span_shapes(gpu, oracle, engine, n1, n2):
    """Shapes that exercise the persistent spans of the second form of the tensor-core search: spans that cross from one
    query tile group into the next after one, two or many candidate steps, a single group shared by a hundred CTAs,
    fewer steps than SMs."""
    if engine != 2:
        pytest.skip("tensor-core engine only")
    a, b = synthetic.config2_pair(n1, n2, seed=n1 + 7 * n2)
    b[n2 - 1] = b[0]
    b[n2 // 3] = a[n1 - 1]
    a[n1 // 2] = a[0]  # equal queries: the cross-check must name the first one
    r, col = gpu.match_top2(a, b, cross_check=True)
    assert_top2_equal(r, oracle.match_top2(a, b))
    assert np.array_equal(col, oracle.match_col_best(a, b))
    assert_top2_equal(gpu.match_top2(a, b), oracle.match_top2(a, b))


def test_empty_and_degenerate_inputs(gpu, oracle):
    a, b = synthetic.config2_pair(40, 30, seed=1)
    r = gpu.match_top2(a, b[:0])  # no candidates: best = +inf, index 0 (match_features.cpp:74)
    assert np.all(r["best_k"] == 0) and np.all(r["best_d"] == 0xFFFF) and np.all(r["second_d"] == 0xFFFF)
    r = gpu.match_top2(a, b[:1])  # one candidate: second = +inf
    assert np.all(r["best_k"] == 0) and np.all(r["second_d"] == 0xFFFF)
    assert_top2_equal(r, oracle.match_top2(a, b[:1]))
    r = gpu.match_top2(a[:0], b)
    assert len(r) == 0
    r, col = gpu.match_top2(a[:0], b, cross_check=True)
    assert len(r) == 0 and np.all(col == 0xFFFFFFFF)
    r, col = gpu.match_top2(a, b[:0], cross_check=True)
    assert len(col) == 0
    same = np.repeat(a[:1], 50, axis=0)  # all candidates identical: best = first, second = same distance
    r = gpu.match_top2(a[:5], same)
    assert np.all(r["best_k"] == 0) and np.array_equal(r["best_d"], r["second_d"])
    assert_top2_equal(r, oracle.match_top2(a[:5], same))
    zeros, ones = np.zeros((3, 8), np.uint64), np.full((3, 8), np.uint64(2**64 - 1))
    r = gpu.match_top2(zeros, ones)  # maximum distance incl. the 26 padding bits of a raw 512-bit row
    assert np.all(r["best_d"] == 512)


@pytest.mark.parametrize("variant", range(N_VARIANTS))
def test_all_kernel_variants(gpu, oracle, default_options, variant):
    gpu.set_option("k1_variant", variant)
    for items in (1, 16, 64):
        gpu.set_option("k1_items_per_sm", items)
        for n1, n2 in ((700, 900), (130, 2049)):
            a, b = synthetic.config2_pair(n1, n2, seed=variant * 17 + n1)
            b[n2 - 1] = b[0]
            r, col = gpu.match_top2(a, b, cross_check=True)
            assert_top2_equal(r, oracle.match_top2(a, b))
            assert np.array_equal(col, oracle.match_col_best(a, b))


@pytest.mark.parametrize("update", [1, 2])
@pytest.mark.parametrize("variant", [0, 1, 8])
def test_both_update_forms(gpu, oracle, default_options, variant, update):
    """k1_update 1 = compare, vote and skip; 2 = branch-free two-smallest (the default at every run length)."""
    gpu.set_option("k1_variant", variant)
    gpu.set_option("k1_update", update)
    for items in (1, 32):
        gpu.set_option("k1_items_per_sm", items)
        for n1, n2 in ((513, 3000), (64, 70)):
            a, b = synthetic.config2_pair(n1, n2, seed=variant * 31 + update)
            b[n2 - 1] = b[0]
            b[n2 // 2] = b[0]
            r, col = gpu.match_top2(a, b, cross_check=True)
            assert_top2_equal(r, oracle.match_top2(a, b))
            assert np.array_equal(col, oracle.match_col_best(a, b))
            assert_top2_equal(gpu.match_top2(a, b), oracle.match_top2(a, b))


def test_split_merge_keeps_position_order(gpu, oracle, default_options):
    # the first minimum must win even when its duplicates fall into different candidate splits / tiles
    a, b = synthetic.config2_pair(64, 8192, seed=9)
    for k in (63, 64, 65, 1000, 4095, 4096, 8191):
        b[k] = b[5]
    b[7000] = a[10]
    b[100] = a[10]
    for items in (1, 4, 64, 256):
        gpu.set_option("k1_items_per_sm", items)
        assert_top2_equal(gpu.match_top2(a, b), oracle.match_top2(a, b))


def test_device_resident_entry_point(gpu, oracle, engine):
    import torch
    n1, n2 = 3000, 2500
    a, b = synthetic.config2_pair(n1, n2, seed=21)
    dq = torch.from_numpy(a.view(np.int64)).cuda()
    dc = torch.from_numpy(b.view(np.int64)).cuda()
    dout = torch.zeros(n1, dtype=torch.int64, device="cuda")
    dcol = torch.zeros(n2, dtype=torch.int32, device="cuda")
    wsb = gpu.match_top2_workspace_bytes(n1, n2, True)
    ws = torch.zeros(wsb + 512, dtype=torch.uint8, device="cuda")
    wsp = (ws.data_ptr() + 255) // 256 * 256
    before = gpu.kernel_launches()
    gpu.match_top2_device(dq.data_ptr(), n1, dc.data_ptr(), n2, dout.data_ptr(), dcol.data_ptr(), wsp, wsb,
                          torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    # K1: one fused kernel (top-2, split merge and cross-check); K1T: expansion, search, finish
    assert gpu.kernel_launches() - before == (1 if engine == 1 else k1t_launches(gpu))
    r = dout.cpu().numpy().view(gpu.TOP2_DTYPE)
    assert_top2_equal(r, oracle.match_top2(a, b))
    assert np.array_equal(dcol.cpu().numpy().view(np.uint32), oracle.match_col_best(a, b))
    with pytest.raises(gpu.OcbError):  # misaligned device pointer is refused, not silently fixed
        gpu.match_top2_device(dq.data_ptr() + 8, n1 - 1, dc.data_ptr(), n2, dout.data_ptr(), None, wsp, wsb, 0)


def test_strided_gather_entry_point(gpu, oracle):
    # rows embedded in 96-byte records at offset 24 (the layout of std::vector<feature_2d>), picked by index lists
    # with repeats and in arbitrary order -- what match_features_subset hands to the library
    a, b = synthetic.config2_pair(900, 1100, seed=21)
    rec1 = np.zeros((900, 96), np.uint8)
    rec2 = np.zeros((1100, 96), np.uint8)
    rec1[:, 24:88] = a.view(np.uint8).reshape(900, 64)
    rec2[:, 24:88] = b.view(np.uint8).reshape(1100, 64)
    rng = np.random.default_rng(3)
    idx1 = rng.integers(0, 900, 700).astype(np.uintp)
    idx2 = rng.permutation(1100)[:1000].astype(np.uintp)
    r, col = gpu.match_top2_strided(rec1.reshape(-1)[24:], 96, idx1, rec2.reshape(-1)[24:], 96, idx2, cross_check=True)
    assert_top2_equal(r, oracle.match_top2(a[idx1], b[idx2]))
    assert np.array_equal(col, oracle.match_col_best(a[idx1], b[idx2]))
    r = gpu.match_top2_strided(a.view(np.uint8).reshape(-1), 64, None, b.view(np.uint8).reshape(-1), 64, None)
    assert_top2_equal(r, oracle.match_top2(a, b))
    r = gpu.match_top2_strided(rec1.reshape(-1)[24:], 96, idx1[:0], rec2.reshape(-1)[24:], 96, idx2)
    assert len(r) == 0


def test_batched_pairs_match_single_calls(gpu, oracle):
    images, pos, pairs = synthetic.grid_survey(3, 3, 700, seed=5)
    images[4] = images[4][:333]  # ragged set sizes
    images[7] = images[7][:0]    # an image without features
    for i, d in enumerate(images):
        gpu.register_descriptors(1000 + i, d)
    plist = [(1000 + a, 1000 + b) for a, b in pairs]
    out, offs = gpu.match_pairs(plist, [len(images[a]) for a, _ in pairs])
    for p, (a, b) in enumerate(pairs):
        n = len(images[a])
        r = out[int(offs[p]):int(offs[p]) + n]
        assert_top2_equal(r, oracle.match_top2(images[a], images[b]))
    with pytest.raises(gpu.OcbError):
        gpu.match_pairs([(1000, 99999)], [len(images[0])])
    for i in range(len(images)):
        gpu.unregister_descriptors(1000 + i)
    with pytest.raises(gpu.OcbError):
        gpu.unregister_descriptors(1000)


def test_batched_registration_shares_one_allocation(gpu, oracle):
    a, b = synthetic.config2_pair(3000, 2500, seed=31)
    rec = np.zeros((3000, 96), np.uint8)
    rec[:, 24:88] = a.view(np.uint8).reshape(3000, 64)
    idx = np.random.default_rng(2).permutation(3000)[:1700].astype(np.uintp)
    gpu.register_descriptors_batch([(7001, rec.reshape(-1)[24:], 96, idx, len(idx)),
                                    (7002, b.view(np.uint8).reshape(-1), 64, None, len(b)),
                                    (7003, b.view(np.uint8).reshape(-1), 64, None, 0)])
    out, offs = gpu.match_pairs([(7001, 7002), (7002, 7001), (7003, 7001), (7001, 7003)], [1700, 2500, 0, 1700])
    assert_top2_equal(out[:1700], oracle.match_top2(a[idx], b))
    assert_top2_equal(out[1700:4200], oracle.match_top2(b, a[idx]))
    assert np.all(out[4200:]["best_d"] == 0xFFFF)  # no candidates
    gpu.unregister_descriptors(7002)  # the shared allocation must survive for the remaining sets
    out, _ = gpu.match_pairs([(7001, 7001)], [1700])
    assert_top2_equal(out, oracle.match_top2(a[idx], a[idx]))
    gpu.register_descriptors(7001, b[:10])  # replacing a batched set by a plain one
    out, _ = gpu.match_pairs([(7001, 7001)], [10])
    assert np.all(out["best_d"] == 0)
    for sid in (7001, 7003):
        gpu.unregister_descriptors(sid)


def test_batched_registration_rejects_duplicate_ids(gpu):
    # the same id twice in one batch used to free the batch's shared allocation while it was being entered
    a, _ = synthetic.config2_pair(64, 8, seed=3)
    flat = a.view(np.uint8).reshape(-1)
    with pytest.raises(gpu.OcbError, match="duplicate"):
        gpu.register_descriptors_batch([(7101, flat, 64, None, 64), (7102, flat, 64, None, 64),
                                        (7101, flat, 64, None, 32)])
    with pytest.raises(gpu.OcbError):  # nothing of the rejected batch was entered
        gpu.unregister_descriptors(7102)


def test_config1_pair_through_the_mirror(gpu, hostlib, config1, golden):
    # BASELINE configs[0]: test_match's flow on the repo pair; golden = the reference's own match_features.cpp
    ia = hostlib.spatially_subsample_feature_indices(config1["a_xy"], config1["a_strength"], 40.0)
    ib = hostlib.spatially_subsample_feature_indices(config1["b_xy"], config1["b_strength"], 40.0)
    assert len(ia) == 4052 and len(ib) == 4198
    i1, i2, d = hostlib.match_features_subset(config1["a_desc"], config1["b_desc"], ia, ib)
    assert len(i1) == 500
    assert np.array_equal(i1, golden["c1_m1"]) and np.array_equal(i2, golden["c1_m2"])
    assert np.array_equal(d, golden["c1_dist"])
    # test/test_match.cpp:26-42
    assert set(i1.tolist()) <= set(ia.tolist()) and set(i2.tolist()) <= set(ib.tolist())


def test_mirror_equals_reference_code_on_subsets(gpu, hostlib, oracle, golden):
    i1, i2, d = hostlib.match_features_subset(golden["c2s_a"], golden["c2s_b"], golden["c2s_i1"], golden["c2s_i2"])
    assert np.array_equal(i1, golden["c2s_m1"]) and np.array_equal(i2, golden["c2s_m2"])
    assert np.array_equal(d, golden["c2s_dist"])
    # repeated / unsorted indices are passed through untouched
    a, b = golden["c2s_a"], golden["c2s_b"]
    idx1 = np.array([5, 5, 900, 3, 5, 0], np.uintp)
    idx2 = np.array([7, 7, 1, 999, 500, 17], np.uintp)
    got = hostlib.match_features_subset(a, b, idx1, idx2)
    want = oracle.match_features_subset(a, b, idx1, idx2)
    assert all(np.array_equal(x, y) for x, y in zip(got, want))
    e = hostlib.match_features_subset(a, b, idx1, idx2[:0])
    assert len(e[0]) == 0
    e = hostlib.match_features_subset(a, b, idx1[:0], idx2)
    assert len(e[0]) == 0


def test_cross_check_flags(gpu, hostlib, oracle):
    a, b = synthetic.config2_pair(900, 800, seed=77)
    i1 = np.arange(900, dtype=np.uintp)
    i2 = np.arange(800, dtype=np.uintp)
    m1, m2, d, mutual = hostlib.match_features_subset(a, b, i1, i2, cross_check=True)
    p1, p2, pd = hostlib.match_features_subset(a, b, i1, i2)
    assert np.array_equal(m1, p1) and np.array_equal(m2, p2) and np.array_equal(d, pd)  # same list, same order
    col = oracle.match_col_best(a, b)
    assert np.array_equal(mutual, col[m2] == m1)
    assert mutual.sum() > 100


def test_full_size_properties_10k(gpu, oracle):
    # BASELINE configs[1] at full size: EVERY query and EVERY candidate column against the oracle (bit-exact), then
    # size-independent properties
    n = 10000
    a, b = synthetic.config2_pair(n, n, seed=1)
    r, col = gpu.match_top2(a, b, cross_check=True)
    bk, bd, sd = oracle.match_top2(a, b)
    assert np.array_equal(r["best_k"], bk) and np.array_equal(r["best_d"], bd) and np.array_equal(r["second_d"], sd)
    assert np.array_equal(col, oracle.match_col_best(a, b))
    assert np.all(r["best_d"] <= r["second_d"])
    # self match: every row finds itself at distance 0, at its first occurrence
    rs = gpu.match_top2(a, a)
    assert np.all(rs["best_d"] == 0) and np.array_equal(rs["best_k"], np.arange(n, dtype=np.uint32))
    # permuting the candidates permutes best_k and leaves the distances alone (no exact ties in random rows
    # between best and second at this noise level would change distances anyway)
    perm = np.random.default_rng(1).permutation(n)
    rp = gpu.match_top2(a, b[perm])
    assert np.array_equal(rp["best_d"], r["best_d"]) and np.array_equal(rp["second_d"], r["second_d"])
    untied = r["best_d"] < r["second_d"]
    assert np.array_equal(perm[rp["best_k"][untied]], r["best_k"][untied])
    # cross-check direction equals the swapped forward search
    rb = gpu.match_top2(b, a)
    assert np.array_equal(col, rb["best_k"])
    # checksum of the ratio-test survivors is reproducible across kernel variants
    keep = r["best_d"].astype(np.float64) * (1.0 / 486) < 0.8 * (r["second_d"].astype(np.float64) * (1.0 / 486))
    assert 4000 < keep.sum() < 6000


def test_full_size_10k_through_the_mirror_equals_the_reference_object_code(gpu, hostlib):
    # configs[1] through the reference-facing entry point against the reference's own compiled match_features.cpp
    # (-mpopcnt build of the same translation unit), every emitted match: indices, distances and sort order
    import oc_oracle
    if not oc_oracle.Reference.available():
        pytest.skip("oracle/_ref not built")
    ref = oc_oracle.Reference(popcnt=True)
    n = 10000
    a, b = synthetic.config2_pair(n, n, seed=1)
    idx = np.arange(n, dtype=np.uintp)
    got = hostlib.match_features_subset(a, b, idx, idx)
    want = ref.match_features_subset(a, b, idx, idx)
    assert len(want[0]) > 4000
    assert all(np.array_equal(x, y) for x, y in zip(got, want))


def test_engines_agree_and_auto_picks_by_size(gpu, oracle):
    """The single-pair search has two engines with one contract: K1 on the integer pipes (XOR + POPC) and K1T on the
    tensor cores (exact s8 contraction, tcgen05.mma.kind::i8, TMEM accumulators). Same records bit for bit, ties and
    cross-check included; by default pairs from 512 x 512 rows up go to the tensor cores."""
    a, b = synthetic.config2_pair(3000, 2777, seed=17)
    b[1500:1540] = b[:40]   # exact ties across candidate tiles and candidate ranges
    b[2700] = a[5]
    want, want_col = oracle.match_top2(a, b), oracle.match_col_best(a, b)
    launches = {}
    for engine in (1, 2, 0):
        gpu.set_option("k1_engine", engine)
        before = gpu.kernel_launches()
        r, col = gpu.match_top2(a, b, cross_check=True)
        launches[engine] = gpu.kernel_launches() - before
        assert_top2_equal(r, want)
        assert np.array_equal(col, want_col)
    gpu.set_option("k1_engine", 0)
    # K1: one kernel; K1T: expand, search, finish
    assert launches[1] == 1 and launches[2] == k1t_launches(gpu) and launches[0] == k1t_launches(gpu)
    before = gpu.kernel_launches()
    gpu.match_top2(a[:100], b[:100])
    assert gpu.kernel_launches() - before == 1  # small pairs stay on the integer pipes


def test_concurrent_callers(gpu, oracle):
    # the reference calls this path from many OpenMP workers at once (pipeline.cpp:42-49)
    import threading
    pairs = [synthetic.config2_pair(300 + 37 * t, 500 + 11 * t, seed=t) for t in range(8)]
    want = [oracle.match_top2(a, b) for a, b in pairs]
    got, errs = [None] * 8, []

    def work(t):
        try:
            for _ in range(5):
                got[t] = gpu.match_top2(*pairs[t])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    for t in range(8):
        assert_top2_equal(got[t], want[t])


def test_concurrent_batched_submissions(gpu, oracle):
    """ocb_match_pairs from several host threads at once (include/ocb.h: host-buffer entry points are thread-safe;
    each thread has its own streams and staging areas, the registered sets are shared)."""
    import threading
    images, pos, pairs = synthetic.grid_survey(3, 4, 600, seed=11)
    for i, d in enumerate(images):
        gpu.register_descriptors(3000 + i, d)
    plist = [(3000 + a, 3000 + b) for a, b in pairs]
    nq = [len(images[a]) for a, _ in pairs]
    want, offs = gpu.match_pairs(plist, nq)
    for p, (a, b) in enumerate(pairs[:6]):
        assert_top2_equal(want[int(offs[p]):int(offs[p]) + nq[p]], oracle.match_top2(images[a], images[b]))
    errs, got = [], [None] * 4

    def work(t):
        try:
            for rep in range(6):
                lo = (t * 7 + rep * 3) % (len(plist) - 5)
                out, _ = gpu.match_pairs(plist[lo:], nq[lo:])
                if not np.array_equal(out, want[int(offs[lo]):]):
                    errs.append(f"thread {t} rep {rep}: records differ")
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(len(images)):
        gpu.unregister_descriptors(3000 + i)
    assert not errs, errs
