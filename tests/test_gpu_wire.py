"""LinkStage over a graph.json checkpoint (ocbw_graph_link, include/ocb_wire.h): read the document, match + RANSAC the
requested pairs on the GPU from the document's own features and camera models, write the edges back in the reference's
format. Every edge must equal the per-pair reference flow (the oracle), the written document must survive the
reference's own deserialize() -> serialize() byte for byte, and the edge ids must be the ones MeasurementGraph::addEdge
would draw."""
import numpy as np
import pytest

import oc_ref_io as R
from opencalibration_b200 import synthetic, wire
from test_gpu_link import check_pair, expected_pair

pytestmark = pytest.mark.gpu


def survey_document(survey, imgs, num_sparse):
    g = wire.Graph()
    cam, dims = survey.camera8(), np.array(survey.image_size, np.uint64)
    ids = []
    for i, (d, xy, s) in enumerate(imgs):
        pose = np.array([survey.positions[i][0], survey.positions[i][1], 10.0, 0, 0, 0, 1.0])
        ids.append(g.add_node(cam, dims, xy, s, d, num_sparse=num_sparse[i] or len(d), path="img_%03d.JPG" % i,
                              pose=pose))
    return g, ids


def test_link_over_a_checkpoint(gpu, hostlib, oracle):
    survey = synthetic.PlanarSurvey(2, 3, 1200, seed=5)
    imgs = [survey.image(i) for i in range(survey.n_images)]
    imgs[3] = tuple(x[:300] for x in imgs[3])
    num_sparse = [0, 600, 0, 0, 0, 0]
    g0, ids = survey_document(survey, imgs, num_sparse)
    text0 = g0.serialize()

    g = wire.Graph(text0)  # from the text, like a run resumed from a checkpoint
    order = {g.node(i, False)["id"]: i for i in range(g.num_nodes)}
    pairs = survey.pairs[:14]
    stats = g.link([(ids[a], ids[b]) for a, b in pairs], threads=4)
    assert g.num_edges == len(pairs) and stats["seconds_total"] > 0
    cam = survey.camera8()
    kept = 0
    for p, (a, b) in enumerate(pairs):
        e = g.edge(p)
        assert (e["source"], e["dest"]) == (ids[a], ids[b])
        exp = expected_pair(oracle, imgs[a], imgs[b], cam, num_sparse=(num_sparse[a], num_sparse[b]))
        got = dict(matches=tuple(np.asarray(v) for v in e["matches"]), H=e["relation"], relation_type=e["relation_type"],
                   poses=np.concatenate([e["poses"][:, 1:], e["poses"][:, :1]], axis=1),  # link_get order: q, t, score
                   inlier_pixels=e["inlier_pixels"], inlier_idx=e["inlier_idx"])
        check_pair(got, exp)
        kept += bool(exp["keep"])
    assert kept >= 8
    # the nodes know their edges
    text1 = g.serialize()
    g1 = wire.Graph(text1)
    assert g1.serialize() == text1 and g1.num_edges == len(pairs)
    # edge ids: the draws of a fresh default-seeded generator, in pair order (graph.hpp:86-100)
    draws = wire.Graph()
    z = (np.zeros((0, 2)), np.zeros(0, np.float32), np.zeros((0, 8), np.uint64))
    expect_ids = [draws.add_node(np.zeros(8), np.zeros(2, np.uint64), *z) for _ in pairs]
    assert [g.edge(p)["id"] for p in range(len(pairs))] == expect_ids
    if R.available():
        again, equal = R.roundtrip(text1)
        assert again == text1 and equal
    # linking the same pairs again replaces the edges in place: same document
    g.link([(ids[a], ids[b]) for a, b in pairs[:5]], threads=2)
    assert g.serialize() == text1
    # unknown node id
    with pytest.raises(wire.OcbError):
        g.link([(ids[0], 12345)])


def test_link_matches_only(gpu, hostlib, oracle):
    survey = synthetic.PlanarSurvey(1, 3, 800, seed=9)
    imgs = [survey.image(i) for i in range(3)]
    g, ids = survey_document(survey, imgs, [0, 0, 0])
    g.link([(ids[0], ids[1]), (ids[2], ids[1])], run_ransac=False)
    for p, (a, b) in enumerate([(0, 1), (2, 1)]):
        ia = oracle.subsample(imgs[a][1], imgs[a][2], 40.0, 0)
        ib = oracle.subsample(imgs[b][1], imgs[b][2], 40.0, 0)
        m1, m2, md = oracle.match_features_subset(imgs[a][0], imgs[b][0], ia, ib)
        e = g.edge(p)
        assert np.array_equal(e["matches"][0], m1) and np.array_equal(e["matches"][1], m2)
        assert np.array_equal(e["matches"][2], md) and len(md) > 50
