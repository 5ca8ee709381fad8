"""CPU: the reference's own Scala sampler tests (scala/subgraph_sampler/src/test/scala/SGSPureSparkV1TaskTest.scala),
restated on the same mock data against the oracle's sampler and the product's host-side hydration / encoder.  The
reference runs these with the non-reproducible shuffle, so - like there - the assertions are the properties a uniform
sample must have; tests/test_gpu_sampler.py holds the CUDA kernel bit-exact to this oracle."""
import numpy as np
import pytest

from gigl_b200 import sample_io as sio
from oracle import oracle as O
from test_sample_assembly import np_edge_rows

# mockUnhydratedEdgeForCurrentTest / mockHydratedEdgeForCurrentTest (SGSPureSparkV1TaskTest.scala:44-97): (src, dst, feature)
EDGES = [(0, 1, 0.5), (0, 2, 1.0), (0, 3, 1.5), (0, 4, 2.0), (0, 5, 2.5), (0, 6, 3.0), (0, 7, 3.5), (0, 8, 4.0), (1, 2, 1.5), (1, 3, 2.0),
         (1, 0, 0.5), (2, 0, 1.0), (3, 0, 1.5), (4, 0, 2.0), (5, 0, 2.5), (6, 0, 3.0), (7, 0, 3.5), (8, 0, 4.0), (2, 1, 1.5), (3, 1, 2.0)]
# mockHydratedNodeForCurrentTest (:99-121): node id -> feature; 9, 10, 11 are isolated
NODE_FEAT = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.01, 0.11]
N = len(NODE_FEAT)


@pytest.fixture(scope="module")
def graph():
    src = np.array([e[0] for e in EDGES])
    dst = np.array([e[1] for e in EDGES])
    ef = np.array([[e[2]] for e in EDGES], dtype=np.float32)
    rowptr, col = O.np_build_in_csr(src, dst, N, True)  # the mock edge list already holds both directions
    return src, dst, ef, rowptr, col


def test_onehop_samples_are_valid(graph):
    """:190-213 - node 0 has 8 in-neighbours > numNeighborsToSample = 3: the sample is 3 of {1..8}."""
    _, _, _, rowptr, col = graph
    nbr, cnt = O.c_sample_khop(rowptr, col, np.array([0], dtype=np.int32), [3])
    assert cnt[0][0] == 3 and len(set(nbr[0].tolist())) == 3 and set(nbr[0].tolist()) <= set(range(1, 9))
    # a node with <= 3 in-neighbours takes all of them (no randomness): IN(1) = {0, 2, 3}
    nbr, cnt = O.c_sample_khop(rowptr, col, np.array([1], dtype=np.int32), [3])
    assert cnt[0][0] == 3 and sorted(nbr[0].tolist()) == [0, 2, 3]


def test_twohop_samples_are_valid(graph):
    """:241-271 - zero-hop 1 (in-degree 3: all of IN(1) = {3, 0, 2} is taken), one-hop 0 (in-degree 8): three of {1..8}."""
    _, _, _, rowptr, col = graph
    nbr, cnt = O.c_sample_khop(rowptr, col, np.array([1], dtype=np.int32), [3, 3])
    hop1 = nbr[0].tolist()
    assert sorted(hop1) == [0, 2, 3]
    j = hop1.index(0)
    two = nbr[1][j * 3:(j + 1) * 3].tolist()
    assert cnt[1][j] == 3 and len(set(two)) == 3 and set(two) <= set(range(1, 9))
    # the other one-hop nodes have in-degree 2: both neighbours, one empty slot
    for k, want in ((2, [0, 1]), (3, [0, 1])):
        jj = hop1.index(k)
        assert sorted(v for v in nbr[1][jj * 3:(jj + 1) * 3].tolist() if v >= 0) == want and cnt[1][jj] == 2


def test_rooted_node_neighborhood_is_valid(graph):
    """:391-507 - createSubgraph for root 0, numNeighborsToSample = 3: every sampled one-hop source has its two-hop edges,
    the node ids of _neighbor_nodes are exactly the ids on _neighbor_edges, the root is in the neighbourhood with its
    own feature (0.0); edges carry the features of mockHydratedEdgeForCurrentTest."""
    src, dst, ef, rowptr, col = graph
    roots = np.arange(N, dtype=np.int32)
    nbr, cnt = O.c_sample_khop(rowptr, col, roots, [3, 3])
    x = np.array(NODE_FEAT, dtype=np.float32)[:, None]
    data, offs = sio.encode_samples(roots, [3, 3], nbr, x, kind="rnn", csr=(rowptr, col), edge_rows=np_edge_rows(src, dst, N, True), edge_feat=ef)
    recs = [sio.parse_sample(r) for r in sio.split_tfrecords(data, verify=True)]
    assert len(recs) == N
    s = recs[0]
    assert s["root_node"]["node_id"] == 0 and s["root_node"]["feature_values"] == [0.0]
    edges = [(e["src_node_id"], e["dst_node_id"]) for e in s["edges"]]
    onehop = [a for a, b in edges if b == 0]
    assert len(onehop) == 3 and set(onehop) <= set(range(1, 9))
    for k in onehop:  # "for each of onehopSrcIds verify that twohopDstId exists"
        assert any(b == k for _, b in edges)
    assert sorted({v for e in edges for v in e}) == sorted(v["node_id"] for v in s["nodes"])
    root_in_nodes = [v for v in s["nodes"] if v["node_id"] == 0]
    assert len(root_in_nodes) == 1 and root_in_nodes[0]["feature_values"] == [0.0]
    feat = {(a, b): f for a, b, f in EDGES}
    for e in s["edges"]:
        assert e["feature_values"] == [feat[(e["src_node_id"], e["dst_node_id"])]]
    for v in s["nodes"]:
        assert np.float32(v["feature_values"][0]) == np.float32(NODE_FEAT[v["node_id"]])


def test_isolated_nodes_are_included(graph):
    """:509-580 - every node gets a RootedNodeNeighborhood; the isolated ones (9, 10, 11 here; 4, 5 in the reference's
    second mock graph) come with no edges and themselves as the only node."""
    src, dst, ef, rowptr, col = graph
    roots = np.arange(N, dtype=np.int32)
    nbr, _ = O.c_sample_khop(rowptr, col, roots, [3, 3])
    data, _ = sio.encode_samples(roots, [3, 3], nbr, np.array(NODE_FEAT, dtype=np.float32)[:, None], kind="rnn")
    recs = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(data, verify=True))}
    assert sorted(recs) == list(range(N))
    assert sorted(r for r, s in recs.items() if not s["edges"]) == [9, 10, 11]
    for r in (9, 10, 11):
        assert [v["node_id"] for v in recs[r]["nodes"]] == [r]
    # the reference's second mock graph (:511-529): edges among 0..3, nodes 0..5 -> 4 and 5 are isolated
    e2 = [(0, 1), (0, 2), (0, 3), (1, 3), (1, 0), (2, 0), (3, 0), (3, 1)]
    rp2, col2 = O.np_build_in_csr(np.array([a for a, _ in e2]), np.array([b for _, b in e2]), 6, True)
    nbr2, _ = O.c_sample_khop(rp2, col2, np.arange(6, dtype=np.int32), [3, 3])
    data2, _ = sio.encode_samples(np.arange(6, dtype=np.int32), [3, 3], nbr2, None, kind="rnn")
    recs2 = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(data2, verify=True))}
    assert sorted(r for r, s in recs2.items() if not s["edges"]) == [4, 5]


# ---- NodeAnchorBasedLinkPredictionTaskTest.scala (same mock graph) ----------------------------------------------------
def test_positive_samples_are_valid(graph):
    """:127-155 - sampleDstNodesUniformly, numPositiveSamples = 2: node 0 has 8 out-edges, its positives are 2 of {1..8}."""
    src, dst, _, _, _ = graph
    orow, ocol = O.np_build_in_csr(dst, src, N, True)  # out-CSR: row u = sorted destinations of u
    pos = O.np_sample_positives(orow, ocol, np.arange(N), 2)
    assert len(pos[0]) == 2 and len(set(pos[0])) == 2 and set(pos[0]) <= set(range(1, 9))
    assert sorted(pos[4]) == [0]                      # out-degree 1 <= 2: the only destination
    assert all(u not in pos for u in (9, 10, 11))     # no out-edge, no positive


def test_positive_neighbourhoods_and_output_validation(graph):
    """:157-194 lookupDstNodeNeighborhood keeps the direction anchor -> positive and attaches the POSITIVE's own
    neighbourhood; :196-230 TaskOutputValidator.validateMainSamples: both endpoints of every supervision edge are among
    the neighbourhood's nodes.  Checked on the product's host encoder fed with the oracle's samples."""
    src, dst, ef, rowptr, col = graph
    orow, ocol = O.np_build_in_csr(dst, src, N, True)
    roots = np.arange(N, dtype=np.int32)
    nbr, _ = O.c_sample_khop(rowptr, col, roots, [3, 3])
    pos_of = O.np_sample_positives(orow, ocol, roots, 2)
    pos = np.full((N, 2), -1, dtype=np.int32)
    for u, lst in pos_of.items():
        pos[u, :len(lst)] = lst
    tree = np.where(pos >= 0, pos, -1).astype(np.int64)  # every positive is a root of this call: its tree index is its id
    x = np.array(NODE_FEAT, dtype=np.float32)[:, None]
    er = np_edge_rows(src, dst, N, True)
    data, offs = sio.encode_samples(roots, [3, 3], nbr, x, kind="nablp", csr=(rowptr, col), edge_rows=er, edge_feat=ef, pos=pos, pos_tree=tree)
    got = {s["root_node"]["node_id"]: s for s in map(sio.parse_nablp_sample, sio.split_tfrecords(data, verify=True))}
    assert sorted(got) == sorted(pos_of)                      # anchors = the nodes with an out-edge; 9, 10, 11 emit nothing
    rnn_data, _ = sio.encode_samples(roots, [3, 3], nbr, x, kind="rnn", csr=(rowptr, col), edge_rows=er, edge_feat=ef)
    rnn = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(rnn_data, verify=True))}
    key = lambda e: (e["src_node_id"], e["dst_node_id"])  # noqa: E731
    for u, s in got.items():
        assert sorted(key(e) for e in s["pos_edges"]) == sorted((u, p) for p in pos_of[u])   # direction: _src_node -> _pos_dst_node
        have_e, have_n = {key(e) for e in s["edges"]}, {v["node_id"] for v in s["nodes"]}
        for p in pos_of[u]:                                                                    # the positive's neighbourhood rides along
            assert {key(e) for e in rnn[p]["edges"]} <= have_e and {v["node_id"] for v in rnn[p]["nodes"]} <= have_n
        assert {key(e) for e in rnn[u]["edges"]} <= have_e
        for e in s["pos_edges"]:                                                               # validateMainSamples
            assert e["src_node_id"] in have_n and e["dst_node_id"] in have_n
        for e in s["edges"]:
            assert e["src_node_id"] in have_n and e["dst_node_id"] in have_n
        assert not s["hard_neg_edges"] and not s["neg_edges"]
