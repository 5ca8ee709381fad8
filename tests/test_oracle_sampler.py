"""Pins the CPU oracle (sampling half) before anything is compared against it.

Goldens: tests/golden/xxh64_kat.json (Spark xxhash64 KAT + independent XXH64 implementation),
tests/golden/snc16_* and nablp27_* (the reference sampler's own fixture inputs / outputs,
extracted by tests/golden/make_golden.py).
"""
import numpy as np
import pytest

from helpers import load_golden, powerlaw_edges, uniform_edges
from oracle import oracle as orc


def test_xxh64_known_answers():
    kat = load_golden("xxh64_kat.json")
    for v, want in kat["hash_int_seed42"]:
        assert orc.c_xxh64_int(v) == want
        assert int(orc.np_xxh64_int(np.array([v]))[0]) == want
    # the three values quoted in SURVEY.md / BASELINE.md
    got = [orc.c_xxh64_int(v) for v in (43, 44, 45)]
    assert got == [8415032306198630212, -5236626025049887674, -5957839281333450960]


def test_spark_doc_kat_chain():
    """xxhash64('Spark', array(123), 2) = 5602566077635097486: the int steps of the chain use
    hashInt(v, seed=<previous hash>), which is what oracle_xxh64_int implements."""
    import struct

    xxhash = pytest.importorskip("xxhash")
    h = xxhash.xxh64(b"Spark", seed=42).intdigest()
    h = orc.lib().oracle_xxh64_int(123, h) & 0xFFFFFFFFFFFFFFFF
    h = orc.lib().oracle_xxh64_int(2, h)
    assert h == 5602566077635097486
    assert struct.pack("<i", 123) == b"\x7b\x00\x00\x00"


def test_hash_is_injective_on_a_window():
    # XXH64 on a 4-byte input is a composition of bijections on the zero-extended input, so no
    # ties inside a window; the (key, idx) tie-break never fires.  Spot-check 1e6 consecutive ints.
    k = orc.np_xxh64_int(np.arange(-500000, 500000))
    assert len(np.unique(k)) == len(k)


@pytest.mark.parametrize("size,s,c", [(0, 5, 42), (1, 0, 42), (7, 3, 84), (33, 2**31 - 5, 42), (1000, -7, 126)])
def test_perm_c_equals_numpy(size, s, c):
    a = orc.c_perm_full(size, s, c)
    b = orc.np_perm(size, s, c)
    assert np.array_equal(a, b)
    assert sorted(a.tolist()) == list(range(size))
    for f in (1, 3, 15, 64):
        assert np.array_equal(orc.c_perm_topk(size, s, c, f), b[:f])


def test_int32_wraparound_assumption():
    # Spark IntegerType '+' wraps (ANSI off): i + s + seed computed mod 2^32.
    s = 2**31 - 3
    p = orc.np_perm(10, s, 42)
    keys = orc.np_xxh64_int(np.array([orc._wrap32(i + s + 42) for i in range(1, 11)]))
    assert np.array_equal(p, np.lexsort((np.arange(10), keys)))
    assert np.array_equal(orc.c_perm_full(10, s, 42), p)


def test_khop_c_equals_python_random_graphs():
    for seed, directed in ((0, True), (1, False), (2, True)):
        n = 300
        src, dst = powerlaw_edges(n, 4000, seed) if seed != 1 else uniform_edges(n, 1500, seed)
        rowptr, col = orc.np_build_in_csr(src, dst, n, directed)
        roots = np.arange(n, dtype=np.int32)
        for fan in ([3, 3], [10, 5], [15, 10], [4], [2, 3, 2]):
            nb_c, ct_c = orc.c_sample_khop(rowptr, col, roots, fan)
            nb_p, ct_p = orc.np_sample_khop(rowptr, col, roots, fan)
            for a, b in zip(nb_c + ct_c, nb_p + ct_p):
                assert np.array_equal(a, b)


def test_khop_duplicate_edges_group_semantics():
    # directed multigraph: vertex 0 has in-edges 1,1,1 ; vertex 1 has in-edges 2,3.
    src = np.array([1, 1, 1, 2, 3])
    dst = np.array([0, 0, 0, 1, 1])
    rowptr, col = orc.np_build_in_csr(src, dst, 4, True)
    nbr, cnt = orc.c_sample_khop(rowptr, col, np.array([0], np.int32), [3, 4])
    assert nbr[0].tolist() == [1, 1, 1] and cnt[0].tolist() == [3]
    # GROUP BY (_0_hop,_1_hop): one group, array = sorted([2,3]*3) = [2,2,2,3,3,3], 4 taken
    assert cnt[1].tolist() == [4, 0, 0]
    got = sorted(nbr[1][:4].tolist())
    assert set(got) <= {2, 3} and len(got) == 4
    n2, c2 = orc.np_sample_khop(rowptr, col, np.array([0], np.int32), [3, 4])
    assert np.array_equal(n2[1], nbr[1]) and np.array_equal(c2[1], cnt[1])


def _graph16():
    g = load_golden("snc16_graph.json")
    e = np.array(g["edges"], dtype=np.int64)
    n = len(g["nodes"])
    return g, orc.np_build_in_csr(e[:, 0], e[:, 1], n, g["is_graph_directed"]), n


def test_reference_sgs_output_structural_rules_snc16():
    """The reference's real sampler output (non-deterministic shuffle) must satisfy every
    structural rule the oracle encodes; where all frontier degrees <= fanout it must be EQUAL."""
    g, (rowptr, col), n = _graph16()
    f = g["num_neighbors_to_sample"]
    out = load_golden("snc16_sgs_output.json")
    assert len(out["unlabeled"]) == n
    nbr, cnt = orc.c_sample_khop(rowptr, col, np.arange(n, dtype=np.int32), [f, f])
    edges_o = orc.tree_to_edges(np.arange(n), nbr, [f, f])
    deg = np.diff(rowptr)
    n_exact = 0
    for s in out["unlabeled"]:
        r = s["root_node"]["node_id"]
        ref_edges = sorted((e["src"], e["dst"]) for e in s["neighborhood"]["edges"])
        ref_nodes = sorted(x["node_id"] for x in s["neighborhood"]["nodes"])
        # rule: nodes == {root} U edge endpoints
        ends = {r} | {a for a, _ in ref_edges} | {b for _, b in ref_edges}
        assert sorted(ends) == ref_nodes
        # no self loops in this fixture => edges with dst == r are exactly hop 1, edges with
        # dst == k (k in hop1) are exactly S2(r, k)
        hop1 = [a for a, b in ref_edges if b == r]
        in_r = col[rowptr[r] : rowptr[r + 1]].tolist()
        assert set(hop1) <= set(in_r) and len(set(hop1)) == len(hop1)
        assert len(hop1) == min(f, deg[r]) == cnt[0][r]  # len == min(fanout, in-degree)
        assert {b for _, b in ref_edges} <= {r} | set(hop1)
        for k in hop1:
            s2 = [a for a, b in ref_edges if b == k]
            assert len(s2) == min(f, deg[k]) and set(s2) <= set(col[rowptr[k] : rowptr[k + 1]].tolist())
        if deg[r] <= f and all(deg[k] <= f for k in in_r):
            assert ref_edges == sorted(edges_o[r])
            n_exact += 1
    assert n_exact >= 3  # root 4 and the two isolated nodes at least
    r4 = next(s for s in out["unlabeled"] if s["root_node"]["node_id"] == 4)
    assert sorted((e["src"], e["dst"]) for e in r4["neighborhood"]["edges"]) == [(3, 9), (4, 9), (6, 9), (9, 4)]


def test_reference_sgs_output_labels_snc16():
    g, _, _ = _graph16()
    out = load_golden("snc16_sgs_output.json")
    labels = {x["node_id"]: x["node_label"] for x in g["nodes"]}
    for s in out["labeled"]:
        r = s["root_node"]["node_id"]
        assert [l["label"] for l in s["root_node_labels"]] == [labels[r]]
        assert s["root_node_labels"][0]["label_type"] == "node_label"


def test_reference_sgs_output_structural_rules_nablp16():
    """NABLP + random-negative RNN outputs of the reference (sampled from the 16-node graph)."""
    g, (rowptr, col), n = _graph16()
    f = 2  # this fixture was generated at fanout 2 (every root: min(2,deg) hop-1 edges, 2+2*2 total)
    out = load_golden("nablp16_sgs_output.json")
    deg = np.diff(rowptr)
    und = {(int(col[j]), v) for v in range(n) for j in range(rowptr[v], rowptr[v + 1])}
    assert len(out["rnn"]) == n
    for s in out["rnn"]:
        r = s["root_node"]["node_id"]
        edges = [(e["src"], e["dst"]) for e in s["neighborhood"]["edges"]]
        assert set(edges) <= und
        ids = {x["node_id"] for x in s["neighborhood"]["nodes"]}
        assert ids == {r} | {a for a, _ in edges} | {b for _, b in edges}
        assert len([1 for a, b in edges if b == r]) == min(f, deg[r])
    for s in out["nablp"]:
        r = s["root_node"]["node_id"]
        assert s["hard_neg_edges"] == [] and s["neg_edges"] == []  # NodeAnchorBasedLinkPredictionTask.scala:388-406
        assert 1 <= len(s["pos_edges"]) <= 2
        for pe in s["pos_edges"]:
            assert pe["src"] == r and (pe["src"], pe["dst"]) in und  # positives are out-edges of the root
        ids = {x["node_id"] for x in s["neighborhood"]["nodes"]}
        assert r in ids and all(pe["dst"] in ids for pe in s["pos_edges"])


def test_weighted_ops_oracle_semantics():
    """np_sample_op_weighted (the checker of gigl_sample_op_weighted_dev): ORDER BY weight DESC LIMIT k of
    NebulaQueryResponseTranslator.scala:90-105 on a hand graph, and the seeded RandomWeighted draw."""
    from oracle import oracle as O

    #          in-neighbours of 0: 1..6 with weights; of 1: one edge; 2: none
    src = np.array([1, 2, 3, 4, 5, 6, 0])
    dst = np.array([0, 0, 0, 0, 0, 0, 1])
    w_rec = np.array([0.5, 2.0, 2.0, np.nan, -1.0, 0.5, 9.0], dtype=np.float32)
    rowptr, col = O.np_build_in_csr(src, dst, 7, True)
    w = w_rec[np.lexsort((src, dst))]  # CSR order: by dst, then src
    roots = np.array([0, 1, 2], dtype=np.int32)
    nbr, cnt = O.np_sample_op_weighted((rowptr, col), w, roots, [4], [], 1, "top_k")
    assert cnt.tolist() == [4, 1, 0]
    assert nbr.reshape(3, 4)[0].tolist() == [2, 3, 1, 6]      # 2.0 (lower id first), 2.0, then the two 0.5s by id
    assert nbr.reshape(3, 4)[1].tolist() == [0, -1, -1, -1]
    nbr, cnt = O.np_sample_op_weighted((rowptr, col), w, roots, [8], [], 1, "top_k")
    assert nbr.reshape(3, 8)[0].tolist() == [2, 3, 1, 6, 5, 4, -1, -1]   # the NaN edge last
    # RandomWeighted: deterministic, a subset of the row, depends on the call number, zero-weight edges never beat positive ones
    a, _ = O.np_sample_op_weighted((rowptr, col), w, roots, [3], [], 1, "random_weighted")
    b, _ = O.np_sample_op_weighted((rowptr, col), w, roots, [3], [], 1, "random_weighted")
    assert np.array_equal(a, b) and set(a.reshape(3, 3)[0]) <= {1, 2, 3, 4, 5, 6}
    diff = [not np.array_equal(a, O.np_sample_op_weighted((rowptr, col), w, roots, [3], [], c, "random_weighted")[0]) for c in range(2, 12)]
    assert any(diff)
    w0 = np.where(np.isnan(w) | (w < 1), 0, w).astype(np.float32)
    z, _ = O.np_sample_op_weighted((rowptr, col), w0, roots, [2], [], 5, "random_weighted")
    assert set(z.reshape(3, 2)[0]) == {2, 3}


def test_sampling_op_config_methods():
    """The three sampling methods of a SamplingOp in the YAML / JSON form (the op names and fields of the reference's
    SamplingOpToNebulaQueryTranslatorTest.scala:44-118); userDefined is rejected like the reference's NotImplementedError."""
    from gigl_b200 import dag

    et = {"srcNodeType": "user", "relation": "to", "dstNodeType": "story"}
    ops = dag.ops_from_config({"samplingOps": [
        {"opName": "SamplingOpRandomUniform", "edgeType": et, "randomUniform": {"numNodesToSample": 10}},
        {"opName": "SamplingOpRandomWeighted", "edgeType": et, "randomWeighted": {"numNodesToSample": 10, "edgeFeatName": "edgeFeatName"}},
        {"opName": "SamplingOpTopK", "edgeType": et, "topK": {"numNodesToSample": 10, "edgeFeatName": "edgeFeatName"}}]})
    assert [(o.sampling_method, o.edge_feat_name, o.num_nodes_to_sample) for o in ops] == [
        ("random_uniform", "", 10), ("random_weighted", "edgeFeatName", 10), ("top_k", "edgeFeatName", 10)]
    for bad in ({"opName": "u", "edgeType": et, "userDefined": {"pathToUdf": "x"}},
                {"opName": "t", "edgeType": et, "topK": {"numNodesToSample": 3}},
                {"opName": "n", "edgeType": et}):
        with pytest.raises(ValueError):
            dag.ops_from_config({"samplingOps": [bad]})
