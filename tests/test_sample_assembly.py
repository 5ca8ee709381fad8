"""CPU: hydration + sample assembly of the host encoder (gigl_encode_samples_ex_host) against the oracle's pure-python
restatement of the reference SQL (oracle.np_assemble_*): edge features, duplicate edge records of directed graphs,
NodeAnchorBasedLinkPredictionSample merging, and the reference sampler's own NABLP fixture output."""
import numpy as np
import pytest

from helpers import load_golden, powerlaw_edges

from gigl_b200 import sample_io as sio
from oracle import oracle as O


def np_edge_rows(src, dst, n, directed):
    """Feature row per slot of np_build_in_csr's CSR (numpy twin of gigl_edge_rows_host)."""
    src, dst = np.asarray(src, np.int64), np.asarray(dst, np.int64)
    idx = np.arange(len(src))
    if not directed:
        lo, hi = np.minimum(src, dst), np.maximum(src, dst)
        order = np.lexsort((idx, hi, lo))
        lo, hi, idx = lo[order], hi[order], idx[order]
        first = np.ones(len(lo), bool)
        first[1:] = (lo[1:] != lo[:-1]) | (hi[1:] != hi[:-1])
        lo, hi, idx = lo[first], hi[first], idx[first]
        loop = lo == hi
        src = np.concatenate([lo, hi[~loop]])
        dst = np.concatenate([hi, lo[~loop]])
        idx = np.concatenate([idx, idx[~loop]])
    order = np.lexsort((idx, src, dst))
    return idx[order].astype(np.int32)


def _feat(t, v):
    return tuple(np.asarray(v, dtype=np.float32).tolist())


def _decode_rnn(data):
    out = {}
    for rec in sio.split_tfrecords(data, verify=True):
        s = sio.parse_sample(rec)
        out[s["root_node"]["node_id"]] = s
    return out


def _edges(sample, key="edges"):
    return sorted((e["src_node_id"], e["dst_node_id"], _feat(None, e["feature_values"])) for e in sample[key])


@pytest.mark.parametrize("directed", [False, True])
def test_rnn_edge_hydration_matches_the_restated_join(directed):
    n, e = 60, 400
    src, dst = powerlaw_edges(n, e, seed=5)  # heavy tail => duplicate records, self loops
    rng = np.random.default_rng(1)
    ef = rng.standard_normal((e, 3)).astype(np.float32)
    x = rng.standard_normal((n, 4)).astype(np.float32)
    rowptr, col = O.np_build_in_csr(src, dst, n, directed)
    rows = np_edge_rows(src, dst, n, directed)
    assert len(rows) == len(col)
    roots = np.arange(n, dtype=np.int32)
    fan = [4, 3]
    nbr, _ = O.c_sample_khop(rowptr, col, roots, fan)
    table = O.np_hydrated_edge_table(src, dst, directed, ef)
    want = O.np_assemble_rnn(roots, nbr, fan, table)
    data, offs = sio.encode_samples(roots, fan, nbr, x, kind="rnn", csr=(rowptr, col), edge_rows=rows, edge_feat=ef)
    got = _decode_rnn(data)
    assert sorted(got) == list(range(n))
    n_dup = 0
    for r in range(n):
        we, wn = want[r]
        assert _edges(got[r]) == we
        assert sorted(v["node_id"] for v in got[r]["nodes"]) == wn
        for v in got[r]["nodes"]:
            assert np.array_equal(np.float32(v["feature_values"]), x[v["node_id"]])
        n_dup += len(we) - len(set((a, b) for a, b, _ in we))
    if directed:
        assert n_dup > 0  # the duplicate-record join really was exercised
    # without edge features the same join still multiplies duplicate records of a directed graph
    data2, _ = sio.encode_samples(roots, fan, nbr, x, kind="rnn", csr=(rowptr, col))
    got2 = _decode_rnn(data2)
    for r in range(n):
        assert [(a, b) for a, b, _ in _edges(got2[r])] == [(a, b) for a, b, _ in want[r][0]]


@pytest.mark.parametrize("directed", [False, True])
def test_nablp_assembly_matches_the_restated_sql(directed):
    n, e = 50, 260
    src, dst = powerlaw_edges(n, e, seed=11)
    rng = np.random.default_rng(2)
    ef = rng.standard_normal((e, 2)).astype(np.float32)
    ef[rng.integers(0, e, 40)] = 0.5  # some duplicate records with byte-identical features (array_distinct merges them)
    x = rng.standard_normal((n, 3)).astype(np.float32)
    rowptr, col = O.np_build_in_csr(src, dst, n, directed)
    out_rowptr, out_col = O.np_build_in_csr(dst, src, n, directed)
    rows = np_edge_rows(src, dst, n, directed)
    roots = np.arange(n, dtype=np.int32)
    fan = [3, 3]
    num_pos = 2
    nbr, _ = O.c_sample_khop(rowptr, col, roots, fan)
    positives = O.np_sample_positives(out_rowptr, out_col, roots, num_pos)
    table = O.np_hydrated_edge_table(src, dst, directed, ef)
    want = O.np_assemble_nablp(roots, nbr, fan, table, positives)
    pos = np.full((n, num_pos), -1, np.int32)
    for u, ps in positives.items():
        pos[u, : len(ps)] = ps
    pos_tree = np.where(pos >= 0, pos, -1).astype(np.int64)  # roots = arange(n): the tree of node p is tree p
    data, offs = sio.encode_samples(roots, fan, nbr, x, kind="nablp", csr=(rowptr, col), edge_rows=rows, edge_feat=ef,
                                    pos=pos, pos_tree=pos_tree)
    got = {}
    for rec in sio.split_tfrecords(data, verify=True):
        s = sio.parse_nablp_sample(rec)
        got[s["root_node"]["node_id"]] = s
    assert sorted(got) == sorted(want) and len(want) > 10
    src_only = 0
    for u, (we, wn, wp) in want.items():
        s = got[u]
        assert _edges(s) == we
        assert sorted(v["node_id"] for v in s["nodes"]) == wn
        assert _edges(s, "pos_edges") == wp
        assert s["hard_neg_edges"] == [] and s["neg_edges"] == []
        assert np.array_equal(np.float32(s["root_node"]["feature_values"]), x[u])
        src_only += rowptr[u + 1] == rowptr[u]
    if directed:
        assert src_only > 0  # source-only anchors (formNeighborhoodForSrcOnlyNodes) were exercised
    # anchors only: the positives' trees ride behind the anchors
    anchors = np.array(sorted(want)[:7], dtype=np.int32)
    extra = np.array(sorted({p for a in anchors for p in positives[int(a)]} - set(anchors.tolist())), dtype=np.int32)
    roots2 = np.concatenate([anchors, extra])
    nbr2, _ = O.c_sample_khop(rowptr, col, roots2, fan)
    where = {int(v): i for i, v in enumerate(roots2)}
    pos2 = pos[anchors]
    pt2 = np.array([[where[int(p)] if p >= 0 else -1 for p in row] for row in pos2], dtype=np.int64)
    data2, offs2 = sio.encode_samples(roots2, fan, nbr2, x, kind="nablp", csr=(rowptr, col), edge_rows=rows, edge_feat=ef,
                                      n_emit=len(anchors), pos=pos2, pos_tree=pt2)
    assert len(offs2) == len(anchors) + 1
    recs = sio.split_tfrecords(data2)
    assert len(recs) == len(anchors)
    for a, rec in zip(anchors, recs):
        s = sio.parse_nablp_sample(rec)
        assert _edges(s) == want[int(a)][0] and _edges(s, "pos_edges") == want[int(a)][2]


def test_nablp_structure_of_the_reference_fixture_output():
    """The reference sampler's own NodeAnchorBasedLinkPredictionSample output (16-node fixture, fanout 2, produced
    with the unseedable shuffle) obeys the merge rule the encoder implements: every positive's sampled 1-hop
    in-edges are inside the merged neighbourhood, edges are distinct, nodes = {root} U endpoints U positives."""
    g = load_golden("snc16_graph.json")
    src, dst = np.array(g["edges"]).T
    rowptr, col = O.np_build_in_csr(src, dst, 16, False)
    deg = np.diff(rowptr)
    out = load_golden("nablp16_sgs_output.json")
    for s in out["nablp"]:
        r = s["root_node"]["node_id"]
        edges = [(e["src"], e["dst"]) for e in s["neighborhood"]["edges"]]
        assert len(edges) == len(set(edges))  # array_distinct
        ids = sorted(v["node_id"] for v in s["neighborhood"]["nodes"])
        assert ids == sorted({r} | {a for a, _ in edges} | {b for _, b in edges} | {pe["dst"] for pe in s["pos_edges"]})
        for v in [r] + [pe["dst"] for pe in s["pos_edges"]]:
            assert len({a for a, b in edges if b == v}) >= min(2, deg[v])  # each merged tree brings its hop-1 edges
    # our encoder on the same graph produces samples for exactly the same anchors (every non-isolated node)
    roots = np.arange(16, dtype=np.int32)
    nbr, _ = O.c_sample_khop(rowptr, col, roots, [2, 2])
    positives = O.np_sample_positives(rowptr, col, roots, 2)  # undirected: the out-CSR is the in-CSR
    pos = np.full((16, 2), -1, np.int32)
    for u, ps in positives.items():
        pos[u, : len(ps)] = ps
    data, _ = sio.encode_samples(roots, [2, 2], nbr, None, kind="nablp", csr=(rowptr, col), pos=pos, pos_tree=pos.astype(np.int64))
    ours = {sio.parse_nablp_sample(r)["root_node"]["node_id"] for r in sio.split_tfrecords(data)}
    assert ours == {s["root_node"]["node_id"] for s in out["nablp"]}


@pytest.mark.parametrize("directed", [False, True])
def test_user_defined_labels_assembly_matches_the_restated_sql(directed):
    """UserDefinedLabelsNodeAnchorBasedLinkPredictionTask: positives / hard negatives come from their own (directed,
    duplicate-keeping, feature-carrying) edge tables; negatives are optional per anchor."""
    n, e = 45, 200
    src, dst = powerlaw_edges(n, e, seed=31)
    rng = np.random.default_rng(4)
    x = rng.standard_normal((n, 2)).astype(np.float32)
    ef = rng.standard_normal((e, 2)).astype(np.float32)
    psrc, pdst = rng.integers(0, n, 60), rng.integers(0, n, 60)
    pdst[:6], psrc[:6] = pdst[6:12], psrc[6:12]  # duplicate label records
    pef = rng.standard_normal((60, 3)).astype(np.float32)
    nsrc, ndst = rng.integers(0, n // 2, 40), rng.integers(0, n, 40)  # only some anchors have negatives
    nef = rng.standard_normal((40, 1)).astype(np.float32)
    rowptr, col = O.np_build_in_csr(src, dst, n, directed)
    roots = np.arange(n, dtype=np.int32)
    fan = [3, 2]
    nbr, _ = O.c_sample_khop(rowptr, col, roots, fan)
    num_pos, num_neg = 2, 3
    # label tables are never bidirectionalised (loadEdgeDataframeIntoSparkSql: only MAIN edges are, :262-273)
    p_out = O.np_build_in_csr(pdst, psrc, n, True)
    n_out = O.np_build_in_csr(ndst, nsrc, n, True)
    positives = O.np_sample_positives(p_out[0], p_out[1], roots, num_pos, call_no=3)
    negatives = O.np_sample_positives(n_out[0], n_out[1], roots, num_neg, call_no=4)
    want = O.np_assemble_nablp(roots, nbr, fan, O.np_hydrated_edge_table(src, dst, directed, ef), positives,
                               pos_table=O.np_hydrated_edge_table(psrc, pdst, True, pef), negatives=negatives,
                               neg_table=O.np_hydrated_edge_table(nsrc, ndst, True, nef))

    def dense(d, k):
        a = np.full((n, k), -1, np.int32)
        for u, vs in d.items():
            a[u, : len(vs)] = vs
        return a

    pos, neg = dense(positives, num_pos), dense(negatives, num_neg)
    main = sio.HostEdgeTable((rowptr, col), np_edge_rows(src, dst, n, directed), ef)
    ptab = sio.HostEdgeTable(O.np_build_in_csr(psrc, pdst, n, True), np_edge_rows(psrc, pdst, n, True), pef)
    ntab = sio.HostEdgeTable(O.np_build_in_csr(nsrc, ndst, n, True), np_edge_rows(nsrc, ndst, n, True), nef)
    data, offs = sio.encode_link_samples(roots, fan, nbr, x, n, pos, pos.astype(np.int64), main, ptab, neg, neg.astype(np.int64), ntab)
    got = {}
    for rec in sio.split_tfrecords(data, verify=True):
        s = sio.parse_nablp_sample(rec)
        got[s["root_node"]["node_id"]] = s
    assert sorted(got) == sorted(want) and len(want) > 10
    with_neg = 0
    for u, (we, wn, wp, wneg) in want.items():
        s = got[u]
        assert _edges(s) == we and sorted(v["node_id"] for v in s["nodes"]) == wn
        assert _edges(s, "pos_edges") == wp and _edges(s, "hard_neg_edges") == wneg
        assert all(len(pe["feature_values"]) == 3 for pe in s["pos_edges"]) and s["neg_edges"] == []
        with_neg += bool(wneg)
    assert 0 < with_neg < len(want)  # the LEFT JOIN was exercised both ways


def test_typed_rnn_from_a_sampling_op_dag_matches_the_restated_union():
    """Heterogeneous RootedNodeNeighborhoods: union of the ops' typed edge / node sets + the root, nodes hydrated with
    their own type's features (GraphDBSampler.scala:129-148, SGSTask.hydrateRnn)."""
    rng = np.random.default_rng(21)
    n_user, n_item, n = 40, 25, 40
    follows = (rng.integers(0, n_user, 250), rng.integers(0, n_user, 250))     # user -> user, condensed edge type 0
    shown = (rng.integers(0, n_item, 200), rng.integers(0, n_user, 200))       # item -> user, 1
    clicks = (rng.integers(0, n_user, 220), rng.integers(0, n_item, 220))      # user -> item, 2
    inc = lambda e: O.np_build_in_csr(e[0], e[1], n, True)  # noqa: E731
    outg = lambda e: O.np_build_in_csr(e[1], e[0], n, True)  # noqa: E731
    roots = np.arange(n_user, dtype=np.int32)
    # ops: friends (users <- follows), seen (items shown to the user), clickers (users who clicked a seen item),
    #      their_clicks (items the friends clicked, OUTGOING over clicks)
    friends, _ = O.np_sample_chain([inc(follows)], roots, [3], [1])
    seen, _ = O.np_sample_chain([inc(shown)], roots, [2], [2])
    clickers, _ = O.np_sample_chain([inc(shown), inc(clicks)], roots, [2, 2], [2, 3])
    their, _ = O.np_sample_chain([inc(follows), outg(clicks)], roots, [3, 2], [1, 4])
    ops = [dict(parent=-1, fanout=3, condensed_edge_type=0, result_node_type=0, nbr=friends[0]),
           dict(parent=-1, fanout=2, condensed_edge_type=1, result_node_type=1, nbr=seen[0]),
           dict(parent=1, fanout=2, condensed_edge_type=2, result_node_type=0, nbr=clickers[1]),
           dict(parent=0, fanout=2, condensed_edge_type=2, result_node_type=1, outgoing=True, nbr=their[1])]
    xu = rng.standard_normal((n_user, 3)).astype(np.float32)
    xi = rng.standard_normal((n_item, 5)).astype(np.float32)
    want = O.np_assemble_dag_rnn(roots, 0, ops)
    data, offs = sio.encode_dag_samples(roots, 0, ops, [xu, xi])
    recs = sio.split_tfrecords(data, verify=True)
    assert len(recs) == n_user and offs[-1] == len(data)
    n_out_edges = 0
    for r, rec in zip(roots, recs):
        s = sio.parse_sample(rec)
        assert s["root_node"]["node_id"] == r and s["root_node"]["condensed_node_type"] == 0
        assert np.array_equal(np.float32(s["root_node"]["feature_values"]), xu[r])
        we, wn = want[int(r)]
        assert sorted((e["condensed_edge_type"], e["src_node_id"], e["dst_node_id"]) for e in s["edges"]) == we
        assert sorted((v["condensed_node_type"], v["node_id"]) for v in s["nodes"]) == wn
        for v in s["nodes"]:
            tab = xu if v["condensed_node_type"] == 0 else xi
            assert np.array_equal(np.float32(v["feature_values"]), tab[v["node_id"]])
        # an OUTGOING op's edges point away from the frontier: user (friend) -> item over `clicks`
        n_out_edges += sum(1 for t, a, b in we if t == 2 and (0, a) in wn and (1, b) in wn)
    assert n_out_edges > 0


def _typed_fixture(seed=33):
    """user / item graph with three edge types; `clicks` carries 2 features and duplicate records, `follows` none."""
    rng = np.random.default_rng(seed)
    n_user, n_item, n = 40, 25, 40
    et = {0: (rng.integers(0, n_user, 250), rng.integers(0, n_user, 250), None),                                   # user -follows-> user
          1: (rng.integers(0, n_item, 200), rng.integers(0, n_user, 200), rng.standard_normal((200, 1)).astype(np.float32)),  # item -shown_to-> user
          2: (rng.integers(0, n_user, 260), rng.integers(0, n_item, 260), rng.standard_normal((260, 2)).astype(np.float32))}  # user -clicks-> item
    et[2][0][200:] = et[2][0][:60]  # duplicate (src, dst) records with their own feature rows
    et[2][1][200:] = et[2][1][:60]
    inc = {t: O.np_build_in_csr(e[0], e[1], n, True) for t, e in et.items()}
    outg = {t: O.np_build_in_csr(e[1], e[0], n, True) for t, e in et.items()}
    tabs = [sio.HostEdgeTable(inc[t], np_edge_rows(et[t][0], et[t][1], n, True), et[t][2]) for t in range(3)]
    xu = rng.standard_normal((n, 3)).astype(np.float32)
    xi = rng.standard_normal((n, 5)).astype(np.float32)
    return n_user, n_item, n, et, inc, outg, tabs, xu, xi


def _typed_edges(sample, key="edges"):
    return sorted(((e["condensed_edge_type"], e["src_node_id"], e["dst_node_id"], _feat(None, e["feature_values"]) or None)
                   for e in sample[key]), key=lambda e: (e[0], e[1], e[2], e[3] or ()))


def _user_dag(inc, outg, roots):
    """friends (1), seen (2), clickers <- seen (3), mixed <- {friends, clickers} over follows (4, two instances),
    their_clicks <- friends OUTGOING over clicks (5)."""
    friends, _ = O.np_sample_chain([inc[0]], roots, [3], [1])
    seen, _ = O.np_sample_chain([inc[1]], roots, [2], [2])
    clickers, _ = O.np_sample_chain([inc[1], inc[2]], roots, [2, 2], [2, 3])
    mixed_a, _ = O.np_sample_chain([inc[0], inc[0]], roots, [3, 2], [1, 4])
    mixed_b, _ = O.np_sample_chain([inc[1], inc[2], inc[0]], roots, [2, 2, 2], [2, 3, 4])
    their, _ = O.np_sample_chain([inc[0], outg[2]], roots, [3, 2], [1, 5])
    return [dict(parent=-1, fanout=3, condensed_edge_type=0, result_node_type=0, nbr=friends[0]),
            dict(parent=-1, fanout=2, condensed_edge_type=1, result_node_type=1, nbr=seen[0]),
            dict(parent=1, fanout=2, condensed_edge_type=2, result_node_type=0, nbr=clickers[1]),
            dict(parent=0, fanout=2, condensed_edge_type=0, result_node_type=0, nbr=mixed_a[1]),
            dict(parent=2, fanout=2, condensed_edge_type=0, result_node_type=0, nbr=mixed_b[2]),
            dict(parent=0, fanout=2, condensed_edge_type=2, result_node_type=1, outgoing=True, nbr=their[1])]


def test_typed_rnn_edge_hydration_and_multi_input_ops():
    """Typed RootedNodeNeighborhoods with the LEFT JOIN against per-type edge records (SGSTask.scala:243-262): one Edge per
    matching record, features of the record; an op with two input ops contributes the union of its two instances."""
    n_user, n_item, n, et, inc, outg, tabs, xu, xi = _typed_fixture()
    roots = np.arange(n_user, dtype=np.int32)
    ops = _user_dag(inc, outg, roots)
    records = O.np_typed_edge_records(et)
    want = O.np_hydrate_typed_rnn(O.np_assemble_dag_rnn(roots, 0, ops), records)
    data, offs = sio.encode_typed_samples(roots, 0, ops, [xu, xi], tabs, kind="rnn")
    recs = sio.split_tfrecords(data, verify=True)
    assert len(recs) == n_user and offs[-1] == len(data)
    multi = dup = 0
    for r, rec in zip(roots, recs):
        s = sio.parse_sample(rec)
        we, wn = want[int(r)]
        assert _typed_edges(s) == we
        assert sorted((v["condensed_node_type"], v["node_id"]) for v in s["nodes"]) == wn
        keys = [e[:3] for e in we]
        dup += len(keys) - len(set(keys))
        a = {int(c) for c in ops[3]["nbr"][r * 6:(r + 1) * 6] if c >= 0}
        b = {int(c) for c in ops[4]["nbr"][r * 8:(r + 1) * 8] if c >= 0}
        multi += bool(a - b) and bool(b - a)
        assert {(0, v) for v in a | b} <= set(wn)
    assert dup > 0 and multi > 0  # duplicate records were joined; both instances of the two-input op contributed
    # hydration switched off: the same keys, no features, one Edge per key
    data0, _ = sio.encode_typed_samples(roots, 0, ops, [xu, xi], tabs, kind="rnn", hydrate_edges=False)
    plain, _ = sio.encode_dag_samples(roots, 0, ops, [xu, xi])
    assert data0 == plain


@pytest.mark.parametrize("include_isolated", [False, True])
def test_typed_nablp_assembly_matches_the_restated_merge(include_isolated):
    """The typed task's main samples: supervision edge type user -clicks-> item; positives = OUTGOING sample over clicks;
    neighbourhood = anchor's DAG merged by key with the item DAG of every positive."""
    n_user, n_item, n, et, inc, outg, tabs, xu, xi = _typed_fixture(seed=34)
    roots = np.arange(n_user, dtype=np.int32)
    ops = _user_dag(inc, outg, roots)
    num_pos = 2
    pos, pcnt = O.np_sample_chain([outg[2]], roots, [num_pos], [len(ops) + 1])
    pos = pos[0].reshape(n_user, num_pos)
    assert (pcnt[0] == 0).any() and (pcnt[0] > 0).any()
    # item DAG: clickers of the item (1), then what those users were shown (2)
    items = np.unique(pos[pos >= 0]).astype(np.int32)
    items = items[items % 5 != 0]  # some positives have no tree in this call: they contribute the node alone
    t1, _ = O.np_sample_chain([inc[2]], items, [2], [1])
    t2, _ = O.np_sample_chain([inc[2], inc[1]], items, [2, 2], [1, 2])
    tops = [dict(parent=-1, fanout=2, condensed_edge_type=2, result_node_type=0, nbr=t1[0]),
            dict(parent=0, fanout=2, condensed_edge_type=1, result_node_type=1, nbr=t2[1])]
    tree = np.where(pos >= 0, np.searchsorted(items, np.maximum(pos, 0)), -1)
    tree = np.where((tree >= 0) & (tree < len(items)) & (items[np.minimum(tree, len(items) - 1)] == pos), tree, -1)
    records = O.np_typed_edge_records(et)
    want = O.np_assemble_typed_nablp(O.np_assemble_dag_rnn(roots, 0, ops), O.np_assemble_dag_rnn(items, 1, tops), 1,
                                     {int(r): pos[i].tolist() for i, r in enumerate(roots)}, 2, records, include_isolated=include_isolated)
    data, offs = sio.encode_typed_samples(roots, 0, ops, [xu, xi], tabs, kind="nablp", pos=pos, pos_tree=tree, pos_condensed_edge_type=2,
                                          target_roots=items, target_node_type=1, target_ops=tops, include_isolated=include_isolated)
    got = {}
    for rec in sio.split_tfrecords(data, verify=True):
        s = sio.parse_nablp_sample(rec)
        got[s["root_node"]["node_id"]] = s
    assert sorted(got) == sorted(want)
    assert (len(got) == n_user) == include_isolated
    assert int((np.diff(offs) > 0).sum()) == len(got)
    for a, (wpe, we, wn) in want.items():
        s = got[a]
        assert np.array_equal(np.float32(s["root_node"]["feature_values"]), xu[a]) and s["root_node"]["condensed_node_type"] == 0
        assert _typed_edges(s, "pos_edges") == wpe
        assert _typed_edges(s) == sorted(we, key=lambda e: (e[0], e[1], e[2], e[3] or ()))
        assert sorted((v["condensed_node_type"], v["node_id"]) for v in s["nodes"]) == wn
        assert not s["hard_neg_edges"] and not s["neg_edges"]
        for v in s["nodes"]:
            assert np.array_equal(np.float32(v["feature_values"]), (xu if v["condensed_node_type"] == 0 else xi)[v["node_id"]])
