"""GPU: the argv-compatible sampler component end to end on the reference's own 16-node fixture
(scala/common/src/test/assets/subgraph_sampler/supervised_node_classification): TFRecords in, TFRecords out."""
import base64
import os

import numpy as np
import pytest
import yaml

from helpers import GOLDEN, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from gigl_b200 import Context

    c = Context(0)
    yield c
    c.close()


def _write_fixture(tmp):
    for name, sub in (("snc16_node_data.tfrecord.b64", "node_data"), ("snc16_edge_data.tfrecord.b64", "edge_data")):
        os.makedirs(tmp / sub, exist_ok=True)
        (tmp / sub / "data.tfrecord").write_bytes(base64.b64decode(open(os.path.join(GOLDEN, name)).read()))
    meta = {"condensedEdgeTypeToPreprocessedMetadata": {"0": {"dstNodeIdKey": "dst", "srcNodeIdKey": "src",
                                                               "mainEdgeInfo": {"tfrecordUriPrefix": "edge_data", "featureDim": 0}}},
            "condensedNodeTypeToPreprocessedMetadata": {"0": {"featureDim": 2, "featureKeys": ["f0", "f1"], "labelKeys": ["node_label"],
                                                               "nodeIdKey": "node_id", "tfrecordUriPrefix": "node_data"}}}
    cfg = {"graphMetadata": {"edgeTypes": [{"dstNodeType": "user", "relation": "friend", "srcNodeType": "user"}], "nodeTypes": ["user"]},
           "taskMetadata": {"nodeBasedTaskMetadata": {"supervisionNodeTypes": ["user"]}},
           "datasetConfig": {"subgraphSamplerConfig": {"numHops": 2, "numNeighborsToSample": 3, "numPositiveSamples": 2}},
           "sharedConfig": {"flattenedGraphMetadata": {"supervisedNodeClassificationOutput": {
               "labeledTfrecordUriPrefix": "output/labeled/samples/", "unlabeledTfrecordUriPrefix": "output/unlabeled/samples/"}},
               "preprocessedMetadataUri": "preprocessed_metadata.yaml"}}
    (tmp / "preprocessed_metadata.yaml").write_text(yaml.safe_dump(meta))
    (tmp / "frozen_gbml_config.yaml").write_text(yaml.safe_dump(cfg))


def test_sampler_component_on_reference_fixture(tmp_path):
    from gigl_b200 import sample_io as sio
    from gigl_b200 import subgraph_sampler
    from oracle import oracle as O

    _write_fixture(tmp_path)
    stats = subgraph_sampler.run("frozen_gbml_config.yaml", "test_job", None, root=str(tmp_path), log=lambda *_: None)
    assert stats["rnn"] == 16 and stats["snc"] == 14  # the reference wrote 16 RootedNodeNeighborhood + 14 labeled samples
    unl = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "output/unlabeled/samples/")))
    lab = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "output/labeled/samples/")))
    rnn = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(unl, verify=True))}
    snc = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(lab, verify=True))}
    assert sorted(rnn) == list(range(16))
    g = load_golden("snc16_graph.json")
    src, dst = np.array(g["edges"]).T
    rowptr, col = O.np_build_in_csr(src, dst, 16, False)
    roots = np.arange(16, dtype=np.int32)
    onbr, _ = O.c_sample_khop(rowptr, col, roots, [3, 3])
    want = O.tree_to_edges(roots, onbr, [3, 3])
    gold = {s["root_node"]["node_id"]: s for s in load_golden("snc16_sgs_output.json")["unlabeled"]}
    gold_l = {s["root_node"]["node_id"] for s in load_golden("snc16_sgs_output.json")["labeled"]}
    feats = {n["node_id"]: (n["f0"], n["f1"]) for n in g["nodes"]}
    for r in range(16):
        e = sorted((d["src_node_id"], d["dst_node_id"]) for d in rnn[r]["edges"])
        assert e == sorted(want[r])                                    # bit-exact index sets vs the oracle
        assert len(rnn[r]["nodes"]) == len(set(n["node_id"] for n in rnn[r]["nodes"]))
        for n in rnn[r]["nodes"]:
            assert np.allclose(n["feature_values"], feats[n["node_id"]], atol=1e-7)
        ge = sorted((d["src"], d["dst"]) for d in gold[r]["neighborhood"]["edges"])
        assert (len(e) == 0) == (len(ge) == 0)                          # isolated nodes: root only, no edges
        if len(ge) == 0:
            assert [n["node_id"] for n in rnn[r]["nodes"]] == [r]
    assert set(snc) == gold_l                                           # same roots carry training samples
    for r, s in snc.items():
        assert s["root_node_labels"][0]["label_type"] == "node_label"
        assert sorted((d["src_node_id"], d["dst_node_id"]) for d in s["edges"]) == sorted(want[r])


def _nablp_fixture(tmp, directed, n=40, e=220, with_edge_feat=True, strategy=None, num_pos=2):
    from helpers import powerlaw_edges, tf_example, tfrecord_bytes

    src, dst = powerlaw_edges(n, e, seed=21)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((n, 3)).astype(np.float32)
    ef = rng.standard_normal((e, 2)).astype(np.float32)
    os.makedirs(tmp / "node_data", exist_ok=True)
    os.makedirs(tmp / "edge_data", exist_ok=True)
    order = rng.permutation(n)
    (tmp / "node_data" / "data.tfrecord").write_bytes(tfrecord_bytes(
        [tf_example({"node_id": int(v), "emb": x[v, :2].tolist(), "score": float(x[v, 2])}) for v in order]))
    (tmp / "edge_data" / "data.tfrecord").write_bytes(tfrecord_bytes(
        [tf_example({"src": int(a), "dst": int(b), "w": ef[i].tolist()}) for i, (a, b) in enumerate(zip(src, dst))]))
    main = {"tfrecordUriPrefix": "edge_data"}
    if with_edge_feat:
        main.update({"featureKeys": ["w"], "featureDim": 2})
    meta = {"condensedEdgeTypeToPreprocessedMetadata": {"0": {"dstNodeIdKey": "dst", "srcNodeIdKey": "src", "mainEdgeInfo": main}},
            "condensedNodeTypeToPreprocessedMetadata": {"0": {"featureDim": 3, "featureKeys": ["emb", "score"], "nodeIdKey": "node_id",
                                                               "tfrecordUriPrefix": "node_data"}}}
    sgs = {"numPositiveSamples": num_pos}
    if strategy is None:
        sgs.update({"numHops": 2, "numNeighborsToSample": 3})
    else:
        sgs["subgraphSamplingStrategy"] = strategy
    et = {"dstNodeType": "user", "relation": "friend", "srcNodeType": "user"}
    cfg = {"graphMetadata": {"edgeTypes": [et], "nodeTypes": ["user"]},
           "taskMetadata": {"nodeAnchorBasedLinkPredictionTaskMetadata": {"supervisionEdgeTypes": [et]}},
           "datasetConfig": {"subgraphSamplerConfig": sgs},
           "sharedConfig": {"isGraphDirected": directed, "preprocessedMetadataUri": "preprocessed_metadata.yaml",
                            "flattenedGraphMetadata": {"nodeAnchorBasedLinkPredictionOutput": {
                                "tfrecordUriPrefix": "output/nablp/samples/",
                                "nodeTypeToRandomNegativeTfrecordUriPrefix": {"user": "output/random_negatives/user/"}}}}}
    (tmp / "preprocessed_metadata.yaml").write_text(yaml.safe_dump(meta))
    (tmp / "frozen_gbml_config.yaml").write_text(yaml.safe_dump(cfg))
    return src, dst, x, ef


def _dag_levels(path, roots, csr_by_op, keys):
    """The oracle's run of one messagePassingPaths entry (the config dict the component reads): op instance key -> padded
    tree, every op expanding the distinct result nodes of its inputs once (oracle.np_sample_dag)."""
    from gigl_b200 import dag
    from oracle import oracle as O

    planned = dag.plan(dag.ops_from_config(path), path["rootNodeType"])
    res = O.np_sample_dag(planned, lambda p: csr_by_op[p.op.op_name], roots)
    return [res[k][0] for k in keys]


def _canon(sample, key="edges"):
    return sorted((e["src_node_id"], e["dst_node_id"], tuple(np.float32(e["feature_values"]).tolist())) for e in sample[key])


@pytest.mark.parametrize("directed", [False, True])
def test_link_prediction_component_matches_restated_reference(tmp_path, directed):
    """NodeAnchorBasedLinkPredictionTask end to end (TFRecords in, RootedNodeNeighborhood + NodeAnchorBasedLinkPredictionSample
    TFRecords out, node and edge features hydrated) against the oracle's restatement of the reference SQL.  Small batches
    force positives whose trees live outside the anchor batch."""
    from gigl_b200 import sample_io as sio
    from gigl_b200 import subgraph_sampler
    from oracle import oracle as O

    n = 40
    src, dst, x, ef = _nablp_fixture(tmp_path, directed)
    stats = subgraph_sampler.run("frozen_gbml_config.yaml", "nablp_job", None, root=str(tmp_path), batch_roots=16, log=lambda *_: None)
    rowptr, col = O.np_build_in_csr(src, dst, n, directed)
    orow, ocol = O.np_build_in_csr(dst, src, n, directed)
    roots = np.arange(n, dtype=np.int32)
    nbr, _ = O.c_sample_khop(rowptr, col, roots, [3, 3])
    table = O.np_hydrated_edge_table(src, dst, directed, ef)
    want_rnn = O.np_assemble_rnn(roots, nbr, [3, 3], table)
    want = O.np_assemble_nablp(roots, nbr, [3, 3], table, O.np_sample_positives(orow, ocol, roots, 2))
    rnn_b = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "output/random_negatives/user/")))
    rnn = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(rnn_b, verify=True))}
    assert sorted(rnn) == list(range(n)) and stats["rnn"] == n
    for r in range(n):
        assert _canon(rnn[r]) == want_rnn[r][0] and sorted(v["node_id"] for v in rnn[r]["nodes"]) == want_rnn[r][1]
        for v in rnn[r]["nodes"]:
            assert np.array_equal(np.float32(v["feature_values"]), x[v["node_id"]])
    main_b = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "output/nablp/samples/")))
    got = {s["root_node"]["node_id"]: s for s in map(sio.parse_nablp_sample, sio.split_tfrecords(main_b, verify=True))}
    assert sorted(got) == sorted(want) and stats["nablp"] == len(want)
    for u, (we, wn, wp) in want.items():
        assert _canon(got[u]) == we and _canon(got[u], "pos_edges") == wp
        assert sorted(v["node_id"] for v in got[u]["nodes"]) == wn


def test_per_hop_fanouts_from_sampling_strategy(tmp_path):
    """`subgraphSamplingStrategy.messagePassingPaths` as a linear chain gives per-hop fanouts [4, 2]; `globalRandomUniform`
    gives [f] * numHops (subgraph_sampling_strategy.proto:38-79)."""
    from gigl_b200 import sample_io as sio
    from gigl_b200 import subgraph_sampler
    from oracle import oracle as O

    et = {"dstNodeType": "user", "relation": "friend", "srcNodeType": "user"}
    strat = {"messagePassingPaths": {"paths": [{"rootNodeType": "user", "samplingOps": [
        {"opName": "hop2", "edgeType": et, "inputOpNames": ["hop1"], "randomUniform": {"numNodesToSample": 2}},
        {"opName": "hop1", "edgeType": et, "randomUniform": {"numNodesToSample": 4}}]}]}}
    assert subgraph_sampler.fanouts_from_config({"subgraphSamplingStrategy": strat}) == [4, 2]
    assert subgraph_sampler.fanouts_from_config({"subgraphSamplingStrategy": {"globalRandomUniform": {
        "numHops": 3, "randomUniformSpec": {"numNodesToSample": 5}}}}) == [5, 5, 5]
    with pytest.raises(ValueError):
        subgraph_sampler.fanouts_from_config({"subgraphSamplingStrategy": {"messagePassingPaths": {"paths": [{"samplingOps": [
            {"opName": "a", "randomUniform": {"numNodesToSample": 2}}, {"opName": "b", "randomUniform": {"numNodesToSample": 2}}]}]}}})
    n = 40
    src, dst, x, ef = _nablp_fixture(tmp_path, False, with_edge_feat=False, strategy=strat, num_pos=1)
    stats = subgraph_sampler.run("frozen_gbml_config.yaml", "dag_job", None, root=str(tmp_path), log=lambda *_: None)
    assert stats["fanouts"] == [4, 2]
    rowptr, col = O.np_build_in_csr(src, dst, n, False)
    roots = np.arange(n, dtype=np.int32)
    nbr, _ = O.c_sample_khop(rowptr, col, roots, [4, 2])
    want = O.tree_to_edges(roots, nbr, [4, 2])
    rnn_b = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "output/random_negatives/user/")))
    rnn = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(rnn_b, verify=True))}
    for r in range(n):
        assert sorted((e["src_node_id"], e["dst_node_id"]) for e in rnn[r]["edges"]) == sorted(want[r])


def test_edge_rows_match_numpy(ctx):
    from helpers import powerlaw_edges
    from test_sample_assembly import np_edge_rows
    from gigl_b200 import Graph

    for directed in (False, True):
        n, e = 300, 5000
        src, dst = powerlaw_edges(n, e, seed=8)
        g = Graph.from_edges_host(ctx, n, src, dst, is_graph_directed=directed)
        rows = ctx.edge_rows_host(n, src, dst, directed)
        assert len(rows) == g.n_edges
        assert np.array_equal(rows, np_edge_rows(src, dst, n, directed))
    assert len(ctx.edge_rows_host(5, np.zeros(0, np.int32), np.zeros(0, np.int32), True)) == 0


def test_user_defined_labels_component_matches_restated_reference(tmp_path):
    """UserDefinedLabelsNodeAnchorBasedLinkPredictionTask end to end: positive / negative label edges in their own
    TFRecord tables (with label-edge features), sampled with permutation calls 3 / 4, hydrated against their tables."""
    from helpers import tf_example, tfrecord_bytes
    from gigl_b200 import sample_io as sio
    from gigl_b200 import subgraph_sampler
    from oracle import oracle as O

    n = 40
    src, dst, x, ef = _nablp_fixture(tmp_path, True)
    rng = np.random.default_rng(9)
    psrc, pdst = rng.integers(0, n, 50), rng.integers(0, n, 50)
    nsrc, ndst = rng.integers(0, n // 2, 30), rng.integers(0, n, 30)
    pef = rng.standard_normal((50, 1)).astype(np.float32)
    for sub, (a, b, f) in (("pos_edges", (psrc, pdst, pef)), ("neg_edges", (nsrc, ndst, None))):
        os.makedirs(tmp_path / sub, exist_ok=True)
        recs = [tf_example({"src": int(u), "dst": int(v), **({"lw": [float(f[i, 0])]} if f is not None else {})})
                for i, (u, v) in enumerate(zip(a, b))]
        (tmp_path / sub / "data.tfrecord").write_bytes(tfrecord_bytes(recs))
    meta = yaml.safe_load((tmp_path / "preprocessed_metadata.yaml").read_text())
    em = meta["condensedEdgeTypeToPreprocessedMetadata"]["0"]
    em["positiveEdgeInfo"] = {"tfrecordUriPrefix": "pos_edges", "featureKeys": ["lw"], "featureDim": 1}
    em["negativeEdgeInfo"] = {"tfrecordUriPrefix": "neg_edges"}
    (tmp_path / "preprocessed_metadata.yaml").write_text(yaml.safe_dump(meta))
    cfg = yaml.safe_load((tmp_path / "frozen_gbml_config.yaml").read_text())
    cfg["datasetConfig"]["subgraphSamplerConfig"].update({"numUserDefinedPositiveSamples": 2, "numUserDefinedNegativeSamples": 2})
    (tmp_path / "frozen_gbml_config.yaml").write_text(yaml.safe_dump(cfg))
    stats = subgraph_sampler.run("frozen_gbml_config.yaml", "udl_job", None, root=str(tmp_path), batch_roots=16, log=lambda *_: None)
    rowptr, col = O.np_build_in_csr(src, dst, n, True)
    roots = np.arange(n, dtype=np.int32)
    nbr, _ = O.c_sample_khop(rowptr, col, roots, [3, 3])
    p_out, n_out = O.np_build_in_csr(pdst, psrc, n, True), O.np_build_in_csr(ndst, nsrc, n, True)
    want = O.np_assemble_nablp(roots, nbr, [3, 3], O.np_hydrated_edge_table(src, dst, True, ef),
                               O.np_sample_positives(p_out[0], p_out[1], roots, 2, call_no=3),
                               pos_table=O.np_hydrated_edge_table(psrc, pdst, True, pef),
                               negatives=O.np_sample_positives(n_out[0], n_out[1], roots, 2, call_no=4),
                               neg_table=O.np_hydrated_edge_table(nsrc, ndst, True, None))
    main_b = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "output/nablp/samples/")))
    got = {s["root_node"]["node_id"]: s for s in map(sio.parse_nablp_sample, sio.split_tfrecords(main_b, verify=True))}
    assert sorted(got) == sorted(want) and stats["nablp"] == len(want) and stats["rnn"] == n
    for u, (we, wn, wp, wneg) in want.items():
        assert _canon(got[u]) == we and sorted(v["node_id"] for v in got[u]["nodes"]) == wn
        assert _canon(got[u], "pos_edges") == wp and _canon(got[u], "hard_neg_edges") == wneg


def test_typed_component_writes_rnn_per_node_type(tmp_path):
    """Two node types / two edge types with per-type SamplingOp DAGs (the shape of the reference's heterogeneous fixture,
    scala/common/src/test/assets/subgraph_sampler/heterogeneous): typed RootedNodeNeighborhood TFRecords per node type."""
    from helpers import tf_example, tfrecord_bytes
    from gigl_b200 import sample_io as sio
    from gigl_b200 import subgraph_sampler
    from oracle import oracle as O

    rng = np.random.default_rng(17)
    n_a, n_p = 30, 45  # authors, papers
    a2p = (rng.integers(0, n_a, 150), rng.integers(0, n_p, 150))
    p2a = (a2p[1].copy(), a2p[0].copy())
    xa = rng.standard_normal((n_a, 2)).astype(np.float32)
    xp = rng.standard_normal((n_p, 3)).astype(np.float32)
    for sub, recs in (("nodes_author", [tf_example({"node_id": int(i), "f": xa[i].tolist()}) for i in range(n_a)]),
                      ("nodes_paper", [tf_example({"node_id": int(i), "g": xp[i].tolist()}) for i in range(n_p)]),
                      ("edges_a2p", [tf_example({"src": int(u), "dst": int(v)}) for u, v in zip(*a2p)]),
                      ("edges_p2a", [tf_example({"src": int(u), "dst": int(v)}) for u, v in zip(*p2a)])):
        os.makedirs(tmp_path / sub, exist_ok=True)
        (tmp_path / sub / "data.tfrecord").write_bytes(tfrecord_bytes(recs))
    et0 = {"srcNodeType": "author", "relation": "author_to_paper", "dstNodeType": "paper"}
    et1 = {"srcNodeType": "paper", "relation": "paper_to_author", "dstNodeType": "author"}
    meta = {"condensedNodeTypeToPreprocessedMetadata": {
                "0": {"nodeIdKey": "node_id", "featureKeys": ["f"], "tfrecordUriPrefix": "nodes_author"},
                "1": {"nodeIdKey": "node_id", "featureKeys": ["g"], "tfrecordUriPrefix": "nodes_paper"}},
            "condensedEdgeTypeToPreprocessedMetadata": {
                "0": {"srcNodeIdKey": "src", "dstNodeIdKey": "dst", "mainEdgeInfo": {"tfrecordUriPrefix": "edges_a2p"}},
                "1": {"srcNodeIdKey": "src", "dstNodeIdKey": "dst", "mainEdgeInfo": {"tfrecordUriPrefix": "edges_p2a"}}}}
    paths = [{"rootNodeType": "paper", "samplingOps": [
                 {"opName": "writers", "edgeType": et0, "randomUniform": {"numNodesToSample": 3}},
                 {"opName": "their_papers", "edgeType": et1, "inputOpNames": ["writers"], "randomUniform": {"numNodesToSample": 2}}]},
             {"rootNodeType": "author", "samplingOps": [
                 {"opName": "papers", "edgeType": et1, "randomUniform": {"numNodesToSample": 2}},
                 {"opName": "written", "edgeType": et0, "randomUniform": {"numNodesToSample": 2}, "samplingDirection": "OUTGOING"}]}]
    cfg = {"graphMetadata": {"condensedEdgeTypeMap": {"0": et0, "1": et1}, "condensedNodeTypeMap": {"0": "author", "1": "paper"},
                             "edgeTypes": [et0, et1], "nodeTypes": ["author", "paper"]},
           "taskMetadata": {"nodeAnchorBasedLinkPredictionTaskMetadata": {"supervisionEdgeTypes": [et1]}},
           "datasetConfig": {"subgraphSamplerConfig": {"numPositiveSamples": 1, "subgraphSamplingStrategy": {"messagePassingPaths": {"paths": paths}}}},
           "sharedConfig": {"isGraphDirected": True, "preprocessedMetadataUri": "preprocessed_metadata.yaml",
                            "flattenedGraphMetadata": {"nodeAnchorBasedLinkPredictionOutput": {
                                "tfrecordUriPrefix": "out/nablp/",
                                "nodeTypeToRandomNegativeTfrecordUriPrefix": {"author": "out/rnn/author/", "paper": "out/rnn/paper/"}}}}}
    (tmp_path / "preprocessed_metadata.yaml").write_text(yaml.safe_dump(meta))
    (tmp_path / "frozen_gbml_config.yaml").write_text(yaml.safe_dump(cfg))
    stats = subgraph_sampler.run("frozen_gbml_config.yaml", "typed_job", None, root=str(tmp_path), batch_roots=20, log=lambda *_: None)
    assert stats["rnn_per_node_type"] == {"paper": n_p, "author": n_a}
    n = max(n_a, n_p)
    inc = lambda e: O.np_build_in_csr(e[0], e[1], n, True)  # noqa: E731
    outg = lambda e: O.np_build_in_csr(e[1], e[0], n, True)  # noqa: E731
    # paper roots: writers <- author_to_paper (call 1), their_papers <- paper_to_author (call 2)
    roots = np.arange(n_p, dtype=np.int32)
    ch = _dag_levels(paths[0], roots, {"writers": inc(a2p), "their_papers": inc(p2a)}, ["writers", "their_papers"])
    want = O.np_assemble_dag_rnn(roots, 1, [dict(parent=-1, fanout=3, condensed_edge_type=0, result_node_type=0, nbr=ch[0]),
                                            dict(parent=0, fanout=2, condensed_edge_type=1, result_node_type=1, nbr=ch[1])])
    raw = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "out/rnn/paper/")))
    got = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(raw, verify=True))}
    assert sorted(got) == list(range(n_p))
    for r in range(n_p):
        assert sorted((e["condensed_edge_type"], e["src_node_id"], e["dst_node_id"]) for e in got[r]["edges"]) == want[r][0]
        assert sorted((v["condensed_node_type"], v["node_id"]) for v in got[r]["nodes"]) == want[r][1]
        for v in got[r]["nodes"]:
            assert np.array_equal(np.float32(v["feature_values"]), (xa if v["condensed_node_type"] == 0 else xp)[v["node_id"]])
    # author roots: papers <- paper_to_author (call 1), written -> author_to_paper OUTGOING (call 2)
    roots = np.arange(n_a, dtype=np.int32)
    c1, _ = O.np_sample_chain([inc(p2a)], roots, [2], [1])
    c2, _ = O.np_sample_chain([outg(a2p)], roots, [2], [2])
    want = O.np_assemble_dag_rnn(roots, 0, [dict(parent=-1, fanout=2, condensed_edge_type=1, result_node_type=1, nbr=c1[0]),
                                            dict(parent=-1, fanout=2, condensed_edge_type=0, result_node_type=1, outgoing=True, nbr=c2[0])])
    raw = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "out/rnn/author/")))
    got = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(raw, verify=True))}
    for r in range(n_a):
        assert sorted((e["condensed_edge_type"], e["src_node_id"], e["dst_node_id"]) for e in got[r]["edges"]) == want[r][0]
        assert sorted((v["condensed_node_type"], v["node_id"]) for v in got[r]["nodes"]) == want[r][1]
    # main samples: supervision edge type paper -paper_to_author-> author; positives = OUTGOING sample over it from every paper
    # (call 3 = after the paper DAG's two ops), neighbourhood = the paper's merged with the author DAG of its positive
    roots = np.arange(n_p, dtype=np.int32)
    pos, _ = O.np_sample_chain([outg(p2a)], roots, [1], [3])
    paper_sets = O.np_assemble_dag_rnn(roots, 1, [dict(parent=-1, fanout=3, condensed_edge_type=0, result_node_type=0, nbr=ch[0]),
                                                  dict(parent=0, fanout=2, condensed_edge_type=1, result_node_type=1, nbr=ch[1])])
    want = O.np_assemble_typed_nablp(paper_sets, want, 0, {int(r): [int(pos[0][r])] for r in roots}, 1)
    raw = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "out/nablp/")))
    got = {s["root_node"]["node_id"]: s for s in map(sio.parse_nablp_sample, sio.split_tfrecords(raw, verify=True))}
    assert sorted(got) == sorted(want) and stats["nablp"] == len(want) and 0 < len(want) <= n_p
    for r, (wpe, we, wn) in want.items():
        assert got[r]["root_node"]["condensed_node_type"] == 1
        assert sorted((e["condensed_edge_type"], e["src_node_id"], e["dst_node_id"]) for e in got[r]["pos_edges"]) == [e[:3] for e in wpe]
        assert sorted((e["condensed_edge_type"], e["src_node_id"], e["dst_node_id"]) for e in got[r]["edges"]) == [e[:3] for e in we]
        assert sorted((v["condensed_node_type"], v["node_id"]) for v in got[r]["nodes"]) == wn


def test_typed_component_hydrates_edge_features_and_isolated_anchors(tmp_path):
    """user / item graph whose supervision edge type carries edge features and duplicate records: RootedNodeNeighborhoods
    carry one Edge per record, main samples one per key, pos_edges one per record; shouldIncludeIsolatedNodesInTraining
    keeps the anchors without a positive; an op with two input ops."""
    from helpers import tf_example, tfrecord_bytes
    from gigl_b200 import sample_io as sio
    from gigl_b200 import subgraph_sampler
    from oracle import oracle as O

    rng = np.random.default_rng(19)
    n_u, n_i = 36, 20
    n = max(n_u, n_i)
    follows = (rng.integers(0, n_u, 160), rng.integers(0, n_u, 160))
    clicks = (rng.integers(0, n_u // 2, 120), rng.integers(0, n_i, 120))  # the upper half of the users never clicks: isolated anchors
    clicks[0][90:] = clicks[0][:30]
    clicks[1][90:] = clicks[1][:30]
    cf = rng.standard_normal((120, 2)).astype(np.float32)
    xu = rng.standard_normal((n_u, 2)).astype(np.float32)
    for sub, recs in (("nodes_user", [tf_example({"node_id": int(i), "f": xu[i].tolist()}) for i in range(n_u)]),
                      ("nodes_item", [tf_example({"node_id": int(i)}) for i in range(n_i)]),
                      ("edges_follows", [tf_example({"src": int(u), "dst": int(v)}) for u, v in zip(*follows)]),
                      ("edges_clicks", [tf_example({"src": int(u), "dst": int(v), "w": cf[j].tolist()}) for j, (u, v) in enumerate(zip(*clicks))])):
        os.makedirs(tmp_path / sub, exist_ok=True)
        (tmp_path / sub / "data.tfrecord").write_bytes(tfrecord_bytes(recs))
    et0 = {"srcNodeType": "user", "relation": "follows", "dstNodeType": "user"}
    et1 = {"srcNodeType": "user", "relation": "clicks", "dstNodeType": "item"}
    meta = {"condensedNodeTypeToPreprocessedMetadata": {
                "0": {"nodeIdKey": "node_id", "featureKeys": ["f"], "tfrecordUriPrefix": "nodes_user"},
                "1": {"nodeIdKey": "node_id", "tfrecordUriPrefix": "nodes_item"}},
            "condensedEdgeTypeToPreprocessedMetadata": {
                "0": {"srcNodeIdKey": "src", "dstNodeIdKey": "dst", "mainEdgeInfo": {"tfrecordUriPrefix": "edges_follows"}},
                "1": {"srcNodeIdKey": "src", "dstNodeIdKey": "dst", "mainEdgeInfo": {"tfrecordUriPrefix": "edges_clicks", "featureKeys": ["w"]}}}}
    paths = [{"rootNodeType": "user", "samplingOps": [
                 {"opName": "liked", "edgeType": et1, "randomUniform": {"numNodesToSample": 2}, "samplingDirection": "OUTGOING"},
                 {"opName": "friends", "edgeType": et0, "randomUniform": {"numNodesToSample": 2}},
                 {"opName": "co_clickers", "edgeType": et1, "inputOpNames": ["liked"], "randomUniform": {"numNodesToSample": 2}},
                 {"opName": "fans", "edgeType": et0, "inputOpNames": ["friends", "co_clickers"], "randomUniform": {"numNodesToSample": 2}}]},
             {"rootNodeType": "item", "samplingOps": [
                 {"opName": "clickers", "edgeType": et1, "randomUniform": {"numNodesToSample": 3}}]}]
    cfg = {"graphMetadata": {"condensedEdgeTypeMap": {"0": et0, "1": et1}, "condensedNodeTypeMap": {"0": "user", "1": "item"},
                             "edgeTypes": [et0, et1], "nodeTypes": ["user", "item"]},
           "taskMetadata": {"nodeAnchorBasedLinkPredictionTaskMetadata": {"supervisionEdgeTypes": [et1]}},
           "datasetConfig": {"subgraphSamplerConfig": {"numPositiveSamples": 2, "subgraphSamplingStrategy": {"messagePassingPaths": {"paths": paths}}}},
           "sharedConfig": {"isGraphDirected": True, "preprocessedMetadataUri": "preprocessed_metadata.yaml",
                            "shouldIncludeIsolatedNodesInTraining": True,
                            "flattenedGraphMetadata": {"nodeAnchorBasedLinkPredictionOutput": {
                                "tfrecordUriPrefix": "out/nablp/",
                                "nodeTypeToRandomNegativeTfrecordUriPrefix": {"user": "out/rnn/user/", "item": "out/rnn/item/"}}}}}
    (tmp_path / "preprocessed_metadata.yaml").write_text(yaml.safe_dump(meta))
    (tmp_path / "frozen_gbml_config.yaml").write_text(yaml.safe_dump(cfg))
    stats = subgraph_sampler.run("frozen_gbml_config.yaml", "typed_ef_job", None, root=str(tmp_path), batch_roots=16, log=lambda *_: None)
    assert stats["rnn_per_node_type"] == {"user": n_u, "item": n_i} and stats["nablp"] == n_u
    inc = lambda e: O.np_build_in_csr(e[0], e[1], n, True)  # noqa: E731
    outg = lambda e: O.np_build_in_csr(e[1], e[0], n, True)  # noqa: E731
    records = O.np_typed_edge_records({0: (follows[0], follows[1], None), 1: (clicks[0], clicks[1], cf)})
    users = np.arange(n_u, dtype=np.int32)
    # every op expands the distinct result nodes of its inputs once (GraphDBSampler.scala:66-82); `fans` has two inputs
    lv = _dag_levels(paths[0], users, {"liked": outg(clicks), "friends": inc(follows), "co_clickers": inc(clicks), "fans": inc(follows)},
                     ["liked", "friends", "co_clickers", "fans@friends", "fans@co_clickers"])
    uops = [dict(parent=-1, fanout=2, condensed_edge_type=1, result_node_type=1, outgoing=True, nbr=lv[0]),
            dict(parent=-1, fanout=2, condensed_edge_type=0, result_node_type=0, nbr=lv[1]),
            dict(parent=0, fanout=2, condensed_edge_type=1, result_node_type=0, nbr=lv[2]),
            dict(parent=1, fanout=2, condensed_edge_type=0, result_node_type=0, nbr=lv[3]),
            dict(parent=2, fanout=2, condensed_edge_type=0, result_node_type=0, nbr=lv[4])]
    items = np.arange(n_i, dtype=np.int32)
    clk, _ = O.np_sample_chain([inc(clicks)], items, [3], [1])
    iops = [dict(parent=-1, fanout=3, condensed_edge_type=1, result_node_type=0, nbr=clk[0])]
    user_sets, item_sets = O.np_assemble_dag_rnn(users, 0, uops), O.np_assemble_dag_rnn(items, 1, iops)

    def canon(sample, key="edges"):
        return sorted(((e["condensed_edge_type"], e["src_node_id"], e["dst_node_id"], tuple(np.float32(e["feature_values"]).tolist()) or None)
                       for e in sample[key]), key=lambda e: (e[0], e[1], e[2], e[3] or ()))

    # RootedNodeNeighborhoods: the user DAG's root ops include `clicks` (features) -> edges hydrated, one per record
    want = O.np_hydrate_typed_rnn(user_sets, records)
    raw = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "out/rnn/user/")))
    got = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(raw, verify=True))}
    dup = 0
    for r in range(n_u):
        assert canon(got[r]) == want[r][0] and sorted((v["condensed_node_type"], v["node_id"]) for v in got[r]["nodes"]) == want[r][1]
        dup += len(want[r][0]) - len({e[:3] for e in want[r][0]})
    assert dup > 0
    # main samples: positives = 2 OUTGOING clicks per user, drawn as call 5 (after the user DAG's four ops)
    pos, pcnt = O.np_sample_chain([outg(clicks)], users, [2], [5])
    want = O.np_assemble_typed_nablp(user_sets, item_sets, 1, {int(r): pos[0][2 * r:2 * r + 2].tolist() for r in users}, 1, records,
                                     include_isolated=True)
    raw = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "out/nablp/")))
    got = {s["root_node"]["node_id"]: s for s in map(sio.parse_nablp_sample, sio.split_tfrecords(raw, verify=True))}
    assert sorted(got) == list(range(n_u)) and (pcnt[0] == 0).any()
    for r, (wpe, we, wn) in want.items():
        assert canon(got[r], "pos_edges") == wpe
        assert canon(got[r]) == sorted(we, key=lambda e: (e[0], e[1], e[2], e[3] or ()))
        assert sorted((v["condensed_node_type"], v["node_id"]) for v in got[r]["nodes"]) == wn


def test_typed_component_weighted_sampling_ops(tmp_path):
    """TopK / RandomWeighted ops in the component's config (subgraph_sampling_strategy.proto:17-58): the op orders the edges
    of its type by the SCALAR edge feature it names - the second of two feature keys here, so the column offset matters -
    and the written RootedNodeNeighborhoods are the oracle's (np_sample_op_weighted over the same CSR positions)."""
    from helpers import tf_example, tfrecord_bytes
    from gigl_b200 import dag
    from gigl_b200 import sample_io as sio
    from gigl_b200 import subgraph_sampler
    from oracle import oracle as O
    from test_sample_assembly import np_edge_rows

    rng = np.random.default_rng(23)
    n_u, n_i = 30, 18
    n = max(n_u, n_i)
    clicks = (rng.integers(0, n_u, 150), rng.integers(0, n_i, 150))
    cf = np.concatenate([rng.standard_normal((150, 2)), rng.integers(0, 4, (150, 1))], axis=1).astype(np.float32)   # [v0, v1 | w], w with ties
    for sub, recs in (("nodes_user", [tf_example({"node_id": int(i)}) for i in range(n_u)]),
                      ("nodes_item", [tf_example({"node_id": int(i)}) for i in range(n_i)]),
                      ("edges_clicks", [tf_example({"src": int(u), "dst": int(v), "v": cf[j, :2].tolist(), "w": float(cf[j, 2])})
                                        for j, (u, v) in enumerate(zip(*clicks))])):
        os.makedirs(tmp_path / sub, exist_ok=True)
        (tmp_path / sub / "data.tfrecord").write_bytes(tfrecord_bytes(recs))
    et1 = {"srcNodeType": "user", "relation": "clicks", "dstNodeType": "item"}
    meta = {"condensedNodeTypeToPreprocessedMetadata": {"0": {"nodeIdKey": "node_id", "tfrecordUriPrefix": "nodes_user"},
                                                        "1": {"nodeIdKey": "node_id", "tfrecordUriPrefix": "nodes_item"}},
            "condensedEdgeTypeToPreprocessedMetadata": {
                "0": {"srcNodeIdKey": "src", "dstNodeIdKey": "dst", "mainEdgeInfo": {"tfrecordUriPrefix": "edges_clicks", "featureKeys": ["v", "w"]}}}}
    paths = [{"rootNodeType": "user", "samplingOps": [
                 {"opName": "top_liked", "edgeType": et1, "topK": {"numNodesToSample": 2, "edgeFeatName": "w"}, "samplingDirection": "OUTGOING"},
                 {"opName": "lucky_clickers", "edgeType": et1, "inputOpNames": ["top_liked"], "randomWeighted": {"numNodesToSample": 2, "edgeFeatName": "w"}}]},
             {"rootNodeType": "item", "samplingOps": [
                 {"opName": "top_clickers", "edgeType": et1, "topK": {"numNodesToSample": 3, "edgeFeatName": "w"}}]}]
    cfg = {"graphMetadata": {"condensedEdgeTypeMap": {"0": et1}, "condensedNodeTypeMap": {"0": "user", "1": "item"},
                             "edgeTypes": [et1], "nodeTypes": ["user", "item"]},
           "taskMetadata": {"nodeAnchorBasedLinkPredictionTaskMetadata": {"supervisionEdgeTypes": [et1]}},
           "datasetConfig": {"subgraphSamplerConfig": {"numPositiveSamples": 1, "subgraphSamplingStrategy": {"messagePassingPaths": {"paths": paths}}}},
           "sharedConfig": {"isGraphDirected": True, "preprocessedMetadataUri": "preprocessed_metadata.yaml",
                            "shouldIncludeIsolatedNodesInTraining": True,
                            "flattenedGraphMetadata": {"nodeAnchorBasedLinkPredictionOutput": {
                                "tfrecordUriPrefix": "out/nablp/",
                                "nodeTypeToRandomNegativeTfrecordUriPrefix": {"user": "out/rnn/user/", "item": "out/rnn/item/"}}}}}
    (tmp_path / "preprocessed_metadata.yaml").write_text(yaml.safe_dump(meta))
    (tmp_path / "frozen_gbml_config.yaml").write_text(yaml.safe_dump(cfg))
    stats = subgraph_sampler.run("frozen_gbml_config.yaml", "weighted_job", None, root=str(tmp_path), batch_roots=16, log=lambda *_: None)
    assert stats["rnn_per_node_type"] == {"user": n_u, "item": n_i}
    inc, outg = O.np_build_in_csr(clicks[0], clicks[1], n, True), O.np_build_in_csr(clicks[1], clicks[0], n, True)
    w_in, w_out = cf[np_edge_rows(clicks[0], clicks[1], n, True), 2], cf[np_edge_rows(clicks[1], clicks[0], n, True), 2]
    records = O.np_typed_edge_records({0: (clicks[0], clicks[1], cf)})
    for rtype, path, n_roots, out_dir in (("user", paths[0], n_u, "out/rnn/user/"), ("item", paths[1], n_i, "out/rnn/item/")):
        planned = dag.plan(dag.ops_from_config(path), rtype)
        roots = np.arange(n_roots, dtype=np.int32)
        res = O.np_sample_dag(planned, lambda p: outg if p.op.sampling_direction == dag.OUTGOING else inc, roots,
                              weights_of=lambda p: w_out if p.op.sampling_direction == dag.OUTGOING else w_in)
        ops = [dict(parent=-1 if p.parent is None else [q.key for q in planned].index(p.parent), fanout=p.op.num_nodes_to_sample,
                    condensed_edge_type=0, result_node_type=1 if p.op.sampling_direction == dag.OUTGOING else 0,
                    outgoing=p.op.sampling_direction == dag.OUTGOING, nbr=res[p.key][0]) for p in planned]
        want = O.np_hydrate_typed_rnn(O.np_assemble_dag_rnn(roots, 0 if rtype == "user" else 1, ops), records)
        raw = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / out_dir)))
        got = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(raw, verify=True))}
        n_edges = 0
        for r in range(n_roots):
            have = sorted(((e["condensed_edge_type"], e["src_node_id"], e["dst_node_id"], tuple(np.float32(e["feature_values"]).tolist()) or None)
                           for e in got[r]["edges"]), key=lambda e: (e[0], e[1], e[2], e[3] or ()))
            assert have == want[r][0], (rtype, r)
            assert sorted((v["condensed_node_type"], v["node_id"]) for v in got[r]["nodes"]) == want[r][1]
            n_edges += len(have)
        assert n_edges > 0
    # the item DAG really is "the 3 heaviest clicks of every item": no dropped in-edge is heavier than a kept one
    rowptr, col = inc
    for it in range(n_i):
        row_w = np.sort(w_in[rowptr[it]:rowptr[it + 1]])[::-1]
        kept = sorted((e["feature_values"][2] for e in got[it]["edges"]), reverse=True)
        assert len(row_w) == 0 or kept[0] == row_w[0]


@pytest.mark.parametrize("directed", [False, True])
def test_component_sharded_over_ranks_writes_the_same_records(tmp_path, directed):
    """One process per GPU: rank r of `world` samples its contiguous share of the roots and writes part-r<rank>-* files.
    The union of the ranks' records is exactly the single-process output (the units are independent: no collective)."""
    import shutil

    from gigl_b200 import sample_io as sio
    from gigl_b200 import subgraph_sampler

    one, many = tmp_path / "one", tmp_path / "many"
    one.mkdir()
    _nablp_fixture(one, directed)
    shutil.copytree(one, many)
    s1 = subgraph_sampler.run("frozen_gbml_config.yaml", "j", None, root=str(one), batch_roots=16, log=lambda *_: None)
    world, tot = 3, {"rnn": 0, "nablp": 0}
    for rank in range(world):
        st = subgraph_sampler.run("frozen_gbml_config.yaml", "j", None, root=str(many), batch_roots=16, log=lambda *_: None,
                                  rank=rank, world=world, device=0)
        assert st["rank"] == rank and st["world"] == world
        for k in tot:
            tot[k] += st[k]
    assert tot["rnn"] == s1["rnn"] and tot["nablp"] == s1["nablp"] and s1["nablp"] > 0
    for sub in ("output/random_negatives/user/", "output/nablp/samples/"):
        files = sio.list_tfrecord_files(str(many / sub))
        assert {os.path.basename(f).split("-")[1] for f in files} == {"r000", "r001", "r002"}
        a = sorted(sio.split_tfrecords(b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(one / sub))), verify=True))
        b = sorted(sio.split_tfrecords(b"".join(open(f, "rb").read() for f in files), verify=True))
        assert a == b


def _hetero_reference_fixture(tmp):
    """The reference's heterogeneous sampler fixture (tests/golden/hetero_*: 15 authors, 19 papers, two featured edge types,
    its frozen config with messagePassingPaths) written out as the files its URIs name."""
    import base64
    import copy

    from helpers import load_golden

    g = load_golden("hetero_graph.json")
    cfg, meta = copy.deepcopy(g["frozen_gbml_config"]), copy.deepcopy(g["preprocessed_metadata"])
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for sub in ("nodes_author", "nodes_paper", "edges_author_to_paper", "edges_paper_to_author"):
        os.makedirs(tmp / sub, exist_ok=True)
        (tmp / sub / "data.tfrecord").write_bytes(base64.decodebytes(open(os.path.join(golden, f"hetero_{sub}.tfrecord.b64"), "rb").read()))
    meta["condensedNodeTypeToPreprocessedMetadata"]["0"]["tfrecordUriPrefix"] = "nodes_author/"
    meta["condensedNodeTypeToPreprocessedMetadata"]["1"]["tfrecordUriPrefix"] = "nodes_paper/"
    meta["condensedEdgeTypeToPreprocessedMetadata"]["0"]["mainEdgeInfo"]["tfrecordUriPrefix"] = "edges_author_to_paper/"
    meta["condensedEdgeTypeToPreprocessedMetadata"]["1"]["mainEdgeInfo"]["tfrecordUriPrefix"] = "edges_paper_to_author/"
    shared = cfg["sharedConfig"]
    shared["preprocessedMetadataUri"] = "preprocessed_metadata.yaml"
    outp = shared["flattenedGraphMetadata"]["nodeAnchorBasedLinkPredictionOutput"]
    outp["tfrecordUriPrefix"] = "out/nablp/"
    outp["nodeTypeToRandomNegativeTfrecordUriPrefix"] = {"author": "out/rnn/author/", "paper": "out/rnn/paper/"}
    (tmp / "preprocessed_metadata.yaml").write_text(yaml.safe_dump(meta))
    (tmp / "frozen_gbml_config.yaml").write_text(yaml.safe_dump(cfg))
    x = {t: np.array([[r["f0"], r["f1"]] for r in sorted(g[k], key=lambda r: r["node_id"])], dtype=np.float32)
         for t, k in ((0, "nodes_author"), (1, "nodes_paper"))}
    edges = {t: (np.array([r["src"] for r in g[k]]), np.array([r["dst"] for r in g[k]]), np.array([[r["f0"], r["f1"]] for r in g[k]], dtype=np.float32))
             for t, k in ((0, "edges_author_to_paper"), (1, "edges_paper_to_author"))}
    return x, edges


def test_typed_component_on_the_reference_heterogeneous_fixture(tmp_path):
    """The reference's own heterogeneous config (scala_spark35/.../subgraph_sampler/heterogeneous/node_anchor_based_link_
    prediction/frozen_gbml_config_graphdb_dblp_local.yaml) end to end: typed hydrated RootedNodeNeighborhoods for authors and
    papers, then the main samples of the author -to-> paper supervision edge type (numPositiveSamples 1,
    numMaxTrainingSamplesToOutput 10), against the oracle's restatement; every hop takes min(fanout, in-degree) neighbours."""
    from gigl_b200 import dag
    from gigl_b200 import sample_io as sio
    from gigl_b200 import subgraph_sampler
    from oracle import oracle as O

    x, edges = _hetero_reference_fixture(tmp_path)
    stats = subgraph_sampler.run("frozen_gbml_config.yaml", "hetero_ref", None, root=str(tmp_path), batch_roots=8, log=lambda *_: None)
    n_a, n_p = len(x[0]), len(x[1])
    assert (n_a, n_p) == (15, 19) and stats["rnn_per_node_type"] == {"author": n_a, "paper": n_p}
    n = max(n_a, n_p)
    inc = {t: O.np_build_in_csr(e[0], e[1], n, True) for t, e in edges.items()}
    outg = {t: O.np_build_in_csr(e[1], e[0], n, True) for t, e in edges.items()}
    records = O.np_typed_edge_records({t: e for t, e in edges.items()})

    def canon(sample, key="edges"):
        return sorted(((e["condensed_edge_type"], e["src_node_id"], e["dst_node_id"], tuple(np.float32(e["feature_values"]).tolist()) or None)
                       for e in sample[key]), key=lambda e: (e[0], e[1], e[2], e[3] or ()))

    sets = {}
    # author roots: op_4 = papers <- (paper to author), op_6 = authors <- (author to paper); paper roots: op_1, op_3 mirrored
    for name, rt, n_roots, t1, t2 in (("author", 0, n_a, 1, 0), ("paper", 1, n_p, 0, 1)):
        roots = np.arange(n_roots, dtype=np.int32)
        types = ["author", "paper"]
        two_ops = [dag.SamplingOp("o1", (types[1 - rt], "r1", types[rt]), 10), dag.SamplingOp("o2", (types[rt], "r2", types[1 - rt]), 10, ["o1"])]
        lv = O.np_sample_dag(dag.plan(two_ops, types[rt]), lambda p: inc[t1] if p.op.op_name == "o1" else inc[t2], roots)
        ch, cc = [lv["o1"][0], lv["o2"][0]], [lv["o1"][1], lv["o2"][1]]
        deg1 = np.diff(inc[t1][0])[roots]
        assert np.array_equal(cc[0], np.minimum(deg1, 10))  # every hop takes min(fanout, in-degree) neighbours
        ops = [dict(parent=-1, fanout=10, condensed_edge_type=t1, result_node_type=1 - rt, nbr=ch[0]),
               dict(parent=0, fanout=10, condensed_edge_type=t2, result_node_type=rt, nbr=ch[1])]
        sets[rt] = O.np_assemble_dag_rnn(roots, rt, ops)
        want = O.np_hydrate_typed_rnn(sets[rt], records)
        raw = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / f"out/rnn/{name}/")))
        got = {s["root_node"]["node_id"]: s for s in map(sio.parse_sample, sio.split_tfrecords(raw, verify=True))}
        assert sorted(got) == list(range(n_roots))
        for r in range(n_roots):
            assert got[r]["root_node"]["condensed_node_type"] == rt and np.array_equal(np.float32(got[r]["root_node"]["feature_values"]), x[rt][r])
            assert canon(got[r]) == want[r][0] and sorted((v["condensed_node_type"], v["node_id"]) for v in got[r]["nodes"]) == want[r][1]
            for v in got[r]["nodes"]:
                assert np.array_equal(np.float32(v["feature_values"]), x[v["condensed_node_type"]][v["node_id"]])
    # main samples: anchors = the first 10 authors (numMaxTrainingSamplesToOutput), one positive each over author -to-> paper,
    # drawn as call 3 (after the author path's two ops); anchors without an out-edge emit nothing
    authors = np.arange(10, dtype=np.int32)
    pos, _ = O.np_sample_chain([outg[0]], authors, [1], [3])
    want = O.np_assemble_typed_nablp({a: sets[0][a] for a in range(10)}, sets[1], 1, {int(a): [int(pos[0][a])] for a in authors}, 0, records)
    raw = b"".join(open(f, "rb").read() for f in sio.list_tfrecord_files(str(tmp_path / "out/nablp/")))
    got = {s["root_node"]["node_id"]: s for s in map(sio.parse_nablp_sample, sio.split_tfrecords(raw, verify=True))} if raw else {}
    assert sorted(got) == sorted(want) and stats["nablp"] == len(want) and len(want) > 0
    for a, (wpe, we, wn) in want.items():
        assert canon(got[a], "pos_edges") == wpe and len(wpe) == 1 and wpe[0][3] is not None  # the supervision edge carries its features
        assert canon(got[a]) == sorted(we, key=lambda e: (e[0], e[1], e[2], e[3] or ()))
        assert sorted((v["condensed_node_type"], v["node_id"]) for v in got[a]["nodes"]) == wn
