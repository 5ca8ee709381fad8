"""GPU: the argv-compatible sampler component end to end on the reference's own 16-node fixture
(scala/common/src/test/assets/subgraph_sampler/supervised_node_classification): TFRecords in, TFRecords out."""
import base64
import os

import numpy as np
import pytest
import yaml

from helpers import GOLDEN, load_golden

pytestmark = pytest.mark.gpu


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
