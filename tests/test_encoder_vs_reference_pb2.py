"""CPU: every sample type the host encoder writes, against the REFERENCE's own generated protobuf classes.
tests/golden/make_golden.py pushed the encoder's bytes for the cases of tests/encoder_cases.py through
snapchat.research.gbml.training_samples_schema_pb2 (in the build container, where the reference is mounted), checked
that the reference parser accepts them with nothing unknown and re-serialises them to the same size, and committed what
it saw.  Here: the encoder still writes exactly those bytes, and the repo's wire parser sees what the reference saw."""
import base64

import numpy as np
import pytest

import encoder_cases
from helpers import load_golden

from gigl_b200 import sample_io as sio


@pytest.fixture(scope="module")
def gold():
    return load_golden("encoder_pb2_views.json")


@pytest.fixture(scope="module")
def now():
    return encoder_cases.cases()


CASES = ["rnn_with_edge_features", "rnn_no_features_types_unset", "snc_labels", "nablp_main_edges", "nablp_user_defined_labels",
         "typed_rnn", "typed_nablp"]


def _f32(v):
    return np.float32(v).tolist()


def _node(d, mine):
    """reference view (HasField-aware dict) vs the repo parser's dict"""
    assert mine["node_id"] == d["node_id"] and _f32(mine["feature_values"]) == _f32(d["feature_values"])
    assert mine.get("condensed_node_type") == d.get("condensed_node_type")


def _edge(d, mine):
    assert (mine["src_node_id"], mine["dst_node_id"]) == (d["src"], d["dst"]) and _f32(mine["feature_values"]) == _f32(d["feature_values"])
    assert mine.get("condensed_edge_type") == d.get("condensed_edge_type")


@pytest.mark.parametrize("name", CASES)
def test_encoder_writes_the_bytes_the_reference_parser_was_shown(gold, now, name):
    kind, data = now[name]
    assert kind == gold[name]["kind"]
    assert data == base64.b64decode(gold[name]["bytes_b64"]), "the encoder's output changed: regenerate the golden with the reference mounted"


@pytest.mark.parametrize("name", CASES)
def test_repo_parser_sees_what_the_reference_parser_saw(gold, name):
    g = gold[name]
    recs = sio.split_tfrecords(base64.b64decode(g["bytes_b64"]), verify=True)
    assert len(recs) == len(g["records"]) > 0
    parse = sio.parse_nablp_sample if g["kind"] == "nablp" else sio.parse_sample
    for rec, ref in zip(recs, g["records"]):
        mine = parse(rec)
        _node(ref["root_node"], mine["root_node"])
        assert len(mine["nodes"]) == len(ref["neighborhood"]["nodes"]) and len(mine["edges"]) == len(ref["neighborhood"]["edges"])
        for a, b in zip(ref["neighborhood"]["nodes"], mine["nodes"]):
            _node(a, b)
        for a, b in zip(ref["neighborhood"]["edges"], mine["edges"]):
            _edge(a, b)
        if g["kind"] == "snc":
            assert mine["root_node_labels"] == ref["root_node_labels"]
        if g["kind"] == "nablp":
            for key in ("pos_edges", "hard_neg_edges", "neg_edges"):
                assert len(mine[key]) == len(ref[key])
                for a, b in zip(ref[key], mine[key]):
                    _edge(a, b)


def test_the_cases_cover_the_wire_features(gold):
    snc = gold["snc_labels"]["records"]
    labels = [r["root_node_labels"][0]["label"] for r in snc]
    assert min(labels) < 0 and 0 in labels and max(labels) > 0          # int32 labels: sign-extended, zero (field omitted), positive
    assert all(r["root_node_labels"][0]["label_type"] == "node_label" for r in snc)
    unset = gold["rnn_no_features_types_unset"]["records"]
    assert all("condensed_node_type" not in r["root_node"] for r in unset)  # optional field left unset
    assert any(r["root_node"]["node_id"] == 0 for r in unset)                # node id 0 = proto3 default, not on the wire
    udl = gold["nablp_user_defined_labels"]["records"]
    assert any(r["hard_neg_edges"] for r in udl) and all(r["pos_edges"] for r in udl) and not any(r["neg_edges"] for r in udl)
    typed = gold["typed_nablp"]["records"]
    assert {v["condensed_node_type"] for r in typed for v in r["neighborhood"]["nodes"]} == {0, 1}
    assert any(e["feature_values"] for r in typed for e in r["pos_edges"])   # the supervision edge type carries features
    assert any(not r["pos_edges"] for r in typed)                            # include_isolated: anchors without a positive
