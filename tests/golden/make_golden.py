#!/usr/bin/env python
"""Regenerate tests/golden/*.json from the reference's own fixtures.

Runs ONLY in the build container (needs /root/reference, which does not exist on
the GPU box).  The JSON files it writes are committed; tests read those, never
/root/reference.

What is captured
----------------
* ``snc16_graph.json``  - the 16-node / 34-edge input graph of
  scala/common/src/test/assets/subgraph_sampler/supervised_node_classification
  ({node,edge}_data/data.tfrecord), decoded from tf.Example records.
* ``snc16_sgs_output.json`` - the reference sampler's real output for that graph
  (scala/common/src/test/assets/split_generator/supervised_node_classification/
  sgs_output/{unlabeled,labeled}/samples/data.tfrecord), decoded with the reference's
  generated training_samples_schema_pb2 (RootedNodeNeighborhood /
  SupervisedNodeClassificationSample).
* ``nablp27_graph.json`` - the 27-node toy graph (+ user-defined pos/neg edges) of
  scala/common/src/test/assets/subgraph_sampler/node_anchor_based_link_prediction.
* ``nablp16_sgs_output.json`` - the reference's NABLP + random-negative RNN outputs under
  split_generator/node_anchor_based_link_prediction/sgs_output; these were sampled from the
  16-node graph above (16 RNN, 14 NABLP samples), not from the 27-node one.
* ``hetero_graph.json`` / ``hetero_sgs_output.json`` / ``hetero_*.tfrecord.b64`` - the reference's heterogeneous fixture
  (scala_spark35/common/src/test/assets/subgraph_sampler/heterogeneous/node_anchor_based_link_prediction: 15 authors,
  19 papers, two featured edge types, the frozen config with ``messagePassingPaths``) and the typed, hydrated
  RootedNodeNeighborhood outputs the reference holds for it
  (.../split_generator/hetero_node_anchor_based_link_prediction/sgs_output/random_negative_rooted_neighborhood_samples).
  ``python make_golden.py hetero`` regenerates only these.
* ``encoder_pb2_views.json`` - the bytes the repo's host encoder writes for every sample type (tests/encoder_cases.py:
  RootedNodeNeighborhood with / without features, SupervisedNodeClassificationSample with negative / zero labels,
  NodeAnchorBasedLinkPredictionSample with main-edge and user-defined positives / hard negatives, the typed variants)
  as the REFERENCE's generated protobuf classes parse them (``python make_golden.py encoder`` regenerates it).
* ``xxh64_kat.json`` - known answers for Spark's ``xxhash64`` (XXH64.hashInt, seed 42)
  computed with the independent python ``xxhash`` package, incl. Spark's documented
  ``xxhash64('Spark', array(123), 2) = 5602566077635097486`` chain check.
* ``*.tfrecord.b64`` - raw reference TFRecord bytes (tiny) to pin the framing/crc code.
"""
import base64
import json
import os
import struct
import sys

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "python"))
HERE = os.path.dirname(os.path.abspath(__file__))
ASSETS = os.path.join(REF, "scala/common/src/test/assets")


def read_tfrecords(path):
    out = []
    with open(path, "rb") as f:
        buf = f.read()
    off = 0
    while off < len(buf):
        (n,) = struct.unpack_from("<Q", buf, off)
        off += 12
        out.append(buf[off : off + n])
        off += n + 4
    return out


def decode_example(rec):
    """Minimal tf.Example decoder -> {key: list}.  (tensorflow is not installed.)"""

    def varint(b, i):
        r = 0
        s = 0
        while True:
            c = b[i]
            i += 1
            r |= (c & 0x7F) << s
            s += 7
            if c < 0x80:
                return r, i

    def fields(b):
        i = 0
        while i < len(b):
            tag, i = varint(b, i)
            fn, wt = tag >> 3, tag & 7
            if wt == 0:
                v, i = varint(b, i)
            elif wt == 2:
                ln, i = varint(b, i)
                v = b[i : i + ln]
                i += ln
            elif wt == 5:
                v = b[i : i + 4]
                i += 4
            elif wt == 1:
                v = b[i : i + 8]
                i += 8
            else:
                raise ValueError(wt)
            yield fn, wt, v

    res = {}
    for fn, _, features in fields(rec):  # Example.features = 1
        for fn2, _, entry in fields(features):  # Features.feature (map) = 1
            key = None
            val = None
            for fn3, _, v in fields(entry):
                if fn3 == 1:
                    key = v.decode()
                else:
                    val = v
            lst = []
            for kind, _, payload in fields(val or b""):  # 1 bytes_list, 2 float_list, 3 int64_list
                for fn5, wt5, v in fields(payload):
                    if kind == 3:
                        if wt5 == 0:
                            lst.append(v if v < (1 << 63) else v - (1 << 64))
                        else:  # packed
                            j = 0
                            while j < len(v):
                                x, j = varint(v, j)
                                lst.append(x if x < (1 << 63) else x - (1 << 64))
                    elif kind == 2:
                        if wt5 == 5:
                            lst.append(struct.unpack("<f", v)[0])
                        else:
                            lst.extend(struct.unpack("<%df" % (len(v) // 4), v))
                    else:
                        lst.append(v.decode("latin1"))
            res[key] = lst
    return res


def node_to_dict(n):
    d = {"node_id": n.node_id, "feature_values": list(n.feature_values)}
    if n.HasField("condensed_node_type"):
        d["condensed_node_type"] = n.condensed_node_type
    return d


def edge_to_dict(e):
    d = {"src": e.src_node_id, "dst": e.dst_node_id, "feature_values": list(e.feature_values)}
    if e.HasField("condensed_edge_type"):
        d["condensed_edge_type"] = e.condensed_edge_type
    return d


def graph_to_dict(g):
    return {"nodes": [node_to_dict(n) for n in g.nodes], "edges": [edge_to_dict(e) for e in g.edges]}


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", name)


def hetero():
    """The heterogeneous fixture of the spark35 sampler and the reference's typed outputs for it."""
    import glob

    from snapchat.research.gbml import training_samples_schema_pb2 as ts

    assets35 = os.path.join(REF, "scala_spark35/common/src/test/assets")
    base = os.path.join(assets35, "subgraph_sampler/heterogeneous/node_anchor_based_link_prediction")
    tables = {"nodes_author": "node_features_dir/user/features", "nodes_paper": "node_features_dir/story/features",
              "edges_author_to_paper": "edge_features_dir/user-to-story/main_edges/features",
              "edges_paper_to_author": "edge_features_dir/story-to-user/main_edges/features"}
    graph = {"source": "scala_spark35/common/src/test/assets/subgraph_sampler/heterogeneous/node_anchor_based_link_prediction",
             "condensed_node_types": {"0": "author", "1": "paper"},
             "condensed_edge_types": {"0": ["author", "to", "paper"], "1": ["paper", "to", "author"]}}
    for name, sub in tables.items():
        files = sorted(glob.glob(os.path.join(base, sub, "*.tfrecord")))
        recs = [decode_example(r) for f in files for r in read_tfrecords(f)]
        graph[name] = [{k: v[0] for k, v in r.items()} for r in recs]
        with open(os.path.join(HERE, "hetero_" + name + ".tfrecord.b64"), "w") as g:
            g.write(base64.encodebytes(b"".join(open(f, "rb").read() for f in files)).decode())
        print("wrote", "hetero_" + name + ".tfrecord.b64")
    import yaml

    graph["frozen_gbml_config"] = yaml.safe_load(open(os.path.join(base, "frozen_gbml_config_graphdb_dblp_local.yaml")))
    graph["preprocessed_metadata"] = yaml.safe_load(open(os.path.join(base, "preprocessed_metadata.yaml")))
    dump("hetero_graph.json", graph)
    sg = os.path.join(assets35, "split_generator/hetero_node_anchor_based_link_prediction/sgs_output")
    out = {"source": "scala_spark35/common/src/test/assets/split_generator/hetero_node_anchor_based_link_prediction/sgs_output",
           "note": "typed RootedNodeNeighborhoods as the reference holds them (uniform samples of a non-reproducible sampler: "
                   "structure and hydration are pinned, not which neighbours were drawn)"}
    for key, sub in (("rnn_author", "random_negative_rooted_neighborhood_samples/user/samples"),
                     ("rnn_paper", "random_negative_rooted_neighborhood_samples/story/samples")):
        lst = []
        for f in sorted(glob.glob(os.path.join(sg, sub, "*.tfrecord"))):
            for r in read_tfrecords(f):
                m = ts.RootedNodeNeighborhood()
                m.ParseFromString(r)
                lst.append({"root_node": node_to_dict(m.root_node), "neighborhood": graph_to_dict(m.neighborhood),
                            "bytes_b64": base64.b64encode(r).decode()})
        out[key] = lst
    dump("hetero_sgs_output.json", out)


def encoder_views():
    """The host encoder's bytes seen through the reference's own generated protobuf classes."""
    from snapchat.research.gbml import training_samples_schema_pb2 as ts

    sys.path.insert(0, os.path.dirname(HERE))               # tests/
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))  # repo root
    import encoder_cases

    from gigl_b200 import sample_io as sio

    cls = {"rnn": ts.RootedNodeNeighborhood, "snc": ts.SupervisedNodeClassificationSample, "nablp": ts.NodeAnchorBasedLinkPredictionSample}
    out = {"generator": "tests/encoder_cases.py -> gigl_encode_*_host -> snapchat.research.gbml.training_samples_schema_pb2"}
    for name, (kind, data) in encoder_cases.cases().items():
        views = []
        for rec in sio.split_tfrecords(data, verify=True):
            m = cls[kind]()
            m.ParseFromString(rec)
            # nothing the reference's schema does not know: re-serialising what it parsed gives back a record of the same size
            assert len(m.SerializeToString()) == len(rec), (name, len(m.SerializeToString()), len(rec))
            v = {"root_node": node_to_dict(m.root_node), "neighborhood": graph_to_dict(m.neighborhood),
                 "has_neighborhood": m.HasField("neighborhood")}
            if kind == "snc":
                v["root_node_labels"] = [{"label_type": l.label_type, "label": l.label} for l in m.root_node_labels]
            if kind == "nablp":
                v["pos_edges"] = [edge_to_dict(e) for e in m.pos_edges]
                v["hard_neg_edges"] = [edge_to_dict(e) for e in m.hard_neg_edges]
                v["neg_edges"] = [edge_to_dict(e) for e in m.neg_edges]
            views.append(v)
        out[name] = {"kind": kind, "bytes_b64": base64.b64encode(data).decode(), "records": views}
    with open(os.path.join(HERE, "encoder_pb2_views.json"), "w") as f:  # compact: half a megabyte when indented
        json.dump(out, f, sort_keys=True, separators=(",", ":"))
        f.write("\n")
    print("wrote encoder_pb2_views.json")


def main():
    from snapchat.research.gbml import training_samples_schema_pb2 as ts

    if len(sys.argv) > 1 and sys.argv[1] == "hetero":
        return hetero()
    if len(sys.argv) > 1 and sys.argv[1] == "encoder":
        return encoder_views()
    hetero()
    encoder_views()

    # ---- 16-node SNC graph -------------------------------------------------
    base = os.path.join(ASSETS, "subgraph_sampler/supervised_node_classification")
    nodes = [decode_example(r) for r in read_tfrecords(base + "/node_data/data.tfrecord")]
    edges = [decode_example(r) for r in read_tfrecords(base + "/edge_data/data.tfrecord")]
    dump(
        "snc16_graph.json",
        {
            "source": "scala/common/src/test/assets/subgraph_sampler/supervised_node_classification",
            "is_graph_directed": False,
            "num_neighbors_to_sample": 3,
            "nodes": [
                {"node_id": n["node_id"][0], "f0": n["f0"][0], "f1": n["f1"][0], "node_label": n["node_label"][0]}
                for n in nodes
            ],
            "edges": [[e["src"][0], e["dst"][0]] for e in edges],
        },
    )
    sg = os.path.join(ASSETS, "split_generator/supervised_node_classification/sgs_output")
    unl = []
    for r in read_tfrecords(sg + "/unlabeled/samples/data.tfrecord"):
        m = ts.RootedNodeNeighborhood()
        m.ParseFromString(r)
        unl.append({"root_node": node_to_dict(m.root_node), "neighborhood": graph_to_dict(m.neighborhood)})
    lab = []
    for r in read_tfrecords(sg + "/labeled/samples/data.tfrecord"):
        m = ts.SupervisedNodeClassificationSample()
        m.ParseFromString(r)
        lab.append(
            {
                "root_node": node_to_dict(m.root_node),
                "neighborhood": graph_to_dict(m.neighborhood),
                "root_node_labels": [{"label_type": l.label_type, "label": l.label} for l in m.root_node_labels],
            }
        )
    dump(
        "snc16_sgs_output.json",
        {
            "source": "scala/common/src/test/assets/split_generator/supervised_node_classification/sgs_output",
            "note": "produced by the reference sampler with the NON-deterministic shuffle; exact only where degree<=fanout",
            "unlabeled": unl,
            "labeled": lab,
        },
    )
    for tag, p in (
        ("snc16_unlabeled", sg + "/unlabeled/samples/data.tfrecord"),
        ("snc16_edge_data", base + "/edge_data/data.tfrecord"),
        ("snc16_node_data", base + "/node_data/data.tfrecord"),
    ):
        with open(p, "rb") as f, open(os.path.join(HERE, tag + ".tfrecord.b64"), "w") as g:
            g.write(base64.encodebytes(f.read()).decode())
        print("wrote", tag + ".tfrecord.b64")

    # ---- 27-node NABLP graph ----------------------------------------------
    base = os.path.join(ASSETS, "subgraph_sampler/node_anchor_based_link_prediction")
    nodes = [decode_example(r) for r in read_tfrecords(base + "/node_data/data.tfrecord")]
    edges = [decode_example(r) for r in read_tfrecords(base + "/edge_data/data.tfrecord")]
    pos = [decode_example(r) for r in read_tfrecords(base + "/user_defined_pos/data.tfrecord")]
    neg = [decode_example(r) for r in read_tfrecords(base + "/user_defined_neg/data.tfrecord")]
    dump(
        "nablp27_graph.json",
        {
            "source": "scala/common/src/test/assets/subgraph_sampler/node_anchor_based_link_prediction",
            "nodes": [{k: v[0] if len(v) == 1 else v for k, v in n.items()} for n in nodes],
            "edges": [{k: v[0] if len(v) == 1 else v for k, v in e.items()} for e in edges],
            "user_defined_pos": [{k: v[0] if len(v) == 1 else v for k, v in e.items()} for e in pos],
            "user_defined_neg": [{k: v[0] if len(v) == 1 else v for k, v in e.items()} for e in neg],
        },
    )
    sg = os.path.join(ASSETS, "split_generator/node_anchor_based_link_prediction/sgs_output")
    nab = []
    for r in read_tfrecords(sg + "/node_anchor_based_link_prediction_samples/data.tfrecord"):
        m = ts.NodeAnchorBasedLinkPredictionSample()
        m.ParseFromString(r)
        nab.append(
            {
                "root_node": node_to_dict(m.root_node),
                "pos_edges": [edge_to_dict(e) for e in m.pos_edges],
                "hard_neg_edges": [edge_to_dict(e) for e in m.hard_neg_edges],
                "neg_edges": [edge_to_dict(e) for e in m.neg_edges],
                "neighborhood": graph_to_dict(m.neighborhood),
            }
        )
    rnn = []
    for r in read_tfrecords(sg + "/random_negative_rooted_neighborhood_samples/user/data.tfrecord"):
        m = ts.RootedNodeNeighborhood()
        m.ParseFromString(r)
        rnn.append({"root_node": node_to_dict(m.root_node), "neighborhood": graph_to_dict(m.neighborhood)})
    dump(
        "nablp16_sgs_output.json",
        {
            "source": "scala/common/src/test/assets/split_generator/node_anchor_based_link_prediction/sgs_output",
            "nablp": nab,
            "rnn": rnn,
        },
    )

    # ---- XXH64 known answers ----------------------------------------------
    import xxhash

    def s64(u):
        return u - (1 << 64) if u >= (1 << 63) else u

    def hash_int(v, seed=42):
        return s64(xxhash.xxh64(struct.pack("<i", v), seed=seed & 0xFFFFFFFFFFFFFFFF).intdigest())

    # Spark doc: SELECT xxhash64('Spark', array(123), 2) -> 5602566077635097486
    h = xxhash.xxh64(b"Spark", seed=42).intdigest()
    h = xxhash.xxh64(struct.pack("<i", 123), seed=h).intdigest()
    h = xxhash.xxh64(struct.pack("<i", 2), seed=h).intdigest()
    assert s64(h) == 5602566077635097486, h
    ints = [0, 1, 2, 43, 44, 45, 84, 85, 126, 1000, 123456789, 2147483647, -2147483648, -1, -42]
    dump(
        "xxh64_kat.json",
        {
            "spark_doc_kat": {"expr": "xxhash64('Spark', array(123), 2)", "value": 5602566077635097486},
            "hash_int_seed42": [[v, hash_int(v)] for v in ints],
            "generator": "python xxhash %s (xxh64 of the 4 little-endian bytes of the int32, seed 42)" % xxhash.VERSION,
        },
    )


if __name__ == "__main__":
    main()
