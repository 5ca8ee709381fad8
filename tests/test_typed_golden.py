"""CPU: the typed (heterogeneous) sample path against the reference's own heterogeneous fixture and the typed, hydrated
RootedNodeNeighborhoods it holds for it (tests/golden/hetero_*.json, made by tests/golden/make_golden.py from
scala_spark35/common/src/test/assets/{subgraph_sampler/heterogeneous, split_generator/hetero_node_anchor_based_link_prediction}).
The reference's graph-DB sampler is not reproducible, so what is pinned is structure, hydration and the wire format."""
import base64
import collections

import numpy as np
import pytest

from helpers import load_golden

from gigl_b200 import sample_io as sio
from oracle import oracle as O
from test_sample_assembly import np_edge_rows

# root type -> (edge type of the root op, edge type of the second op): config paths op_4 -> op_6 (author), op_1 -> op_3 (paper)
CHAIN = {0: (1, 0), 1: (0, 1)}
FANOUT = 10


@pytest.fixture(scope="module")
def fx():
    g, out = load_golden("hetero_graph.json"), load_golden("hetero_sgs_output.json")
    x = {}
    for t, key in ((0, "nodes_author"), (1, "nodes_paper")):
        tab = np.zeros((max(r["node_id"] for r in g[key]) + 1, 2), dtype=np.float32)
        for r in g[key]:
            tab[r["node_id"]] = (r["f0"], r["f1"])
        x[t] = tab
    edges = {}
    for t, key in ((0, "edges_author_to_paper"), (1, "edges_paper_to_author")):
        recs = g[key]
        edges[t] = (np.array([r["src"] for r in recs]), np.array([r["dst"] for r in recs]),
                    np.array([[r["f0"], r["f1"]] for r in recs], dtype=np.float32))
    return g, out, x, edges


def _canon_edges(edges):
    return sorted((e.get("condensed_edge_type", 0), e.get("src", e.get("src_node_id", 0)), e.get("dst", e.get("dst_node_id", 0)),
                   tuple(np.float32(e["feature_values"]).tolist())) for e in edges)


def _canon_nodes(nodes):
    return sorted((v.get("condensed_node_type", 0), v["node_id"], tuple(np.float32(v["feature_values"]).tolist())) for v in nodes)


def test_reference_typed_outputs_obey_the_rules_the_encoder_implements(fx):
    g, out, x, edges = fx
    rec = {t: collections.defaultdict(list) for t in edges}
    for t, (s, d, f) in edges.items():
        for a, b, row in zip(s, d, f):
            rec[t][(int(a), int(b))].append(tuple(row.tolist()))
    assert len(out["rnn_author"]) == len(g["nodes_author"]) == 15 and len(out["rnn_paper"]) == len(g["nodes_paper"]) == 19
    for key, rt in (("rnn_author", 0), ("rnn_paper", 1)):
        t1, t2 = CHAIN[rt]
        assert sorted(s["root_node"]["node_id"] for s in out[key]) == list(range(len(out[key])))  # one RNN per node of the type
        for s in out[key]:
            root = s["root_node"]
            # (these fixture records carry an un-hydrated root_node; the root's features are in `nodes`)
            assert root["condensed_node_type"] == rt and root["feature_values"] in ([], x[rt][root["node_id"]].tolist())
            E = s["neighborhood"]["edges"]
            hop1 = {e["src"] for e in E if e["condensed_edge_type"] == t1 and e["dst"] == root["node_id"]}
            ends = {(rt, root["node_id"])}
            per_dst = collections.Counter()
            for e in E:
                t = e["condensed_edge_type"]
                # INCOMING ops: Edge(sampled -> frontier node); hop 1 hangs off the root, hop 2 off a hop-1 node
                assert (t == t1 and e["dst"] == root["node_id"]) or (t == t2 and e["dst"] in hop1)
                # hydrated by the LEFT JOIN on (_from, _to, _condensed_edge_type): a real record and ITS feature row
                assert tuple(np.float32(e["feature_values"]).tolist()) in rec[t][(e["src"], e["dst"])]
                st, dt = (0, 1) if t == 0 else (1, 0)
                ends |= {(st, e["src"]), (dt, e["dst"])}
                per_dst[(t, e["dst"])] += 1
            assert max(per_dst.values(), default=0) <= FANOUT
            assert len(set((e["condensed_edge_type"], e["src"], e["dst"]) for e in E)) == len(E)  # a SET of edges
            nodes = s["neighborhood"]["nodes"]
            assert {(v["condensed_node_type"], v["node_id"]) for v in nodes} == ends and len(nodes) == len(ends)
            for v in nodes:  # every node carries the feature row of its OWN type's table
                assert tuple(np.float32(v["feature_values"]).tolist()) == tuple(x[v["condensed_node_type"]][v["node_id"]].tolist())


def test_parser_reads_the_reference_bytes(fx):
    _, out, _, _ = fx
    for key in ("rnn_author", "rnn_paper"):
        for s in out[key]:
            got = sio.parse_sample(base64.b64decode(s["bytes_b64"]))
            assert got["root_node"]["node_id"] == s["root_node"]["node_id"]
            assert got["root_node"]["condensed_node_type"] == s["root_node"]["condensed_node_type"]
            assert _canon_nodes(got["nodes"]) == _canon_nodes(s["neighborhood"]["nodes"])
            assert _canon_edges(got["edges"]) == _canon_edges(s["neighborhood"]["edges"])


@pytest.mark.parametrize("key,rt", [("rnn_author", 0), ("rnn_paper", 1)])
def test_typed_encoder_reproduces_the_reference_records(fx, key, rt):
    """The reference's own index sets (which neighbours its sampler drew) laid out as the two ops' padded trees and pushed
    through gigl_encode_typed_samples_host with the fixture's node / edge tables: the records must come out as the
    reference wrote them - same typed nodes with the same feature rows, same edges with the same edge features."""
    g, out, x, edges = fx
    t1, t2 = CHAIN[rt]
    samples = sorted(out[key], key=lambda s: s["root_node"]["node_id"])
    roots = np.array([s["root_node"]["node_id"] for s in samples], dtype=np.int32)
    nbr1 = np.full((len(roots), FANOUT), -1, dtype=np.int32)
    nbr2 = np.full((len(roots), FANOUT, FANOUT), -1, dtype=np.int32)
    for i, s in enumerate(samples):
        E = s["neighborhood"]["edges"]
        hop1 = sorted(e["src"] for e in E if e["condensed_edge_type"] == t1 and e["dst"] == roots[i])
        nbr1[i, :len(hop1)] = hop1
        for j, k in enumerate(hop1):
            hop2 = sorted(e["src"] for e in E if e["condensed_edge_type"] == t2 and e["dst"] == k)
            nbr2[i, j, :len(hop2)] = hop2
    ops = [dict(parent=-1, fanout=FANOUT, condensed_edge_type=t1, result_node_type=1 - rt, nbr=nbr1.reshape(-1)),
           dict(parent=0, fanout=FANOUT, condensed_edge_type=t2, result_node_type=rt, nbr=nbr2.reshape(-1))]
    n = max(len(x[0]), len(x[1]))
    tabs = [sio.HostEdgeTable(O.np_build_in_csr(s_, d_, n, True), np_edge_rows(s_, d_, n, True), f_) for s_, d_, f_ in (edges[0], edges[1])]
    xs = [np.concatenate([x[t], np.zeros((n - len(x[t]), 2), np.float32)]) for t in (0, 1)]
    data, offs = sio.encode_typed_samples(roots, rt, ops, xs, tabs, kind="rnn")
    recs = sio.split_tfrecords(data, verify=True)
    assert len(recs) == len(samples)
    for s, rec in zip(samples, recs):
        got = sio.parse_sample(rec)
        assert got["root_node"]["node_id"] == s["root_node"]["node_id"] and got["root_node"]["condensed_node_type"] == rt
        assert np.array_equal(np.float32(got["root_node"]["feature_values"]), x[rt][s["root_node"]["node_id"]])  # hydrateRnn hydrates the root
        assert _canon_nodes(got["nodes"]) == _canon_nodes(s["neighborhood"]["nodes"])
        assert _canon_edges(got["edges"]) == _canon_edges(s["neighborhood"]["edges"])
        # and the record has the reference's exact size - same fields on the wire, only their order inside the sets differs -
        # plus the root's packed feature row (tag + length + 2 floats), which the fixture's root_node lacks
        assert len(rec) == len(base64.b64decode(s["bytes_b64"])) + (10 if not s["root_node"]["feature_values"] else 0)
