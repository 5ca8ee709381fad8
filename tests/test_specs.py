"""The Trainer / Inferencer plug-in (gigl_b200.specs): host logic on the CPU (batch building from the reference
sampler's own output protos, interface shape), training end to end on the GPU."""
import base64
import os

import numpy as np
import pytest

from helpers import GOLDEN, load_golden


def _b64(name):
    return base64.b64decode(open(os.path.join(GOLDEN, name)).read())


def test_batch_from_the_reference_samplers_output_follows_the_graph_builder_rules():
    """pyg_graph_builder_test.py / *_batching_test.py semantics: nodes de-duplicated in first-seen order, edges
    de-duplicated, features stacked by local id, roots located - on the reference sampler's real RNN records."""
    import torch

    from gigl_b200 import sample_io as sio
    from gigl_b200.specs import batch_from_sample_protos

    recs = sio.split_tfrecords(_b64("snc16_unlabeled.tfrecord.b64"))
    samples = [sio.parse_sample(r) for r in recs]
    b = batch_from_sample_protos(samples, torch.device("cpu"))
    gold = load_golden("snc16_sgs_output.json")["unlabeled"]
    want_nodes, want_edges, feat = [], set(), {}
    for s in gold:
        for nd in s["neighborhood"]["nodes"]:
            if nd["node_id"] not in feat:
                want_nodes.append(nd["node_id"])
                feat[nd["node_id"]] = nd["feature_values"]
        for ed in s["neighborhood"]["edges"]:
            want_edges.add((ed["src"], ed["dst"]))
    ids = b.node_ids.numpy().tolist()
    assert ids == want_nodes
    assert np.allclose(b.x.numpy(), np.array([feat[i] for i in ids], dtype=np.float32))
    ei = b.edge_index.numpy()
    got_edges = [(ids[a], ids[c]) for a, c in zip(ei[0], ei[1])]
    assert len(got_edges) == len(set(got_edges)) and set(got_edges) == want_edges
    assert [ids[i] for i in b.root_node_indices.tolist()] == [s["root_node"]["node_id"] for s in gold]
    assert b.root_node_labels is None


def test_spec_has_the_reference_operator_interface():
    from gigl_b200.specs import GraphSageB200Spec

    spec = GraphSageB200Spec(optim_lr="0.005", num_epochs="2", out_dim="3", hid_dim="8", in_dim="5", main_sample_batch_size="4")
    for name in ("init_model", "setup_for_training", "train", "eval", "infer_batch", "score", "model", "supports_distributed_training"):
        assert hasattr(spec, name)
    m = spec.init_model(None)
    assert spec.model is m and m.graph_backend == "PYG" and spec.supports_distributed_training
    assert sorted(m.state_dict()) == sorted(f"convs.{l}.{k}" for l in range(2) for k in ("lin_l.weight", "lin_l.bias", "lin_r.weight"))
    sd = {k: v.clone() + 1 for k, v in m.state_dict().items()}
    m2 = GraphSageB200Spec(out_dim="3", hid_dim="8", in_dim="5").init_model(None, state_dict=sd)
    assert all((m2.state_dict()[k] == sd[k]).all() for k in sd)
    g = GraphSageB200Spec(model="gcn", out_dim="3", in_dim="5").init_model({"in_dim": 5})
    assert sorted(g.state_dict()) == ["conv1.bias", "conv1.lin.weight", "conv2.bias", "conv2.lin.weight"]


@pytest.mark.gpu
@pytest.mark.parametrize("kind,prune", [("graphsage", "true"), ("graphsage", "false"), ("gcn", "false")])
def test_trainer_learns_a_homophilous_toy_graph(kind, prune):
    """cfg[0]-shaped plumbing test (toy 1k-node graph, fanout [10, 5], 2 layers): the reference's own trainer test asserts
    that training runs and metrics beat chance (pyg_training_test.py); here the loader, forward, backward all run on the GPU."""
    import torch

    from gigl_b200 import Context, Graph
    from gigl_b200.specs import GraphSageB200Spec, ResidentGraphLoader

    rng = np.random.default_rng(1)
    n, classes, F = 1000, 4, 16
    labels = rng.integers(0, classes, n)
    same = [np.flatnonzero(labels == c) for c in range(classes)]
    src = rng.integers(0, n, 5000)
    dst = np.array([rng.choice(same[labels[s]]) if rng.random() < 0.9 else rng.integers(0, n) for s in src])
    x = (np.eye(classes)[labels] @ rng.standard_normal((classes, F)) + 1.5 * rng.standard_normal((n, F))).astype(np.float32)
    ctx = Context.on_torch_stream(0)
    g = Graph.from_edges_host(ctx, n, src, dst, is_graph_directed=False)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda()
    perm = rng.permutation(n)
    tr, va, te = perm[:600], perm[600:800], perm[800:]
    torch.manual_seed(0)
    spec = GraphSageB200Spec(optim_lr="0.02", num_epochs="8", out_dim=str(classes), hid_dim="32", in_dim=str(F), model=kind,
                             prune_to_roots=prune, is_training=False)
    spec.init_model(None).cuda()
    before = [p.detach().clone() for p in spec.model.parameters()]
    spec.setup_for_training()
    mk = lambda ids, sh: ResidentGraphLoader(g, xd, ids, [10, 5], 128, 2, labels=yd, shuffle=sh)
    spec.loaders = {"train_main": mk(tr, True), "val_main": mk(va, False), "test_main": mk(te, False)}
    res = spec.train(None, torch.device("cuda", 0))
    acc = spec.eval(None, torch.device("cuda", 0))["acc"]
    assert acc > 0.6 and res["best_val_acc"] > 0.6, (res, acc)  # chance = 0.25
    assert all(not torch.equal(a, b) for a, b in zip(before, spec.model.parameters()))
    out = spec.infer_batch(next(iter(spec.loaders["test_main"])), torch.device("cuda", 0))
    assert out.embeddings.shape == (128, classes) and out.predictions.shape == (128,)


@pytest.mark.gpu
def test_pruned_and_full_inference_agree_on_the_roots():
    import torch

    from gigl_b200 import Context, Graph
    from gigl_b200.specs import GraphSageB200Spec, ResidentGraphLoader

    rng = np.random.default_rng(2)
    n, F = 3000, 24
    src, dst = rng.integers(0, n, 40000), rng.integers(0, n, 40000)
    ctx = Context.on_torch_stream(0)
    g = Graph.from_edges_host(ctx, n, src, dst, is_graph_directed=True)
    xd = torch.from_numpy(rng.standard_normal((n, F)).astype(np.float32)).cuda()
    torch.manual_seed(3)
    a = GraphSageB200Spec(out_dim="5", hid_dim="16", in_dim=str(F), prune_to_roots="true")
    a.init_model(None).cuda()
    b = GraphSageB200Spec(out_dim="5", hid_dim="16", in_dim=str(F), prune_to_roots="false")
    b.init_model(None, state_dict=a.model.state_dict()).cuda()
    for batch in ResidentGraphLoader(g, xd, np.arange(0, n, 3), [6, 4], 256, 2):
        ea = a.infer_batch(batch).embeddings
        eb = b.infer_batch(batch).embeddings
        assert float((ea - eb).abs().max()) <= 1e-5 * max(1.0, float(eb.abs().max()))


# ---- the reference's NodeAnchorBasedLinkPrediction / SupervisedNodeClassification batching tests -----------------------
def _node(i):
    return {"node_id": i, "condensed_node_type": 0, "feature_values": [0.0]}


def _edge(s, d, feat=None):
    return {"src_node_id": s, "dst_node_id": d, "condensed_edge_type": 0, "feature_values": list(feat or [])}


def _nablp(root, nodes, edges, pos, hard_neg):
    return {"root_node": _node(root), "nodes": [_node(i) for i in nodes], "edges": [_edge(*e) for e in edges],
            "pos_edges": [_edge(*e) for e in pos], "hard_neg_edges": [_edge(*e) for e in hard_neg], "neg_edges": []}


# node_anchor_based_link_prediction_batching_test.py:60-190: triangle rooted at 0 (pos 0->1, hard neg 0->3), line rooted at 3
# (pos 3->4, hard neg 3->0), chain rooted at 2 (pos 2->3, hard neg 2->4)
_TRIANGLE = _nablp(0, [0, 1, 2, 3], [(0, 1), (0, 2), (1, 2)], [(0, 1)], [(0, 3)])
_LINE = _nablp(3, [3, 4, 0], [(3, 4)], [(3, 4)], [(3, 0)])
_CHAIN = _nablp(2, [1, 2, 3, 4], [(1, 2), (2, 3)], [(2, 3)], [(2, 4)])


@pytest.mark.parametrize("pair,n_nodes,edges,pos,hard_neg", [
    ((_TRIANGLE, _LINE), 5, {(0, 1), (0, 2), (1, 2), (3, 4)}, {0: [1], 3: [4]}, {0: [3], 3: [0]}),      # :391-445 without edge overlap
    ((_TRIANGLE, _CHAIN), 5, {(0, 1), (0, 2), (1, 2), (2, 3)}, {0: [1], 2: [3]}, {0: [3], 2: [4]})])    # :447-502 edge 1->2 shared
def test_link_batch_collation_on_the_reference_batching_cases(pair, n_nodes, edges, pos, hard_neg):
    import torch

    from gigl_b200.specs import link_batch_from_sample_protos

    b = link_batch_from_sample_protos(pair, torch.device("cpu"))
    ids = b.graph.node_ids.tolist()
    assert len(ids) == n_nodes == len(set(ids))
    ei = b.graph.edge_index.numpy()
    got = [(ids[a], ids[c]) for a, c in zip(ei[0], ei[1])]
    assert len(got) == len(edges) and set(got) == edges                      # the shared edge is message-passed once
    assert [ids[r] for r in b.root_nodes.tolist()] == [s["root_node"]["node_id"] for s in pair]
    for data, want in ((b.pos_supervision_edge_data, pos), (b.hard_neg_supervision_edge_data, hard_neg)):
        assert list(data) == [0]                                             # one condensed edge type
        m = data[0].root_node_to_target_node_id
        assert {ids[r]: [ids[t] for t in tg.tolist()] for r, tg in m.items()} == want
        assert data[0].label_edge_features is None and all(tg.dtype == torch.int64 for tg in m.values())
    assert b.edge_attr is None


def test_link_batch_carries_edge_features_of_user_defined_labels():
    """:504-558 - a sample rooted at 5 with feature-rich message-passing edges (2 values) and user-defined label edges (3)."""
    import torch

    from gigl_b200.specs import link_batch_from_sample_protos

    s = _nablp(5, [5, 6, 7], [(5, 6, [0, 1]), (6, 7, [0, 1])], [(5, 6, [0, 1, 2])], [(5, 7, [0, 1, 2])])
    b = link_batch_from_sample_protos([s], torch.device("cpu"))
    assert b.graph.node_ids.numel() == 3 and b.graph.edge_index.shape[1] == 2
    pos, neg = b.pos_supervision_edge_data[0], b.hard_neg_supervision_edge_data[0]
    assert len(pos.root_node_to_target_node_id) == 1 and len(neg.root_node_to_target_node_id) == 1
    assert b.edge_attr.shape == (2, 2)
    for d in (pos, neg):
        for r in d.root_node_to_target_node_id:
            assert d.label_edge_features[r].shape == (1, 3)


def test_node_batch_labels_follow_the_sample_order():
    """supervised_node_classification_batching_test.py:173-235: triangle + line -> 5 nodes / 4 edges, triangle + chain -> 4 / 4,
    root labels in sample order."""
    import torch

    from gigl_b200.specs import batch_from_sample_protos

    def snc(root, nodes, edges, label):
        return {"root_node": _node(root), "nodes": [_node(i) for i in nodes], "edges": [_edge(*e) for e in edges],
                "root_node_labels": [{"label_type": "cls", "label": label}]}

    tri, line, chain = snc(0, [0, 1, 2], [(0, 1), (0, 2), (1, 2)], 7), snc(3, [3, 4], [(3, 4)], 2), snc(1, [1, 2, 3], [(1, 2), (2, 3)], 5)
    for pair, n, e, labels in (((tri, line), 5, 4, [7, 2]), ((tri, chain), 4, 4, [7, 5])):
        b = batch_from_sample_protos(pair, torch.device("cpu"))
        assert b.node_ids.numel() == n and b.edge_index.shape[1] == e and b.root_node_labels.tolist() == labels


def test_link_batch_on_the_reference_samplers_own_nablp_records():
    """The 14 NodeAnchorBasedLinkPredictionSamples the reference's sampler wrote for its toy graph (tests/golden/
    nablp16_sgs_output.json), collated as one batch: the union of their neighbourhoods, every root's positives located."""
    import torch

    from gigl_b200.specs import link_batch_from_sample_protos

    def edge(e):
        return {"src_node_id": e["src"], "dst_node_id": e["dst"], "condensed_edge_type": e["condensed_edge_type"], "feature_values": e["feature_values"]}

    gold = load_golden("nablp16_sgs_output.json")["nablp"]
    samples = [{"root_node": s["root_node"], "nodes": s["neighborhood"]["nodes"], "edges": [edge(e) for e in s["neighborhood"]["edges"]],
                "pos_edges": [edge(e) for e in s["pos_edges"]], "hard_neg_edges": [edge(e) for e in s["hard_neg_edges"]], "neg_edges": []}
               for s in gold]
    b = link_batch_from_sample_protos(samples, torch.device("cpu"))
    ids = b.graph.node_ids.tolist()
    assert sorted(ids) == sorted({n["node_id"] for s in gold for n in s["neighborhood"]["nodes"]})
    ei = b.graph.edge_index.numpy()
    got = {(ids[a], ids[c]) for a, c in zip(ei[0], ei[1])}
    assert got == {(e["src"], e["dst"]) for s in gold for e in s["neighborhood"]["edges"]} and len(got) == ei.shape[1]
    assert [ids[r] for r in b.root_nodes.tolist()] == [s["root_node"]["node_id"] for s in gold]
    pos = b.pos_supervision_edge_data[0].root_node_to_target_node_id
    n_pos = 0
    for s, r in zip(gold, b.root_nodes.tolist()):
        if s["pos_edges"]:
            assert [ids[t] for t in pos[r].tolist()] == [e["dst"] for e in s["pos_edges"]]
            n_pos += len(s["pos_edges"])
    assert n_pos > 0
