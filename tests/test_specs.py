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
