"""GPU: SamplingOp DAGs over typed edges - every op of a tree-shaped DAG against the oracle's per-chain restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _typed_graph(rng, n_user=60, n_item=35):
    n = max(n_user, n_item)
    et = {("user", "follows", "user"): (rng.integers(0, n_user, 400), rng.integers(0, n_user, 400)),
          ("item", "shown_to", "user"): (rng.integers(0, n_item, 300), rng.integers(0, n_user, 300)),
          ("user", "clicks", "item"): (rng.integers(0, n_user, 350), rng.integers(0, n_item, 350))}
    return n, n_user, et


def test_tree_dag_matches_oracle_chains():
    import torch
    from gigl_b200 import Context, Graph, dag
    from oracle import oracle as O

    rng = np.random.default_rng(12)
    n, n_user, et = _typed_graph(rng)
    ctx = Context.on_torch_stream(0)
    ops = [dag.SamplingOp("friends", ("user", "follows", "user"), 3),
           dag.SamplingOp("seen", ("item", "shown_to", "user"), 2),
           dag.SamplingOp("clickers", ("user", "clicks", "item"), 2, ["seen"]),
           dag.SamplingOp("fof", ("user", "follows", "user"), 2, ["friends"]),
           dag.SamplingOp("their_clicks", ("user", "clicks", "item"), 2, ["friends"], dag.OUTGOING)]
    graphs, csrs = {}, {}
    for key, (s, d) in et.items():
        for direction in (dag.INCOMING, dag.OUTGOING):
            graphs[(key, direction)] = Graph.from_edges_host(ctx, n, s, d, is_graph_directed=True, by_source=direction == dag.OUTGOING)
            csrs[(key, direction)] = O.np_build_in_csr(s, d, n, True) if direction == dag.INCOMING else O.np_build_in_csr(d, s, n, True)
    roots = np.arange(0, n_user, 2, dtype=np.int32)
    res = dag.sample_dag(graphs, torch.from_numpy(roots).cuda(), ops, "user")
    ctx.sync()
    planned = {p.op.op_name: p for p in dag.plan(ops, "user")}
    assert [planned[o.op_name].call_no for o in ops] == [1, 2, 3, 4, 5]
    by_name = {o.op_name: o for o in ops}
    for name, (nbr, cnt, fan) in res.items():
        chain = planned[name].chain
        want_nbr, want_cnt = O.np_sample_chain([csrs[(by_name[c].edge_type, by_name[c].sampling_direction)] for c in chain], roots, fan,
                                               [planned[c].call_no for c in chain])
        assert np.array_equal(nbr.cpu().numpy(), want_nbr[-1]), name
        assert np.array_equal(cnt.cpu().numpy(), want_cnt[-1]), name
        assert (cnt.cpu().numpy() > 0).any()
    # the host-buffer entry point gives the same op output
    g = graphs[(("user", "clicks", "item"), dag.INCOMING)]
    seen = res["seen"][0].cpu().numpy()
    nbr_h, cnt_h = g.sample_op_host(roots, [2, 2], [seen], planned["clickers"].call_no)
    assert np.array_equal(nbr_h, res["clickers"][0].cpu().numpy()) and np.array_equal(cnt_h, res["clickers"][1].cpu().numpy())


def test_linear_chain_of_ops_equals_khop():
    import torch
    from gigl_b200 import Context, Graph, dag
    from helpers import powerlaw_edges

    ctx = Context.on_torch_stream(0)
    n = 500
    s, d = powerlaw_edges(n, 6000, seed=2)
    g = Graph.from_edges_host(ctx, n, s, d, is_graph_directed=False)
    et = ("n", "e", "n")
    ops = [dag.SamplingOp("h1", et, 7), dag.SamplingOp("h2", et, 4, ["h1"]), dag.SamplingOp("h3", et, 2, ["h2"])]
    roots = torch.arange(0, n, 3, dtype=torch.int32).cuda()
    res = dag.sample_dag({(et, dag.INCOMING): g}, roots, ops, "n")
    nbr, cnt = g.sample_khop(roots, [7, 4, 2])
    ctx.sync()
    for h, name in enumerate(["h1", "h2", "h3"]):
        assert torch.equal(res[name][0], nbr[h]) and torch.equal(res[name][1], cnt[h])


def test_op_with_several_inputs_runs_once_per_input():
    """An op with two input ops expands the union of their result nodes (GraphDBSampler.scala:66-82): one instance per
    input, each equal to the oracle's chain through that input; a child of such an op has one instance per instance."""
    import torch
    from gigl_b200 import Context, Graph, dag
    from oracle import oracle as O

    rng = np.random.default_rng(13)
    n, n_user, et = _typed_graph(rng)
    ctx = Context.on_torch_stream(0)
    follows, shown, clicks = ("user", "follows", "user"), ("item", "shown_to", "user"), ("user", "clicks", "item")
    ops = [dag.SamplingOp("friends", follows, 3),
           dag.SamplingOp("seen", shown, 2),
           dag.SamplingOp("clickers", clicks, 2, ["seen"]),
           dag.SamplingOp("mixed", follows, 2, ["friends", "clickers"]),
           dag.SamplingOp("deep", shown, 2, ["mixed"])]
    planned = dag.plan(ops, "user")
    assert [p.key for p in planned] == ["friends", "seen", "clickers", "mixed@friends", "mixed@clickers", "deep@mixed@friends",
                                        "deep@mixed@clickers"]
    assert [p.call_no for p in planned] == [1, 2, 3, 4, 4, 5, 5]
    graphs = {(k, dag.INCOMING): Graph.from_edges_host(ctx, n, s, d, is_graph_directed=True) for k, (s, d) in et.items()}
    csr = {k: O.np_build_in_csr(s, d, n, True) for k, (s, d) in et.items()}
    roots = np.arange(0, n_user, 2, dtype=np.int32)
    res = dag.sample_dag(graphs, torch.from_numpy(roots).cuda(), ops, "user")
    ctx.sync()
    by_key = {p.key: p for p in planned}
    for p in planned:
        chain = [by_key[k] for k in p.chain]
        want_nbr, want_cnt = O.np_sample_chain([csr[c.op.edge_type] for c in chain], roots, p.fanouts, [c.call_no for c in chain])
        assert np.array_equal(res[p.key][0].cpu().numpy(), want_nbr[-1]), p.key
        assert np.array_equal(res[p.key][1].cpu().numpy(), want_cnt[-1]), p.key
    assert (res["deep@mixed@clickers"][1].cpu().numpy() > 0).any()
    enc = dag.encoder_ops(planned, res, {follows: 0, shown: 1, clicks: 2}, {"user": 0, "item": 1})
    assert [o["parent"] for o in enc] == [-1, -1, 1, 0, 2, 3, 4]


def test_plan_rejects_unsupported_dags():
    from gigl_b200 import dag

    et = ("user", "follows", "user")
    with pytest.raises(ValueError):
        dag.plan([dag.SamplingOp("a", et, 2), dag.SamplingOp("c", et, 2, ["a", "a"])], "user")  # the same input twice
    with pytest.raises(ValueError):
        dag.plan([dag.SamplingOp("a", et, 2), dag.SamplingOp("c", et, 2, ["a", "nope"])], "user")
    with pytest.raises(ValueError):
        dag.plan([dag.SamplingOp("a", ("item", "shown_to", "user"), 2), dag.SamplingOp("b", et, 2, ["a"])], "user")  # b expands users, a yields items
    with pytest.raises(ValueError):
        dag.plan([dag.SamplingOp("a", et, 2, ["b"]), dag.SamplingOp("b", et, 2, ["a"])], "user")
