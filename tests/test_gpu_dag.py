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
    want = O.np_sample_dag(dag.plan(ops, "user"), lambda p: csrs[(p.op.edge_type, p.op.sampling_direction)], roots)
    for name, (nbr, cnt, fan) in res.items():
        assert np.array_equal(nbr.cpu().numpy(), want[name][0]), name
        assert np.array_equal(cnt.cpu().numpy(), want[name][1]), name
        assert (cnt.cpu().numpy() > 0).any()
    # expanding every slot instead (one expansion per path) is each op's chain through the pure-Spark style sampler
    res_paths = dag.sample_dag(graphs, torch.from_numpy(roots).cuda(), ops, "user", distinct_frontier=False)
    ctx.sync()
    by_name = {o.op_name: o for o in ops}
    for name, (nbr, cnt, fan) in res_paths.items():
        chain = planned[name].chain
        want_nbr, want_cnt = O.np_sample_chain([csrs[(by_name[c].edge_type, by_name[c].sampling_direction)] for c in chain], roots, fan,
                                               [planned[c].call_no for c in chain])
        assert np.array_equal(nbr.cpu().numpy(), want_nbr[-1]) and np.array_equal(cnt.cpu().numpy(), want_cnt[-1]), name
    # the host-buffer entry point gives the same op output
    g = graphs[(("user", "clicks", "item"), dag.INCOMING)]
    seen = ctx.frontier_distinct(res["seen"][0], 2).cpu().numpy()
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
    res = dag.sample_dag({(et, dag.INCOMING): g}, roots, ops, "n", distinct_frontier=False)  # one expansion per path = the k-hop sampler
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
    want = O.np_sample_dag(planned, lambda p: csr[p.op.edge_type], roots)
    for p in planned:
        assert np.array_equal(res[p.key][0].cpu().numpy(), want[p.key][0]), p.key
        assert np.array_equal(res[p.key][1].cpu().numpy(), want[p.key][1]), p.key
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


def test_distinct_frontier_set_semantics_beyond_two_hops():
    """GraphDBSampler.scala:66-82: an op expands the SET of its parents' result nodes, every distinct node once per root.
    On a small dense graph a 3-op chain reaches the same node along many paths; per root and op, every distinct frontier
    node with in-neighbours is expanded exactly once, so it receives at most numNodesToSample sampled edges from that op -
    expanding every path instead exceeds the bound on the same inputs.  Also: the kernel equals the oracle's set reduction,
    with and without earlier lists."""
    import torch
    from gigl_b200 import Context, Graph, dag
    from helpers import powerlaw_edges
    from oracle import oracle as O

    ctx = Context.on_torch_stream(0)
    n = 60
    s, d = powerlaw_edges(n, 900, seed=5)
    g = Graph.from_edges_host(ctx, n, s, d, is_graph_directed=False)
    et = ("n", "e", "n")
    fan = [4, 3, 2]
    ops = [dag.SamplingOp("h1", et, fan[0]), dag.SamplingOp("h2", et, fan[1], ["h1"]), dag.SamplingOp("h3", et, fan[2], ["h2"])]
    roots = np.arange(n, dtype=np.int32)
    roots_d = torch.from_numpy(roots).cuda()
    rowptr, col = g.csr_host()
    deg = np.diff(rowptr)

    def edges_per_frontier_node(res):
        for parent, me, slots, f in (("h1", "h2", fan[0], fan[1]), ("h2", "h3", fan[0] * fan[1], fan[2])):
            par = res[parent][0].cpu().numpy().reshape(n, slots)
            cnt = res[me][1].cpu().numpy().reshape(n, slots)
            nbr = res[me][0].cpu().numpy().reshape(n, slots, f)
            for r in range(n):
                per_node = {}
                for sl in range(slots):
                    v = int(par[r, sl])
                    if v >= 0 and cnt[r, sl] > 0:
                        per_node.setdefault(v, []).append(nbr[r, sl, :cnt[r, sl]].tolist())
                for v, groups in per_node.items():
                    yield r, me, v, groups, f

    res = dag.sample_dag({(et, dag.INCOMING): g}, roots_d, ops, "n")
    ctx.sync()
    n_multi_path = 0
    for r, me, v, groups, f in edges_per_frontier_node(res):
        assert len(groups) == 1, (r, me, v)                       # expanded once ...
        assert len(groups[0]) == min(f, deg[v])                   # ... with min(fanout, in-degree) neighbours
    # every distinct frontier node with in-neighbours IS expanded
    for parent, me, slots in (("h1", "h2", fan[0]), ("h2", "h3", fan[0] * fan[1])):
        par = res[parent][0].cpu().numpy().reshape(n, slots)
        cnt = res[me][1].cpu().numpy().reshape(n, slots)
        for r in range(n):
            want = {int(v) for v in par[r] if v >= 0 and deg[v] > 0}
            got = {int(par[r, sl]) for sl in range(slots) if cnt[r, sl] > 0}
            assert got == want
            n_multi_path += len([v for v in par[r] if v >= 0]) - len({int(v) for v in par[r] if v >= 0})
    assert n_multi_path > 0  # the graph does reach nodes along several paths
    over = 0
    res_paths = dag.sample_dag({(et, dag.INCOMING): g}, roots_d, ops, "n", distinct_frontier=False)
    ctx.sync()
    for r, me, v, groups, f in edges_per_frontier_node(res_paths):
        over += len(groups) > 1
    assert over > 0
    # the kernel against the oracle's set reduction
    rng = np.random.default_rng(3)
    cur = rng.integers(-1, 12, size=40 * 24).astype(np.int32)
    p1 = rng.integers(-1, 12, size=40 * 5).astype(np.int32)
    p2 = rng.integers(-1, 12, size=40 * 3).astype(np.int32)
    cur_d, p1_d, p2_d = (torch.from_numpy(a).cuda() for a in (cur, p1, p2))
    assert np.array_equal(ctx.frontier_distinct(cur_d, 24).cpu().numpy(), O.np_frontier_distinct(cur, 24))
    assert np.array_equal(ctx.frontier_distinct(cur_d, 24, [(p1_d, 5), (p2_d, 3)]).cpu().numpy(),
                          O.np_frontier_distinct(cur, 24, [(p1, 5), (p2, 3)]))


@pytest.mark.parametrize("fanout", [3, 40])
def test_weighted_ops_match_oracle(fanout):
    """TopK / RandomWeighted ops (subgraph_sampling_strategy.proto:17-36, NebulaQueryResponseTranslator.scala:73-105): every
    op instance of a DAG that mixes the three sampling methods, bit-exact against oracle.np_sample_op_weighted; weights
    with ties, negatives, zeros, a NaN and an infinity."""
    import torch
    from gigl_b200 import Context, Graph, dag
    from oracle import oracle as O

    rng = np.random.default_rng(31)
    n, n_user, et = _typed_graph(rng, n_user=90, n_item=50)
    hub = (np.full(300, 7), rng.integers(0, n_user, 300))   # user 7 OUTGOING hub / a long in-row for its followers' view
    s0, d0 = et[("user", "follows", "user")]
    et[("user", "follows", "user")] = (np.concatenate([s0, hub[1]]), np.concatenate([d0, hub[0]]))
    ctx = Context.on_torch_stream(0)
    ops = [dag.SamplingOp("best_friends", ("user", "follows", "user"), fanout, [], dag.INCOMING, "top_k", "w"),
           dag.SamplingOp("seen", ("item", "shown_to", "user"), 2),
           dag.SamplingOp("lucky_clickers", ("user", "clicks", "item"), 2, ["seen"], dag.INCOMING, "random_weighted", "w"),
           dag.SamplingOp("fof", ("user", "follows", "user"), 2, ["best_friends"], dag.INCOMING, "random_weighted", "w"),
           dag.SamplingOp("their_top_clicks", ("user", "clicks", "item"), 2, ["best_friends"], dag.OUTGOING, "top_k", "w")]
    graphs, csrs, wts, wts_np = {}, {}, {}, {}
    for key, (s, d) in et.items():
        rec_w = rng.integers(-3, 6, len(s)).astype(np.float32) / 2      # many ties, zeros, negatives
        rec_w[rng.integers(0, len(s), 3)] = [np.nan, np.inf, -0.0]
        for direction in (dag.INCOMING, dag.OUTGOING):
            a, b = (s, d) if direction == dag.INCOMING else (d, s)
            graphs[(key, direction)] = Graph.from_edges_host(ctx, n, s, d, is_graph_directed=True, by_source=direction == dag.OUTGOING)
            csrs[(key, direction)] = O.np_build_in_csr(a, b, n, True)
            rows = ctx.edge_rows_host(n, a, b, True)
            assert np.array_equal(np.asarray(a)[rows], csrs[(key, direction)][1])   # position -> record is the CSR's own order
            wts_np[(key, direction, "w")] = rec_w[rows]
            wts[(key, direction, "w")] = torch.from_numpy(rec_w[rows]).cuda()
    roots = np.arange(0, n_user, 2, dtype=np.int32)
    res = dag.sample_dag(graphs, torch.from_numpy(roots).cuda(), ops, "user", weights=wts)
    ctx.sync()
    planned = dag.plan(ops, "user")
    want = O.np_sample_dag(planned, lambda p: csrs[(p.op.edge_type, p.op.sampling_direction)], roots,
                           weights_of=lambda p: wts_np[(p.op.edge_type, p.op.sampling_direction, p.op.edge_feat_name)])
    for name, (nbr, cnt, fan) in res.items():
        assert np.array_equal(nbr.cpu().numpy(), want[name][0]), name
        assert np.array_equal(cnt.cpu().numpy(), want[name][1]), name
        assert (cnt.cpu().numpy() > 0).any(), name
    cnt = res["best_friends"][1].cpu().numpy()
    rowptr = csrs[(("user", "follows", "user"), dag.INCOMING)][0]
    assert np.array_equal(cnt, np.minimum(fanout, rowptr[roots + 1] - rowptr[roots]))   # LIMIT k of the row, NaN edges included
    with pytest.raises(ValueError):
        dag.sample_dag(graphs, torch.from_numpy(roots).cuda(), ops, "user")   # weighted ops without weights
