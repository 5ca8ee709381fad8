"""GPU: parity on the SHAPES of BASELINE.json's larger configs at sizes the oracle finishes in seconds -
configs[3] (the g1b generator: directed RMAT(0.57, 0.19, 0.19, 0.05), duplicates kept, hubs, F = 128, 128 -> 128 -> 128,
fanout [15, 10]) and configs[2] (MAG240M as the reference runs it, examples/MAG240M/preprocessor_config.py:72-106: papers and
authors cast to ONE node type, F = 769 with the author rows zero except column 0, directed, numNeighborsToSample = 15 for
both hops; plus its heterogeneous form, two edge types through the typed sampler).  Index sets bit-exact, embeddings within
the float bar of helpers.rel_err."""
import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _rmat(n, e, seed_shift=0):
    from gigl_b200 import synth

    return synth.rmat_edges_numpy(n, e, seed=synth.GEN_SEED + seed_shift)


def test_g1b_shaped_batch_matches_the_oracle():
    """The bench's g1b graph at 1/500 scale (same generator and parameters, so the same hub structure): sampler bit-exact,
    collated node / edge sets exact, root embeddings within the float bar - for a batch of 4096 roots."""
    import torch

    from gigl_b200 import Batch, Context, Graph, SageModel, synth
    from oracle import oracle as O

    n, e, F, fan = 200_000, 2_000_000, 128, [15, 10]
    src, dst = _rmat(n, e)
    ctx = Context.on_torch_stream(0)
    g = Graph.from_edges_host(ctx, n, src.astype(np.int32), dst.astype(np.int32), is_graph_directed=True)
    rowptr, col = O.np_build_in_csr(src, dst, n, True)
    gr, gc = g.csr_host()
    assert np.array_equal(gr, rowptr) and np.array_equal(gc, col)
    deg = np.diff(rowptr)
    assert deg.max() > 2000 and (deg == 0).sum() > n // 20  # hubs and untouched vertices, as on the full-size graph
    roots = np.concatenate([np.flatnonzero(deg > 0)[1000:1000 + 3840], np.flatnonzero(deg == 0)[:256]]).astype(np.int32)
    rt = torch.from_numpy(roots).cuda()
    nbr, cnt = g.sample_khop(rt, fan)
    onbr, ocnt = O.c_sample_khop(rowptr, col, roots, fan)
    for h in range(2):
        assert np.array_equal(nbr[h].cpu().numpy(), onbr[h]) and np.array_equal(cnt[h].cpu().numpy(), ocnt[h])
    assert (onbr[1] >= 0).sum() > 50_000
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, F)).astype(np.float32)
    layers = synth.sage_weights(rng, [F, 128, 128])
    batch = Batch(ctx, n)
    batch.collate(rt, fan, nbr, 2)
    node_ids, ei = batch.export()
    o_nodes, o_ei, _ = O.c_collate(n, roots, onbr, fan)
    assert set(node_ids.cpu().numpy().tolist()) == set(o_nodes.tolist())
    ids = node_ids.cpu().numpy().astype(np.int64)
    got_edges = np.unique((ids[ei[1].cpu().numpy()] << 32) | ids[ei[0].cpu().numpy()])
    want_edges = np.unique((o_nodes[o_ei[1]] << 32) | o_nodes[o_ei[0]])
    assert np.array_equal(got_edges, want_edges)
    out = batch.sage_forward(SageModel(ctx, layers), torch.from_numpy(x).cuda()).cpu().numpy()
    ref64 = O.batch_sage_embeddings(x, roots, onbr, fan, layers, f64=True, n_graph_nodes=n)
    ref32 = O.batch_sage_embeddings(x, roots, onbr, fan, layers, n_graph_nodes=n)
    assert rel_err(out, ref64, ref32) < 1e-5


def _mag_like(n_paper, n_author, cites, writes, seed=3):
    """papers 0 .. n_paper - 1, authors n_paper .. (the casted id ranges of preprocessor_config.py:72-76); paper rows carry
    768 features behind the degree column, author rows are zero except column 0; stored with 3 zero pad columns (772 = 4 * 193)"""
    rng = np.random.default_rng(seed)
    cs, cd = _rmat(n_paper, cites, 1)
    ws = n_paper + rng.integers(0, n_author, writes)
    wd = (rng.zipf(1.5, writes) - 1) % n_paper
    n = n_paper + n_author
    F = 772
    x = np.zeros((n, F), dtype=np.float32)
    x[:n_paper, 1:769] = rng.standard_normal((n_paper, 768)).astype(np.float32)
    deg = np.bincount(np.concatenate([cd, wd, ws]), minlength=n).astype(np.float32)
    x[:, 0] = np.log1p(deg)
    return n, F, (cs, cd), (ws, wd), x


def test_mag_like_homogeneous_cast_matches_the_oracle():
    """configs[2] as the reference runs it: ONE node type, both edge tables merged, directed, fanout [15, 15], F = 769 (+3)."""
    import torch

    from gigl_b200 import Batch, Context, Graph, SageModel, synth
    from oracle import oracle as O

    n, F, (cs, cd), (ws, wd), x = _mag_like(30_000, 30_000, 300_000, 120_000)
    src, dst = np.concatenate([cs, ws]), np.concatenate([cd, wd])
    fan = [15, 15]
    ctx = Context.on_torch_stream(0)
    g = Graph.from_edges_host(ctx, n, src.astype(np.int32), dst.astype(np.int32), is_graph_directed=True)
    rowptr, col = O.np_build_in_csr(src, dst, n, True)
    roots = np.concatenate([np.arange(0, 1500, dtype=np.int32), np.arange(30_000, 30_500, dtype=np.int32)])  # papers and authors
    rt = torch.from_numpy(roots).cuda()
    nbr, cnt = g.sample_khop(rt, fan)
    onbr, ocnt = O.c_sample_khop(rowptr, col, roots, fan)
    for h in range(2):
        assert np.array_equal(nbr[h].cpu().numpy(), onbr[h]) and np.array_equal(cnt[h].cpu().numpy(), ocnt[h])
    assert (ocnt[0][1500:] == 0).all()  # authors have no in-edges in the directed cast: neighbourless roots
    layers = synth.sage_weights(np.random.default_rng(5), [F, 64, 32])
    batch = Batch(ctx, n)
    batch.collate(rt, fan, nbr, 2)
    out = batch.sage_forward(SageModel(ctx, layers), torch.from_numpy(x).cuda()).cpu().numpy()
    ref64 = O.batch_sage_embeddings(x, roots, onbr, fan, layers, f64=True, n_graph_nodes=n)
    ref32 = O.batch_sage_embeddings(x, roots, onbr, fan, layers, n_graph_nodes=n)
    assert rel_err(out, ref64, ref32) < 1e-5


def test_mag_like_two_edge_types_through_the_typed_sampler():
    """The heterogeneous form of configs[2]: paper roots expand `cites` (paper <- paper) and `writes` (paper <- author), then
    the authors' other papers (author -> paper, OUTGOING over `writes`); every op instance bit-exact against the oracle's DAG
    run on a graph of 60k nodes / 420k typed edges."""
    import torch

    from gigl_b200 import Context, Graph, dag
    from oracle import oracle as O

    n, F, (cs, cd), (ws, wd), x = _mag_like(30_000, 30_000, 300_000, 120_000)
    ws = ws - 30_000  # typed ids: authors 0 .. 29999 in their own id space
    cites, writes = ("paper", "cites", "paper"), ("author", "writes", "paper")
    ops = [dag.SamplingOp("cited", cites, 15), dag.SamplingOp("writers", writes, 15),
           dag.SamplingOp("cited2", cites, 10, ["cited"]),
           dag.SamplingOp("their_papers", writes, 5, ["writers"], dag.OUTGOING)]
    ctx = Context.on_torch_stream(0)
    nn = 30_000
    tables = {cites: (cs, cd), writes: (ws, wd)}
    graphs, csrs = {}, {}
    for et, (s_, d_) in tables.items():
        for direction in (dag.INCOMING, dag.OUTGOING):
            graphs[(et, direction)] = Graph.from_edges_host(ctx, nn, s_.astype(np.int32), d_.astype(np.int32), is_graph_directed=True,
                                                            by_source=direction == dag.OUTGOING)
            csrs[(et, direction)] = O.np_build_in_csr(s_, d_, nn, True) if direction == dag.INCOMING else O.np_build_in_csr(d_, s_, nn, True)
    roots = np.arange(100, 100 + 384, dtype=np.int32)
    res = dag.sample_dag(graphs, torch.from_numpy(roots).cuda(), ops, "paper")
    ctx.sync()
    want = O.np_sample_dag(dag.plan(ops, "paper"), lambda p: csrs[(p.op.edge_type, p.op.sampling_direction)], roots)
    for key, (nbr, cnt, _) in res.items():
        assert np.array_equal(nbr.cpu().numpy(), want[key][0]) and np.array_equal(cnt.cpu().numpy(), want[key][1]), key
        assert (cnt.cpu().numpy() > 0).any(), key
