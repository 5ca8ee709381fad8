"""GPU parity: batch collation + layer-wise GraphSAGE over the coalesced batch graph, through the
C-ABI, against the oracle's restatement of the reference collate + model(x, edge_index)[roots]."""
import numpy as np
import pytest

from helpers import powerlaw_edges, uniform_edges

pytestmark = pytest.mark.gpu
RTOL = 1e-5  # north_star: within 1e-5 relative on fp32 embeddings


def _orc():
    from oracle import oracle as orc

    return orc


def _rel(got, ref, ref32=None):
    """max-norm relative error, after the element-wise bound of helpers.rel_err has been asserted (absolute term calibrated
    on the fp32 oracle's own error when `ref32` is given)"""
    from helpers import rel_err

    return rel_err(got, ref, ref32)


@pytest.fixture(scope="module")
def env():
    import torch

    from gigl_b200 import Context

    ctx = Context.on_torch_stream(0)
    yield ctx, torch.device("cuda:0")
    ctx.close()


def _setup(ctx, dev, n, e, F, directed, seed, gen=powerlaw_edges):
    import torch

    from gigl_b200 import Batch, Graph

    orc = _orc()
    src, dst = gen(n, e, seed)
    g = Graph.from_edges_host(ctx, n, src, dst, is_graph_directed=directed)
    rowptr, col = orc.np_build_in_csr(src, dst, n, directed)
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, F)).astype(np.float32)
    xt = torch.from_numpy(x).to(dev)
    g.set_features(xt)
    return g, Batch(ctx, n), rowptr, col, x, xt, rng


@pytest.mark.parametrize("directed", [True, False])
@pytest.mark.parametrize("fan,dims", [([3, 2], [16, 8, 4]), ([15, 10], [100, 64, 47]), ([4, 3, 2], [12, 8, 8, 5]), ([5], [7, 3]),
                                      ([5, 4], [200, 300, 24]), ([3, 2], [600, 8, 4]), ([40, 3], [128, 256, 16])])
def test_collate_and_forward_match_oracle(env, directed, fan, dims):
    import torch

    from gigl_b200 import SageModel, synth

    ctx, dev = env
    orc = _orc()
    n = 3000
    g, batch, rowptr, col, x, xt, rng = _setup(ctx, dev, n, 40000, dims[0], directed, 11)
    roots = rng.permutation(n)[:257].astype(np.int32)
    layers = synth.sage_weights(rng, dims)
    model = SageModel(ctx, layers)
    roots_t = torch.from_numpy(roots).to(dev)
    nbr, cnt = g.sample_khop(roots_t, fan)
    sizes = batch.collate(roots_t, fan, nbr, len(layers))
    out = batch.sage_forward(model, xt)
    ctx.sync()
    onbr, _ = orc.c_sample_khop(rowptr, col, roots, fan)
    for a, b in zip(nbr, onbr):
        assert np.array_equal(a.cpu().numpy(), b)
    ref = orc.batch_sage_embeddings(x, roots, onbr, fan, layers, f64=True)
    assert _rel(out.cpu().numpy(), ref, orc.batch_sage_embeddings(x, roots, onbr, fan, layers)) < RTOL
    # the exported batch graph equals the oracle's collation (as sets; local ids are arbitrary)
    node_ids, ei = batch.export()
    ctx.sync()
    onodes, oei, oroot = orc.np_collate(roots, onbr, fan)
    node_ids, ei = node_ids.cpu().numpy().astype(np.int64), ei.cpu().numpy()
    assert np.array_equal(node_ids[: len(roots)], roots) and sizes[0] == len(roots)
    assert np.array_equal(np.sort(node_ids), np.sort(onodes)) and len(np.unique(node_ids)) == len(node_ids)
    got = np.sort((node_ids[ei[1]] << 32) | node_ids[ei[0]])
    want = np.sort((onodes[oei[1]] << 32) | onodes[oei[0]])
    assert np.array_equal(got, want) and batch.n_edges == len(want)
    # a second batch through the same workspace (dense maps are cleaned between batches)
    roots2 = rng.permutation(n)[:100].astype(np.int32)
    r2 = torch.from_numpy(roots2).to(dev)
    nbr2, _ = g.sample_khop(r2, fan)
    batch.collate(r2, fan, nbr2, len(layers))
    out2 = batch.sage_forward(model, xt)
    ctx.sync()
    onbr2, _ = orc.c_sample_khop(rowptr, col, roots2, fan)
    assert _rel(out2.cpu().numpy(), orc.batch_sage_embeddings(x, roots2, onbr2, fan, layers, f64=True),
                orc.batch_sage_embeddings(x, roots2, onbr2, fan, layers)) < RTOL


def test_duplicate_roots_isolated_roots_and_empty(env):
    import torch

    from gigl_b200 import SageModel, synth

    ctx, dev = env
    orc = _orc()
    n = 500
    g, batch, rowptr, col, x, xt, rng = _setup(ctx, dev, n, 900, 20, True, 5, gen=uniform_edges)
    deg = np.diff(rowptr)
    iso = np.flatnonzero(deg == 0)[:5]
    roots = np.concatenate([[7, 7, 9], iso, [7]]).astype(np.int32)
    fan = [4, 3]
    layers = synth.sage_weights(rng, [20, 16, 8])
    model = SageModel(ctx, layers)
    rt = torch.from_numpy(roots).to(dev)
    nbr, _ = g.sample_khop(rt, fan)
    batch.collate(rt, fan, nbr, 2)
    out = batch.sage_forward(model, xt).cpu().numpy()
    onbr, _ = orc.c_sample_khop(rowptr, col, roots, fan)
    ref = orc.batch_sage_embeddings(x, roots, onbr, fan, layers, f64=True)
    assert _rel(out, ref, orc.batch_sage_embeddings(x, roots, onbr, fan, layers)) < RTOL
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[-1])
    # isolated root: embedding = W_r2 relu(W_r1 x + b1) + b2, no neighbours anywhere
    # empty batch
    e0 = torch.zeros(0, dtype=torch.int32, device=dev)
    nbr0, _ = g.sample_khop(e0, fan)
    assert batch.collate(e0, fan, nbr0, 2)[0] == 0
    assert batch.sage_forward(model, xt).shape == (0, 8)


def test_host_entry_point_end_to_end(env):
    from gigl_b200 import SageModel, synth

    ctx, dev = env
    orc = _orc()
    n = 4000
    g, batch, rowptr, col, x, xt, rng = _setup(ctx, dev, n, 80000, 32, False, 21)
    g.set_features_host(x)
    fan = [10, 5]
    layers = synth.sage_weights(rng, [32, 64, 16])
    model = SageModel(ctx, layers)
    roots = np.arange(0, n, 3, dtype=np.int32)
    out, nbr, cnt = g.infer_khop_sage_host(batch, model, roots, fan, return_samples=True)
    onbr, ocnt = orc.c_sample_khop(rowptr, col, roots, fan)
    for h in range(2):
        assert np.array_equal(nbr[h], onbr[h]) and np.array_equal(cnt[h], ocnt[h])
    assert _rel(out, orc.batch_sage_embeddings(x, roots, onbr, fan, layers, f64=True), orc.batch_sage_embeddings(x, roots, onbr, fan, layers)) < RTOL


@pytest.mark.parametrize("fan,directed", [([10, 5], False), ([7], True), ([4, 3, 2], True), ([128, 2], False)])
def test_host_entry_point_packed_index_sets(env, fan, directed):
    """gigl_infer_khop_sage_packed_host: one-byte counts + the filled slots only; unpacking restores the padded tree of
    the oracle exactly, the embeddings are those of the padded call bit for bit, and the call repeats on one workspace."""
    from gigl_b200 import SageModel, synth, unpack_tree

    ctx, dev = env
    orc = _orc()
    n = 3000
    g, batch, rowptr, col, x, xt, rng = _setup(ctx, dev, n, 50000, 32, directed, 29)
    g.set_features_host(x)
    layers = synth.sage_weights(rng, [32, 24, 8])
    model = SageModel(ctx, layers)
    for step in range(2):
        roots = rng.permutation(n)[:500 + 37 * step].astype(np.int32)
        out_p, packed, cnt_u8 = g.infer_khop_sage_packed_host(batch, model, roots, fan)
        out, nbr, cnt = g.infer_khop_sage_host(batch, model, roots, fan, return_samples=True)
        assert np.array_equal(out, out_p)
        onbr, ocnt = orc.c_sample_khop(rowptr, col, roots, fan)
        unbr, ucnt = unpack_tree(packed, cnt_u8, fan)
        assert len(packed) == sum(int((a >= 0).sum()) for a in onbr) and len(packed) < sum(len(a) for a in onbr)
        for h in range(len(fan)):
            assert cnt_u8[h].dtype == np.uint8
            assert np.array_equal(unbr[h], onbr[h]) and np.array_equal(ucnt[h], ocnt[h])
            assert np.array_equal(nbr[h], onbr[h]) and np.array_equal(cnt[h], ocnt[h])
    # a packed buffer that is too small is an error, not a truncated result
    from gigl_b200 import GiglError

    small = (np.empty(10, dtype=np.int32), [np.empty(len(c), dtype=np.uint8) for c in cnt_u8])
    with pytest.raises(GiglError):
        g.infer_khop_sage_packed_host(batch, model, roots, fan, packed_out=small)
    out2, _, _ = g.infer_khop_sage_packed_host(batch, model, roots, fan)  # the ctx is usable after the error
    assert np.array_equal(out2, out_p)
    # the bit-stream form: ceil(log2(n)) bits per id, the same entries in the same order
    from gigl_b200 import unpack_bits

    out_b, words, cnt_b, n_ids, bits = g.infer_khop_sage_bitpacked_host(batch, model, roots, fan)
    assert bits == 12 and n_ids == len(packed) and words.dtype == np.uint32 and len(words) == (n_ids * bits + 31) // 32
    assert np.array_equal(out_b, out_p) and all(np.array_equal(a, c) for a, c in zip(cnt_b, cnt_u8))
    assert np.array_equal(unpack_bits(words, n_ids, bits), packed)
    ids = np.empty(n_ids, dtype=np.int32)
    assert ctx._L.gigl_unpack_bits_host(words.ctypes.data, n_ids, bits, ids.ctypes.data) == 0 and np.array_equal(ids, packed)
    with pytest.raises(GiglError):
        g.infer_khop_sage_bitpacked_host(batch, model, roots, fan, packed_out=(np.empty(3, dtype=np.uint32), cnt_b))


def test_large_batch_properties(env):
    """BASELINE-size style: 65k roots on a 1M-node power-law graph; size-independent properties:
    idempotence (bit-identical embeddings run to run despite atomics in the id assignment),
    batch-size invariance does NOT hold (coalescing) but permutation of the roots permutes the rows."""
    import torch

    from gigl_b200 import Batch, Graph, SageModel, synth

    ctx, dev = env
    n, e, F = 1_000_000, 20_000_000, 100
    src, dst = synth.rmat_edges_torch(n, e, dev)
    g = Graph.from_edges_dev(ctx, n, src, dst, is_graph_directed=False)
    xt = synth.features_torch(n, F, dev)
    g.set_features(xt)
    rng = np.random.default_rng(0)
    layers = synth.sage_weights(rng, [F, 128, 47])
    model = SageModel(ctx, layers)
    batch = Batch(ctx, n)
    fan = [15, 10]
    gen = torch.Generator(device=dev).manual_seed(3)
    roots = torch.randperm(n, device=dev, generator=gen)[:65536].to(torch.int32)
    nbr, _ = g.sample_khop(roots, fan)
    batch.collate(roots, fan, nbr, 2)
    a = batch.sage_forward(model, xt).clone()
    batch.collate(roots, fan, nbr, 2)
    b = batch.sage_forward(model, xt).clone()
    ctx.sync()
    assert torch.equal(a, b)
    perm = torch.randperm(roots.numel(), device=dev, generator=gen)
    roots_p = roots[perm]
    nbr_p, _ = g.sample_khop(roots_p, fan)
    batch.collate(roots_p, fan, nbr_p, 2)
    c = batch.sage_forward(model, xt)
    ctx.sync()
    assert torch.equal(c, a[perm])
    assert bool(torch.isfinite(a).all())


@pytest.mark.parametrize("F", [100, 37, 300])
def test_halo_staging_gives_identical_embeddings(env, F):
    """gigl_batch_set_halo_staging: layer 1 gathers from a per-batch copy of the unique nodes' rows (what a sharded feature
    table uses to move one row per node over NVLink) - bit-identical to gathering from the table itself, and repeatable
    across batches on one workspace."""
    import torch

    from gigl_b200 import SageModel, synth

    ctx, dev = env
    orc = _orc()
    n, fan = 5000, [6, 4]
    g, batch, rowptr, col, x, xt, rng = _setup(ctx, dev, n, 60000, F, False, 23)
    layers = synth.sage_weights(rng, [F, 48, 9])
    model = SageModel(ctx, layers)
    for it in range(3):
        roots = rng.permutation(n)[:400 + 50 * it].astype(np.int32)
        roots_t = torch.from_numpy(roots).to(dev)
        nbr, _ = g.sample_khop(roots_t, fan)
        batch.set_halo_staging(False)
        batch.collate(roots_t, fan, nbr, 2)
        direct = batch.sage_forward(model, xt).clone()
        batch.set_halo_staging(True)
        staged = batch.sage_forward(model, xt).clone()       # same collation, staged layer 1
        batch.collate(roots_t, fan, nbr, 2)
        staged2 = batch.sage_forward(model, xt).clone()      # staging from a fresh collation
        ctx.sync()
        assert torch.equal(direct, staged) and torch.equal(direct, staged2)
        onbr, _ = orc.c_sample_khop(rowptr, col, roots, fan)
        ref = orc.batch_sage_embeddings(x, roots, onbr, fan, layers, f64=True)
        assert _rel(staged.cpu().numpy(), ref, orc.batch_sage_embeddings(x, roots, onbr, fan, layers)) < RTOL
    batch.set_halo_staging(False)
