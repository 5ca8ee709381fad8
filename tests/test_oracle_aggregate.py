"""Pins the CPU oracle (aggregate half): hand-computed cases + C fp32 vs C fp64 vs numpy fp64.

PyG 2.5.3 is not installable here, so the pins are: hand-computable graphs,
an independent numpy restatement, and a torch restatement (index_add_ + F.linear, the very ops
PyG's SAGEConv lowers to: MeanAggregation -> scatter -> index_add_, Linear -> F.linear).
"""
import numpy as np
import torch

from oracle import oracle as orc


def test_sage_hand_computed_4_nodes():
    # edges j->i: 1->0, 2->0, 2->0 (duplicate counted twice), 0->3 ; node 1,2 have no in-edges
    ei = np.array([[1, 2, 2, 0], [0, 0, 0, 3]])
    x = np.array([[1.0, 2.0], [3.0, 5.0], [7.0, 11.0], [13.0, 17.0]], np.float32)
    Wl = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], np.float32)
    Wr = np.array([[2.0, 0.0], [0.0, 2.0], [0.0, 0.0]], np.float32)
    bl = np.array([0.5, -0.5, 0.0], np.float32)
    out = orc.c_sage_conv(x, ei, Wl, bl, Wr)
    m0 = (x[1] + 2 * x[2]) / 3
    want = np.array(
        [
            [m0[0] + 0.5 + 2 * 1, m0[1] - 0.5 + 2 * 2, m0[0] + m0[1]],
            [0.5 + 6, -0.5 + 10, 0],  # no in-edge: mean = 0
            [0.5 + 14, -0.5 + 22, 0],
            [1 + 0.5 + 26, 2 - 0.5 + 34, 3],
        ],
        np.float32,
    )
    np.testing.assert_allclose(out, want, rtol=1e-6)
    np.testing.assert_allclose(orc.c_sage_conv(x, ei, Wl, bl, Wr, f64=True), want, rtol=1e-6)
    np.testing.assert_allclose(orc.np_sage_conv(x, ei, Wl, bl, Wr), want, rtol=1e-6)


def test_gcn_hand_computed():
    # 3 nodes, edges 0->1, 1->2, 2->2 (existing self loop kept once), duplicate 2->2
    ei = np.array([[0, 1, 2, 2], [1, 2, 2, 2]])
    x = np.array([[1.0], [2.0], [4.0]], np.float32)
    W = np.array([[1.0], [10.0]], np.float32)
    b = np.array([0.0, 1.0], np.float32)
    # after add_remaining_self_loops: 0->1, 1->2, 0->0, 1->1, 2->2 ; deg = [1, 2, 2]
    d = np.array([1.0, 2.0, 2.0]) ** -0.5
    xp = x @ W.T
    want = np.zeros((3, 2))
    want[0] = d[0] * d[0] * xp[0]
    want[1] = d[0] * d[1] * xp[0] + d[1] * d[1] * xp[1]
    want[2] = d[1] * d[2] * xp[1] + d[2] * d[2] * xp[2]
    want += b
    np.testing.assert_allclose(orc.c_gcn_conv(x, ei, W, b), want, rtol=1e-6)
    np.testing.assert_allclose(orc.c_gcn_conv(x, ei, W, b, f64=True), want, rtol=1e-12)
    np.testing.assert_allclose(orc.np_gcn_conv(x, ei, W, b), want, rtol=1e-12)


def _torch_sage(x, ei, Wl, bl, Wr):
    x = torch.from_numpy(x)
    src, dst = torch.from_numpy(ei[0]), torch.from_numpy(ei[1])
    agg = torch.zeros_like(x).index_add_(0, dst, x[src])
    cnt = torch.zeros(x.shape[0]).index_add_(0, dst, torch.ones(len(dst))).clamp(min=1)
    out = torch.nn.functional.linear(agg / cnt[:, None], torch.from_numpy(Wl), torch.from_numpy(bl))
    return (out + torch.nn.functional.linear(x, torch.from_numpy(Wr))).numpy()


def test_sage_random_c_vs_numpy_vs_torch():
    rng = np.random.default_rng(3)
    for n, e, F, O in ((50, 400, 16, 8), (500, 6000, 100, 47), (200, 0, 7, 3), (300, 5000, 128, 64)):
        x = rng.standard_normal((n, F)).astype(np.float32)
        ei = rng.integers(0, n, (2, e))
        Wl = (rng.standard_normal((O, F)) / np.sqrt(F)).astype(np.float32)
        Wr = (rng.standard_normal((O, F)) / np.sqrt(F)).astype(np.float32)
        bl = rng.standard_normal(O).astype(np.float32)
        ref64 = orc.c_sage_conv(x, ei, Wl, bl, Wr, f64=True)
        np.testing.assert_allclose(orc.np_sage_conv(x, ei, Wl, bl, Wr), ref64, rtol=1e-10, atol=1e-12)
        scale = np.abs(ref64).max()
        for got in (orc.c_sage_conv(x, ei, Wl, bl, Wr), _torch_sage(x, ei, Wl, bl, Wr)):
            assert np.abs(got - ref64).max() <= 1e-5 * scale
        r = orc.c_sage_conv(x, ei, Wl, bl, Wr, relu=True)
        assert (r >= 0).all()


def test_gcn_random_c_vs_numpy():
    rng = np.random.default_rng(4)
    n, e, F, O = 400, 3000, 33, 16
    x = rng.standard_normal((n, F)).astype(np.float32)
    ei = rng.integers(0, n, (2, e))
    W = (rng.standard_normal((O, F)) / np.sqrt(F)).astype(np.float32)
    b = rng.standard_normal(O).astype(np.float32)
    ref = orc.np_gcn_conv(x, ei, W, b)
    np.testing.assert_allclose(orc.c_gcn_conv(x, ei, W, b, f64=True), ref, rtol=1e-9, atol=1e-12)
    assert np.abs(orc.c_gcn_conv(x, ei, W, b) - ref).max() <= 1e-5 * np.abs(ref).max()


def test_collate_is_union_of_per_root_subgraphs():
    O = orc
    """np_collate == the reference's GraphBuilder semantics: union of the per-root edge sets with
    nodes / edges de-duplicated; roots come first; embeddings select the root rows."""
    from helpers import powerlaw_edges

    src, dst = powerlaw_edges(400, 5000, 2)
    rowptr, col = O.np_build_in_csr(src, dst, 400, True)
    roots = np.array([5, 9, 5, 100, 399], dtype=np.int32)
    fan = [4, 3]
    nbr, _ = O.c_sample_khop(rowptr, col, roots, fan)
    nodes, ei, ri = O.np_collate(roots, nbr, fan)
    want = set()
    for lst in O.tree_to_edges(roots, nbr, fan):
        want |= set(lst)
    got = set(zip(nodes[ei[0]].tolist(), nodes[ei[1]].tolist()))
    assert got == want and ei.shape[1] == len(want)
    assert len(set(nodes.tolist())) == len(nodes) and np.array_equal(nodes[ri], roots)
    assert set(nodes.tolist()) == set(roots.tolist()) | {s for s, _ in want} | {d for _, d in want}
    # embeddings: whole-batch forward then root rows; duplicate roots get identical rows
    rng = np.random.default_rng(0)
    x = rng.standard_normal((400, 6)).astype(np.float32)
    layers = [(rng.standard_normal((5, 6)).astype(np.float32), rng.standard_normal(5).astype(np.float32),
               rng.standard_normal((5, 6)).astype(np.float32)),
              (rng.standard_normal((3, 5)).astype(np.float32), None, rng.standard_normal((3, 5)).astype(np.float32))]
    out = O.batch_sage_embeddings(x, roots, nbr, fan, layers, f64=True)
    assert out.shape == (5, 3) and np.array_equal(out[0], out[2])
    full = O.sage_model(x[nodes], ei, layers, f64=True)
    assert np.array_equal(out, full[ri])


def test_fast_collate_torch_forward_and_torch_csr_agree_with_the_plain_oracle():
    from helpers import powerlaw_edges

    O = orc
    n = 600
    src, dst = powerlaw_edges(n, 9000, 7)
    for directed in (True, False):
        rowptr, col = O.np_build_in_csr(src, dst, n, directed)
        tr, tc = O.torch_build_in_csr(torch.from_numpy(src), torch.from_numpy(dst), n, directed)
        assert np.array_equal(tr.numpy(), rowptr) and np.array_equal(tc.numpy(), col)
    roots = np.array([3, 50, 3, 77, 599, 12], dtype=np.int32)
    fan = [5, 3]
    nbr, _ = O.c_sample_khop(rowptr, col, roots, fan)
    n1, e1, r1 = O.np_collate(roots, nbr, fan)
    n2, e2, r2 = O.np_collate_fast(roots, nbr, fan)
    assert np.array_equal(n1[r1], n2[r2]) and set(n1.tolist()) == set(n2.tolist())
    k1 = np.sort((n1[e1[1]] << 32) | n1[e1[0]])
    k2 = np.sort((n2[e2[1]] << 32) | n2[e2[0]])
    assert np.array_equal(k1, k2)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, 10)).astype(np.float32)
    layers = [(rng.standard_normal((8, 10)).astype(np.float32), rng.standard_normal(8).astype(np.float32),
               rng.standard_normal((8, 10)).astype(np.float32)),
              (rng.standard_normal((4, 8)).astype(np.float32), rng.standard_normal(4).astype(np.float32),
               rng.standard_normal((4, 8)).astype(np.float32))]
    a = O.sage_model(x[n2], e2, layers, f64=True)
    b = O.torch_sage_forward(x[n2], e2, layers)
    assert np.abs(a - b).max() / max(1.0, np.abs(a).max()) < 1e-5
