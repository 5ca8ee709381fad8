"""GPU parity of the TRAINING path: forward + backward of SAGEConv / GraphSAGE / GCNConv through gigl_b200.nn (C-ABI
kernels under torch autograd) vs the restated layers under torch CPU autograd in fp64 (oracle.torch_sage_grads /
torch_gcn_grads).  Tolerance: 1e-5 relative (BASELINE.json north_star), as max|got - ref| <= 1e-5 * max(1, max|ref|), for
outputs and parameter gradients.  Input gradients of hub nodes are fp32 sums over thousands of out-edges: there torch's
own fp32 autograd (the reference's arithmetic) is itself up to ~3e-5 away from fp64, so the bar for grad_x is
"no further from fp64 than the reference's fp32 arithmetic, plus 1e-5" - both errors are measured in the test."""
import numpy as np
import pytest

from helpers import powerlaw_edges, uniform_edges

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _orc():
    from oracle import oracle as orc

    return orc


def _rel(got, ref):
    got = got.detach().cpu().numpy() if hasattr(got, "detach") else got
    return float(np.abs(got.astype(np.float64) - ref).max() / max(1.0, np.abs(ref).max()))


def _layers(rng, dims, bias=True):
    out = []
    for a, b in zip(dims[:-1], dims[1:]):
        s = 1.0 / np.sqrt(a)
        out.append((rng.uniform(-s, s, (b, a)).astype(np.float32), rng.uniform(-s, s, b).astype(np.float32) if bias else None,
                    rng.uniform(-s, s, (b, a)).astype(np.float32)))
    return out


@pytest.mark.parametrize("R,M,N", [(1, 1, 1), (7, 5, 3), (1000, 47, 200), (4096, 256, 200), (5000, 256, 512), (333, 130, 77),
                                   (20000, 16, 1538), (0, 8, 8)])
def test_linear_tn_matches_fp64(R, M, N):
    import torch

    from gigl_b200 import Context

    ctx = Context.on_torch_stream(0)
    rng = np.random.default_rng(R + M + N)
    G = rng.standard_normal((R, M)).astype(np.float32)
    A = rng.standard_normal((R, N)).astype(np.float32)
    got = ctx.linear_tn(torch.from_numpy(G).cuda(), torch.from_numpy(A).cuda())
    ref = G.astype(np.float64).T @ A.astype(np.float64)
    assert _rel(got, ref) < RTOL
    got2 = ctx.linear_tn(torch.from_numpy(G).cuda(), torch.from_numpy(A).cuda())
    assert torch.equal(got, got2), "split-K reduction must be run-to-run deterministic"
    acc = ctx.linear_tn(torch.from_numpy(G).cuda(), torch.from_numpy(A).cuda(), out=got.clone(), accumulate=True)
    assert _rel(acc, 2 * ref) < RTOL


@pytest.mark.parametrize("n,e,F,O,relu,bias", [(1, 0, 4, 4, False, True), (300, 4000, 16, 7, True, True), (1000, 20000, 100, 47, False, True),
                                                (777, 9000, 33, 5, True, False), (2000, 50000, 128, 128, True, True),
                                                (600, 30000, 2, 3, False, True)])
def test_sage_conv_backward(n, e, F, O, relu, bias):
    import torch

    from gigl_b200 import nn as gnn

    orc = _orc()
    rng = np.random.default_rng(n + e)
    src, dst = powerlaw_edges(n, e, seed=n) if e else (np.zeros(0, np.int64), np.zeros(0, np.int64))
    ei = np.stack([src, dst])
    x = rng.standard_normal((n, F)).astype(np.float32)
    layers = _layers(rng, [F, O], bias)
    go = rng.standard_normal((n, O)).astype(np.float32)
    # the oracle applies ReLU only between layers; emulate a fused-ReLU single layer with a 2-layer identity trick:
    ref_out, ref_gx, ref_g = _sage_ref(orc, x, ei, layers[0], go, relu)
    gx_fp32 = _sage_ref(orc, x, ei, layers[0], go, relu, f64=False)[1]
    conv = gnn.SAGEConv(F, O, bias=bias).cuda()
    with torch.no_grad():
        conv.lin_l.weight.copy_(torch.from_numpy(layers[0][0]))
        if bias:
            conv.lin_l.bias.copy_(torch.from_numpy(layers[0][1]))
        conv.lin_r.weight.copy_(torch.from_numpy(layers[0][2]))
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    out = conv(xt, torch.from_numpy(ei).cuda(), relu=relu)
    out.backward(torch.from_numpy(go).cuda())
    assert _rel(out, ref_out) < RTOL
    assert _rel(xt.grad, ref_gx) < RTOL + _rel(gx_fp32, ref_gx)
    assert _rel(conv.lin_l.weight.grad, ref_g[0]) < RTOL
    assert _rel(conv.lin_r.weight.grad, ref_g[2]) < RTOL
    if bias:
        assert _rel(conv.lin_l.bias.grad, ref_g[1]) < RTOL


def _sage_ref(orc, x, ei, layer, go, relu, f64=True):
    """single SAGEConv (+ optional fused ReLU) forward/backward in fp64 (or fp32) on the CPU"""
    import torch
    import torch.nn.functional as Fn

    dt = torch.float64 if f64 else torch.float32
    xt = torch.tensor(x, dtype=dt, requires_grad=True)
    Wl = torch.tensor(layer[0], dtype=dt, requires_grad=True)
    bl = None if layer[1] is None else torch.tensor(layer[1], dtype=dt, requires_grad=True)
    Wr = torch.tensor(layer[2], dtype=dt, requires_grad=True)
    src, dst = torch.as_tensor(ei[0]), torch.as_tensor(ei[1])
    n = x.shape[0]
    cnt = torch.zeros(n, dtype=dt).index_add_(0, dst, torch.ones(dst.numel(), dtype=dt)).clamp(min=1.0)
    agg = torch.zeros_like(xt).index_add(0, dst, xt.index_select(0, src)) / cnt[:, None]
    out = Fn.linear(agg, Wl, bl) + Fn.linear(xt, Wr)
    if relu:
        out = out.relu()
    out.backward(torch.tensor(go, dtype=dt))
    return out.detach().numpy(), xt.grad.numpy(), (Wl.grad.numpy(), None if bl is None else bl.grad.numpy(), Wr.grad.numpy())


def test_graphsage_two_layers_matches_autograd_and_is_deterministic():
    import torch

    from gigl_b200 import nn as gnn

    orc = _orc()
    rng = np.random.default_rng(5)
    n, e, dims = 3000, 60000, [100, 256, 47]
    src, dst = powerlaw_edges(n, e, seed=9)
    ei = np.stack([src, dst])
    x = rng.standard_normal((n, dims[0])).astype(np.float32)
    layers = _layers(rng, dims)
    go = rng.standard_normal((n, dims[-1])).astype(np.float32)
    ref_out, ref_gx, ref_g = orc.torch_sage_grads(x, ei, layers, grad_out=go)
    _, gx_fp32, g_fp32 = orc.torch_sage_grads(x, ei, layers, grad_out=go, f64=False)
    model = gnn.GraphSAGE(dims[0], dims[1], 2, dims[2]).cuda()
    assert sorted(model.state_dict().keys()) == sorted(
        [f"convs.{l}.{k}" for l in range(2) for k in ("lin_l.weight", "lin_l.bias", "lin_r.weight")])
    sd = {}
    for l, (Wl, bl, Wr) in enumerate(layers):
        sd[f"convs.{l}.lin_l.weight"], sd[f"convs.{l}.lin_l.bias"], sd[f"convs.{l}.lin_r.weight"] = map(torch.from_numpy, (Wl, bl, Wr))
    model.load_state_dict(sd)
    runs = []
    for _ in range(2):
        model.zero_grad()
        xt = torch.from_numpy(x).cuda().requires_grad_(True)
        out = model(xt, torch.from_numpy(ei).cuda())
        out.backward(torch.from_numpy(go).cuda())
        runs.append((out.detach().clone(), xt.grad.clone(), [p.grad.clone() for p in model.parameters()]))
    out, gx, pg = runs[0]
    assert _rel(out, ref_out) < RTOL and _rel(gx, ref_gx) < RTOL + _rel(gx_fp32, ref_gx)
    for l in range(2):  # layer 0's gradients flow through layer 1's input gradient (the fp32 hub sums): same band
        assert _rel(model.convs[l].lin_l.weight.grad, ref_g[l][0]) < RTOL + _rel(g_fp32[l][0], ref_g[l][0])
        assert _rel(model.convs[l].lin_l.bias.grad, ref_g[l][1]) < RTOL + _rel(g_fp32[l][1], ref_g[l][1])
        assert _rel(model.convs[l].lin_r.weight.grad, ref_g[l][2]) < RTOL + _rel(g_fp32[l][2], ref_g[l][2])
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    assert all(torch.equal(a, b) for a, b in zip(runs[0][2], runs[1][2])), "backward must be bit-identical run to run"


def test_graphsage_pruned_levels_equal_full_rows():
    """level_sizes (the collated batch's dependency levels) computes fewer rows but the same numbers / gradients."""
    import torch

    from gigl_b200 import Batch, Context, Graph
    from gigl_b200 import nn as gnn

    orc = _orc()
    rng = np.random.default_rng(11)
    n, e, dims, fan = 5000, 60000, [32, 64, 16], [5, 3]
    src, dst = powerlaw_edges(n, e, seed=3)
    ctx = Context.on_torch_stream(0)
    g = Graph.from_edges_host(ctx, n, src, dst, is_graph_directed=True)
    roots = torch.arange(0, n, 7, dtype=torch.int32, device="cuda")
    nbr, cnt = g.sample_khop(roots, fan)
    b = Batch(ctx, n)
    levels = b.collate(roots, fan, nbr, 2)
    node_ids, ei = b.export()
    xg = rng.standard_normal((n, dims[0])).astype(np.float32)
    xb = xg[node_ids.cpu().numpy()]
    layers = _layers(rng, dims)
    B = roots.numel()
    go = rng.standard_normal((B, dims[-1])).astype(np.float32)
    ref_out, _, ref_g = orc.torch_sage_grads(xb, ei.cpu().numpy(), layers, grad_out=go, level_sizes=[B, levels[1]], x_requires_grad=False)
    model = gnn.GraphSAGE(dims[0], dims[1], 2, dims[2]).cuda()
    sd = {}
    for l, (Wl, bl, Wr) in enumerate(layers):
        sd[f"convs.{l}.lin_l.weight"], sd[f"convs.{l}.lin_l.bias"], sd[f"convs.{l}.lin_r.weight"] = map(torch.from_numpy, (Wl, bl, Wr))
    model.load_state_dict(sd)
    out = model(torch.from_numpy(xb).cuda(), ei, level_sizes=levels)
    assert out.shape == (B, dims[-1])
    out.backward(torch.from_numpy(go).cuda())
    assert _rel(out, ref_out) < RTOL
    for l in range(2):
        assert _rel(model.convs[l].lin_l.weight.grad, ref_g[l][0]) < RTOL
        assert _rel(model.convs[l].lin_l.bias.grad, ref_g[l][1]) < RTOL
        assert _rel(model.convs[l].lin_r.weight.grad, ref_g[l][2]) < RTOL
    # and the un-pruned model gives the same root rows
    full = model(torch.from_numpy(xb).cuda(), ei)[:B]
    assert _rel(full, ref_out) < RTOL


@pytest.mark.parametrize("n,e,F,O,relu", [(300, 4000, 16, 7, True), (1000, 20000, 100, 16, False), (50, 0, 8, 4, True), (700, 9000, 33, 5, True)])
def test_gcn_conv_backward(n, e, F, O, relu):
    import torch

    from gigl_b200 import nn as gnn

    orc = _orc()
    rng = np.random.default_rng(n)
    src, dst = uniform_edges(n, e, seed=n) if e else (np.zeros(0, np.int64), np.zeros(0, np.int64))
    if e:
        src[:10] = dst[:10]  # explicit self loops are collapsed into the implicit one
    ei = np.stack([src, dst])
    x = rng.standard_normal((n, F)).astype(np.float32)
    W = rng.uniform(-0.3, 0.3, (O, F)).astype(np.float32)
    bvec = rng.uniform(-0.3, 0.3, O).astype(np.float32)
    go = rng.standard_normal((n, O)).astype(np.float32)
    ref_out, ref_gx, ref_gW, ref_gb = orc.torch_gcn_grads(x, ei, W, bvec, relu=relu, grad_out=go)
    gx_fp32 = orc.torch_gcn_grads(x, ei, W, bvec, relu=relu, grad_out=go, f64=False)[1]
    conv = gnn.GCNConv(F, O).cuda()
    with torch.no_grad():
        conv.lin.weight.copy_(torch.from_numpy(W))
        conv.bias.copy_(torch.from_numpy(bvec))
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    out = conv(xt, torch.from_numpy(ei).cuda(), relu=relu)
    out.backward(torch.from_numpy(go).cuda())
    assert _rel(out, ref_out) < RTOL
    assert _rel(xt.grad, ref_gx) < RTOL + _rel(gx_fp32, ref_gx)
    assert _rel(conv.lin.weight.grad, ref_gW) < RTOL
    assert _rel(conv.bias.grad, ref_gb) < RTOL


def test_two_layer_gcn_trains_on_a_separable_toy_task():
    """The reference's own training test asserts 'accuracy above chance / parameters changed' (pyg_training_test.py);
    same bar here, end to end through our kernels + torch.optim.Adam."""
    import torch
    import torch.nn.functional as Fn

    from gigl_b200 import nn as gnn

    rng = np.random.default_rng(0)
    n, classes, F = 600, 3, 12
    labels = rng.integers(0, classes, n)
    src = rng.integers(0, n, 6000)
    # homophilous edges: connect to a node of the same class with probability 0.9
    same = [np.flatnonzero(labels == c) for c in range(classes)]
    dst = np.array([rng.choice(same[labels[s]]) if rng.random() < 0.9 else rng.integers(0, n) for s in src])
    x = (np.eye(classes)[labels] @ rng.standard_normal((classes, F)) + 2.0 * rng.standard_normal((n, F))).astype(np.float32)
    ei = torch.from_numpy(np.stack([np.concatenate([src, dst]), np.concatenate([dst, src])])).cuda()
    torch.manual_seed(0)
    model = gnn.TwoLayerGCN(F, classes, is_training=False).cuda()
    before = [p.detach().clone() for p in model.parameters()]
    opt = torch.optim.Adam(model.parameters(), lr=0.05, weight_decay=5e-4)
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda()
    losses = []
    for _ in range(60):
        opt.zero_grad()
        loss = Fn.cross_entropy(model(xt, ei), yt)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    acc = float((model(xt, ei).argmax(1) == yt).float().mean())
    assert losses[-1] < 0.6 * losses[0] and acc > 0.8, (losses[0], losses[-1], acc)
    assert all(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
