"""Kernel-level timings on a products-like synthetic graph (dev tool; bench.py is the contract)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gigl_b200 import Context, Graph, synth

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=2_449_029)
ap.add_argument("--edges", type=int, default=61_859_140)
ap.add_argument("--F", type=int, default=100)
ap.add_argument("--H", type=int, default=256)
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--what", default="sample,gather,sage")
a = ap.parse_args()
dev = torch.device("cuda:0")
ctx = Context.on_torch_stream(0)
t0 = time.time()
src, dst = synth.rmat_edges_torch(a.nodes, a.edges, dev)
torch.cuda.synchronize(); t1 = time.time()
g = Graph.from_edges_dev(ctx, a.nodes, src, dst, is_graph_directed=False)
ctx.sync(); t2 = time.time()
del src, dst
rowptr, col = g.csr_tensors()
deg = (rowptr[1:] - rowptr[:-1])
print(json.dumps({"gen_s": t1 - t0, "build_s": t2 - t1, "n": g.n_nodes, "e": g.n_edges, "max_deg": int(deg.max()),
                  "zero_deg_frac": float((deg == 0).float().mean())}), flush=True)

def timeit(fn, iters):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))

what = a.what.split(",")
if "sample" in what:
    for fan in ([15, 10], [10, 5]):
        gen = torch.Generator(device=dev).manual_seed(1)
        roots = torch.randperm(a.nodes, device=dev, generator=gen)[: a.batch].to(torch.int32)
        out = g.sample_khop(roots, fan)
        med, mn = timeit(lambda: g.sample_khop(roots, fan, out=out), a.iters)
        ctx.sync()
        nE = int((out[0][0] >= 0).sum() + (out[0][1] >= 0).sum())
        # hashes: sum of degrees over frontier
        d1 = deg[roots.long()].sum().item()
        h1 = out[0][0]; v = h1[h1 >= 0].long(); d2 = deg[v].sum().item()
        print(json.dumps({"op": "sample_khop", "fanout": fan, "batch": a.batch, "ms_med": med, "ms_min": mn,
                          "roots_per_s": a.batch / med * 1e3, "sampled_edges": nE, "hashes": d1 + d2,
                          "hashes_per_s": (d1 + d2) / med * 1e3}), flush=True)
if "gather" in what or "sage" in what:
    x = synth.features_torch(a.nodes, a.F, dev)
    rng = np.random.default_rng(0)
    (Wl, bl, Wr), = synth.sage_weights(rng, [a.F, a.H])
    Wl, bl, Wr = (torch.from_numpy(t).to(dev) for t in (Wl, bl, Wr))
    agg = torch.empty(a.nodes, a.F, device=dev)
    if "gather" in what:
        med, mn = timeit(lambda: ctx.gather_mean(x, rowptr, col, out=agg), a.iters)
        byts = g.n_edges * (4 * a.F + 4) + (a.nodes + 1) * 8 + a.nodes * 4 * a.F
        print(json.dumps({"op": "gather_mean_fullgraph", "F": a.F, "ms_med": med, "ms_min": mn, "edges_per_s": g.n_edges / med * 1e3,
                          "alg_GBps": byts / med / 1e6}), flush=True)
    if "sage" in what:
        out = torch.empty(a.nodes, a.H, device=dev)
        med, mn = timeit(lambda: ctx.sage_conv(x, rowptr, col, Wl, bl, Wr, relu=True, out=out), a.iters)
        fl = 2.0 * a.nodes * 2 * a.F * a.H
        print(json.dumps({"op": "sage_conv_fullgraph", "F": a.F, "H": a.H, "ms_med": med, "ms_min": mn,
                          "edges_per_s": g.n_edges / med * 1e3, "gemm_TFLOPs_if_all_gemm": fl / med / 1e9}), flush=True)
ctx.sync()
