"""Sizes of the destination rows a batch collation meets (products-like bench step): entries per dst before de-duplication,
how many rows / entries fall in each size class.  Analysis input for the bucketed collation (csrc/batch_collate.cu)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import WORKLOADS, build_inputs_torch  # noqa: E402
from gigl_b200 import Context, Graph  # noqa: E402
from gigl_b200.sharding import root_batches  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "products-like"]
dev = torch.device("cuda", 0)
ctx = Context.on_torch_stream(0)
src, dst, x, layers = build_inputs_torch(wl, dev)
g = Graph.from_edges_dev(ctx, wl["nodes"], src, dst, is_graph_directed=wl["directed"])
del src, dst, x
fan = [15, 10]
B = 65536
out = {}
for step in (3, 11):
    roots = torch.from_numpy(root_batches(wl["nodes"], 0, 1, B, 1, start_step=step)[0]).to(dev)
    nbr, cnt = g.sample_khop(roots, fan)
    ctx.sync()
    par1 = roots.repeat_interleave(fan[0])
    par2 = nbr[0].repeat_interleave(fan[1])
    d = torch.cat([par1[nbr[0] >= 0], par2[nbr[1] >= 0]]).long()
    s = torch.cat([nbr[0][nbr[0] >= 0], nbr[1][nbr[1] >= 0]]).long()
    rows, counts = torch.unique(d, return_counts=True)
    uk = torch.unique(d * (1 << 32) + s)
    c = counts.cpu().numpy()
    edges = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 65536, 1 << 30]
    hist = {}
    lo = 1
    for hi in edges[1:]:
        m = (c >= lo) & (c < hi)
        hist[f"{lo}-{hi - 1}"] = {"rows": int(m.sum()), "entries": int(c[m].sum())}
        lo = hi
    top = np.sort(c)[::-1][:16].tolist()
    # how many parent slots feed the biggest rows (atomic contention per dst)
    p2 = nbr[0][nbr[0] >= 0].long()
    _, pc = torch.unique(p2, return_counts=True)
    out[f"step{step}"] = {"entries": int(c.sum()), "unique_edges": int(uk.numel()), "rows": int(len(c)), "hist": hist, "top_rows": top,
                          "top_parent_slots_per_dst": torch.sort(pc, descending=True).values[:8].tolist()}
print(json.dumps(out))
