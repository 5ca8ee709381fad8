import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gigl_b200 import Context
ctx = Context.on_torch_stream(0)
R, M, N = 16, 128, 64
print("== m mapping: G one-hot (r,m), A ones -> rows of C that are nonzero")
for (r, m) in [(0,0),(0,1),(0,4),(0,31),(0,32),(0,64),(1,0),(1,1),(1,4),(7,0),(8,0),(9,5),(15,127)]:
    G = torch.zeros(R, M, device='cuda'); G[r, m] = 1
    A = torch.ones(R, N, device='cuda')
    C = ctx.linear_tn(G, A).cpu().numpy()
    nz = np.argwhere(np.abs(C) > 1e-6)
    rows = sorted(set(nz[:,0].tolist())); cols = sorted(set(nz[:,1].tolist()))
    print((r, m), "rows", rows[:8], "ncols", len(cols), "val", C[rows[0], cols[0]] if rows else None)
print("== n mapping: G ones, A one-hot (r,n)")
for (r, n) in [(0,0),(0,1),(0,4),(0,31),(0,32),(1,0),(1,4),(8,3),(15,63)]:
    G = torch.ones(R, M, device='cuda')
    A = torch.zeros(R, N, device='cuda'); A[r, n] = 1
    C = ctx.linear_tn(G, A).cpu().numpy()
    nz = np.argwhere(np.abs(C) > 1e-6)
    rows = sorted(set(nz[:,0].tolist())); cols = sorted(set(nz[:,1].tolist()))
    print((r, n), "cols", cols[:8], "nrows", len(rows))
print("== k mapping: G one-hot (r,0), A one-hot (r2,0)")
for r in [0,1,7,8]:
    hits = []
    for r2 in range(R):
        G = torch.zeros(R, M, device='cuda'); G[r, 0] = 1
        A = torch.zeros(R, N, device='cuda'); A[r2, 0] = 1
        C = ctx.linear_tn(G, A).cpu().numpy()
        if np.abs(C).max() > 1e-6: hits.append((r2, np.argwhere(np.abs(C)>1e-6)[0].tolist()))
    print(r, hits)
