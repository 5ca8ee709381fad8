"""dev tool: projection GEMM timing (CUDA events) at the bench's layer-1 shape; GIGL_GEMM_DBG selects what is skipped."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigl_b200 import Context
ctx = Context.on_torch_stream(0)
dev = torch.device("cuda:0")
K, N, M = int(os.environ.get("K", 200)), int(os.environ.get("N", 256)), int(os.environ.get("M", 375000))
W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
A = torch.randn(M, K, device=dev); C = torch.empty(M, N, device=dev)
for _ in range(3):
    ctx.linear(A, W, b, relu=True, out=C)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ctx.linear(A, W, b, relu=True, out=C)
e1.record(); torch.cuda.synchronize()
print("dbg", os.environ.get("GIGL_GEMM_DBG", "0"), "ms per linear() incl. operand split", e0.elapsed_time(e1) / 10)
