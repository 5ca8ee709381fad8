import sys, os, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from gigl_b200 import Batch, Context, Graph, SageModel
from gigl_b200.sharding import root_batches
wl = bench.WORKLOADS["products-like"]; fan=[15,10]; B=65536
dev=torch.device("cuda",0)
ctx=Context(0) if os.environ.get("OWN_STREAM") else Context.on_torch_stream(0)
src,dst,x,layers=bench.build_inputs_torch(wl,dev)
g=Graph.from_edges_dev(ctx,wl["nodes"],src,dst,is_graph_directed=False); del src,dst
g.set_features(x); model=SageModel(ctx,layers); batch=Batch(ctx,wl["nodes"])
batches=root_batches(wl["nodes"],0,1,B,23)
roots_pin=[torch.from_numpy(b).pin_memory() for b in batches]
out_pin=torch.empty((B,wl["O"]),dtype=torch.float32).pin_memory()
nbr_pin,cnt_pin,width=[],[],1
for f in fan:
    cnt_pin.append(torch.empty(B*width,dtype=torch.int32).pin_memory()); width*=f
    nbr_pin.append(torch.empty(B*width,dtype=torch.int32).pin_memory())
s_out=([t.numpy() for t in nbr_pin],[t.numpy() for t in cnt_pin])
def run(samples):
    for i in range(3): g.infer_khop_sage_host(batch,model,roots_pin[i].numpy(),fan,return_samples=samples,out=out_pin.numpy(),samples_out=s_out if samples else None)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for i in range(3,23): g.infer_khop_sage_host(batch,model,roots_pin[i].numpy(),fan,return_samples=samples,out=out_pin.numpy(),samples_out=s_out if samples else None)
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/20*1e3
print("e2e with samples ms", run(True)); print("e2e without samples ms", run(False))
