#!/usr/bin/env python
"""bench.py - the hot path of BASELINE.json on N B200s of one node.

A "step" = one batch of B roots through the whole path: k-hop rooted-neighbourhood sampling
(fanout [15, 10], GiGL's deterministic hash permutation) -> batch collation -> 2-layer GraphSAGE
over the coalesced batch graph -> root embeddings.

  N = 1   BASELINE.json configs[1]: the ogbn-products-shaped synthetic graph (N = 2,449,029, 61.86M undirected pairs
          mirrored, F = 100 fp32, 100 -> 256 -> 47), whole CSR + features HBM-resident.
  N > 1   the PARTITIONED path of north_star: CSR replicated, the feature table sharded 1/N by node range and mapped
          flat over NVLink, every rank samples its own contiguous range of roots and pulls its batch's remote rows
          (the remote-neighbour feature halo) once per unique node; the replicated variant is measured beside it
          (key "replicated").  At N = 8 the 1e8-node / 1e9-edge graph of configs[3] is run as well (key "g1b").

    python bench.py --gpus N --steps K --warmup W            # the CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import glob
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sampled-subgraphs/sec"
UNIT = "subgraphs/s"
WORKLOADS = {
    # name: nodes, input edge records (undirected pairs unless directed), F, hidden, out, BASELINE.json config it mirrors
    "products-like": dict(nodes=2_449_029, pairs=61_859_140, F=100, H=256, O=47, directed=False, cfg="configs[1]"),
    "products-like-f128": dict(nodes=2_449_029, pairs=61_859_140, F=128, H=256, O=47, directed=False, cfg="configs[1] (F = 128 variant)"),
    "toy-1k": dict(nodes=1_000, pairs=5_000, F=16, H=16, O=7, directed=False, cfg="configs[0]"),
    # N = 1e8, E = 1e9 directed RMAT, F = 128; CSR (4.8 GB) on every GPU, features (51 GB) sharded or replicated
    "g1b": dict(nodes=100_000_000, pairs=1_000_000_000, F=128, H=128, O=128, directed=True, cfg="configs[3]"),
    # a 1/64 scale model of g1b for quick checks of the same code path
    "g1b-small": dict(nodes=1_562_500, pairs=15_625_000, F=128, H=128, O=128, directed=True, cfg="configs[3] at 1/64 scale"),
    # BASELINE.json configs[2], MAG240M as the reference runs it (examples/MAG240M/preprocessor_config.py:72-106,
    # task_config.yaml: papers and authors cast to ONE node type, directed, numNeighborsToSample = 15 for both hops,
    # F = 1 degree column + 768 paper features, author rows zero behind column 0) at 1/64 of its 244 M nodes / 1.68 B edges;
    # the 769 columns are stored as 772 (3 zero columns and zero weight columns: the same sums)
    "mag-like": dict(nodes=3_814_606, pairs=26_250_000, F=772, H=256, O=128, directed=True, fanout="15,15", zero_upper_half=True,
                     cfg="configs[2] (homogeneous cast, 1/64 scale)"),
    # BASELINE.json configs[4]: Inferencer full-graph embedding export = layer-wise inference over EVERY node:
    # N = 1e8 / E = 2e9, 128 -> 128 -> 128; g2b-small = 1/64 scale.  Metric of these two: aggregated-edges/sec.
    "g2b": dict(nodes=100_000_000, pairs=2_000_000_000, F=128, H=128, O=128, directed=True, layerwise=True, cfg="configs[4]"),
    "g2b-eighth": dict(nodes=12_500_000, pairs=250_000_000, F=128, H=128, O=128, directed=True, layerwise=True,
                       cfg="configs[4] at 1/8 scale (one GPU's share of the 8-GPU run)"),
    "g2b-small": dict(nodes=1_562_500, pairs=31_250_000, F=128, H=128, O=128, directed=True, layerwise=True, cfg="configs[4] at 1/64 scale"),
}


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gigl_b200", choices=["gigl_b200", "reference"])
    ap.add_argument("--workload", default="products-like", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=65536, help="roots per step per GPU (both arms)")
    ap.add_argument("--fanout", default=None, help="per-hop fanouts (default: the workload's, 15,10 unless it says otherwise)")
    ap.add_argument("--cpu-steps", type=int, default=2, help="timed steps of the cpu_baseline leg of the product arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-full-graph", action="store_true")
    ap.add_argument("--e2e-padded", action="store_true", help="e2e returns the padded-tree index sets instead of the packed form")
    ap.add_argument("--e2e-ids", choices=["bits", "int32"], default="bits",
                    help="packed e2e form: ids as a ceil(log2 n)-bit stream (gigl_infer_khop_sage_bitpacked_host) or as int32")
    ap.add_argument("--features", default="auto", choices=["auto", "sharded", "replicated"],
                    help="auto = local table at N = 1, sharded over the N GPUs (NVLink halo) at N > 1")
    ap.add_argument("--halo", default="staged", choices=["staged", "direct"],
                    help="sharded features: 'staged' copies each unique batch node's row into local HBM once per step and "
                         "gathers from the copy; 'direct' loads one (mostly remote) row per unique edge inside the gather kernel")
    ap.add_argument("--hot-rows", type=float, default=None,
                    help="sharded features: fraction of the table (highest in-degree vertices) replicated on every GPU "
                         "(default: GIGL_HOT_ROWS or 0 = off)")
    ap.add_argument("--streams", type=int, default=3,
                    help="batches in flight per GPU: step i runs on stream i %% S with its own context / workspace, so the host "
                         "reads and kernel tails of one batch are covered by the other's kernels (1 = strictly one after the other)")
    ap.add_argument("--no-extras", action="store_true", help="N > 1: skip the replicated variant and the g1b record")
    ap.add_argument("--with-g1b", action="store_true", help="run the g1b record at any N (default: N = 8 only)")
    return ap.parse_args(argv)


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML polled every
    2 ms from a thread (the timed region is tens of milliseconds, too short for `nvidia-smi -lms`)."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
           "hw_power_brake_slowdown": 0x80}

    def __init__(self, device_index: int):
        self.idx = device_index
        self.samples = []
        self.reasons = 0
        self.stop_flag = False
        self.h = None
        self.t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            idx = device_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[device_index])
                except (ValueError, IndexError):
                    idx = device_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                try:
                    self.reasons |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    pass
            time.sleep(0.002)

    def start(self):
        if self.h is None:
            return
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=1)
        reasons = sorted(k for k, bit in self.BAD.items() if self.reasons & bit)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s, where from).  Kernels here are timed inside a step, never alone."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained") or d.get("bf16_tflops") or 2250.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 2250.0, "fallback (B200_PROFILING.md: 6.65 TB/s HBM, 2.25 PFLOP/s dense bf16)"


def source_digest():
    """sha256 over the CUDA sources: ties a committed ncu DRAM-traffic record to the kernels it was measured on."""
    h = hashlib.sha256()
    for p in sorted(glob.glob(os.path.join(ROOT, "gigl_b200", "csrc", "*.cu*"))):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def measured_traffic():
    """{kernel phase: DRAM read+write bytes per launch} from the round's ncu capture (profiles/r2_traffic.json, written
    by scripts/ncu_traffic.py from the committed CSV) - only if it was taken on the CUDA sources that are running now."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        if d.get("source_digest") == source_digest():
            return d
    except Exception:
        pass
    return None


def bind_to_gpu_numa_node(device_index: int):
    """N > 1 only: pins this rank's threads to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer is
    allocated, so the e2e path's host buffers live on the socket the GPU's PCIe link hangs off.  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        idx = device_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                idx = int(vis.split(",")[device_index])
            except (ValueError, IndexError):
                idx = device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {w * 64 + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


from gigl_b200.sharding import max_over_ranks, root_batches  # noqa: E402


# ----------------------------------------------------------------------------------------------
# CPU pipeline = the reference algorithm restated (oracle/): sampler and collation in C + OpenMP,
# feature lookup and aggregate with torch's own CPU kernels.  Timed as the baseline; never the product path.
# ----------------------------------------------------------------------------------------------
def cpu_step(O, rowptr, col, x, roots, fan, layers, threads, n_nodes):
    import torch

    nbr, _ = O.c_sample_khop(rowptr, col, roots, fan, n_threads=threads)
    t1 = time.perf_counter()
    node_ids, ei, root_idx = O.c_collate(n_nodes, roots, nbr, fan, n_threads=threads)
    xb = x(node_ids) if callable(x) else x.index_select(0, torch.from_numpy(node_ids)).numpy()
    t2 = time.perf_counter()
    out = O.torch_sage_forward(xb, ei, layers, n_threads=threads)[root_idx]
    return out, ei.shape[1], t1, t2


def run_cpu_baseline(rowptr, col, x, fan, layers, n_nodes, n_roots, steps, warmup):
    import torch

    from oracle import oracle as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    if not callable(x):
        x = torch.from_numpy(x) if isinstance(x, np.ndarray) else x
    batches = root_batches(n_nodes, 0, 1, n_roots, steps + warmup)
    for b in batches[:warmup]:
        cpu_step(O, rowptr, col, x, b, fan, layers, threads, n_nodes)
    t_s = t_c = t_a = 0.0
    edges = 0
    per_step = []
    t0 = time.perf_counter()
    for b in batches[warmup:]:
        ta = time.perf_counter()
        _, e, t1, t2 = cpu_step(O, rowptr, col, x, b, fan, layers, threads, n_nodes)
        tb = time.perf_counter()
        t_s += t1 - ta
        t_c += t2 - t1
        t_a += tb - t2
        per_step.append((tb - ta) * 1e3)
        edges += e * len(layers)
    dt = time.perf_counter() - t0
    return {"value": n_roots * steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} step(s) x {n_roots} roots (ids in order) of the same graph/fanout/model, every host core: C+OpenMP sampler "
                      f"and collation, torch-CPU feature lookup and SAGE on the whole batch graph as the reference does",
            "seconds": dt, "phase_seconds": {"sample": t_s, "collate_and_lookup": t_c, "aggregate": t_a},
            "ms_per_step_min": float(np.min(per_step)), "ms_per_step_median": float(np.median(per_step)),
            "aggregated_edges_per_s": edges / max(t_a, 1e-9)}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ----------------------------------------------------------------------------------------------
def fanout_of(args, wl):
    return [int(v) for v in (args.fanout or wl.get("fanout", "15,10")).split(",")]


def build_inputs_torch(wl, dev, with_features=True):
    """Synthetic graph + features + weights (seeded; identical on every box)."""
    from gigl_b200 import synth

    src, dst = synth.rmat_edges_torch(wl["nodes"], wl["pairs"], dev)
    x = synth.features_torch(wl["nodes"], wl["F"], dev) if with_features else None
    if x is not None and wl.get("zero_upper_half"):  # MAG240M cast: the author rows carry the degree column only
        x[wl["nodes"] // 2:, 1:] = 0
        x[:, 769:] = 0
    layers = synth.sage_weights(np.random.default_rng(synth.GEN_SEED), [wl["F"], wl["H"], wl["O"]])
    return src, dst, x, layers


def workload_config(wl_name, wl, fan, batch, world):
    """The same dict for both arms (--impl gigl_b200 / reference) of one (workload, N)."""
    edges = (f"{wl['pairs']} directed edges (duplicates kept)" if wl["directed"] else
             f"{wl['pairs']} undirected pairs de-duplicated+mirrored")
    return {"workload": f"BASELINE.json {wl['cfg']} shape: {wl_name} synthetic RMAT(0.57,0.19,0.19,0.05) graph, "
                        f"N={wl['nodes']}, {edges}, F={wl['F']} fp32, "
                        f"2-hop fanout {fan}, GraphSAGE {wl['F']}->{wl['H']}->{wl['O']} inference on the coalesced batch graph",
            "roots_per_step_per_gpu": batch, "fanout": fan, "seed": {"generator": 20260101, "sampler_base_seed": 42, "first_call_no": 1},
            "roots": "rank r of N takes the contiguous id range [r*N_nodes/N, (r+1)*N_nodes/N), step s its next B ids",
            "l2": f"every step reads different roots from a {wl['nodes'] * wl['F'] * 4 / 1e9:.2f} GB feature table + the CSR "
                  "(>> 126 MB L2); no flush needed"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; Spark / PyG are not runnable
    here) on the host cores, same workload / metric / roots per step."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    wl = WORKLOADS[args.workload]
    fan = fanout_of(args, wl)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    src, dst, x, layers = build_inputs_torch(wl, dev)  # input preparation only (torch ops, not timed)
    rowptr, col = O.torch_build_in_csr(src, dst, wl["nodes"], wl["directed"])
    rowptr, col, x = rowptr.cpu().numpy(), col.cpu().numpy(), x.cpu()
    del src, dst
    n_roots = min(args.batch, wl["nodes"] // max(world, 1))
    res = run_cpu_baseline(rowptr, col, x, fan, layers, wl["nodes"], n_roots, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["seconds"] / args.steps * 1e3,
            "ms_per_step_min": res["ms_per_step_min"], "ms_per_step_median": res["ms_per_step_median"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, wl, fan, n_roots, world),
            "note": "the reference's CPU pipeline restated (oracle port) on rank 0's host cores; each step = one batch of the workload",
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "phase_seconds": res["phase_seconds"], "aggregated_edges_per_s": res["aggregated_edges_per_s"],
            "cpu_model": cpu_model()}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# the CUDA path
# ----------------------------------------------------------------------------------------------
class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist

        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.numa_cpus = bind_to_gpu_numa_node(self.local) if self.world > 1 else None  # before CUDA / pinned allocations
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL's version banner must not land on stdout beside the JSON line: an explicit level also wins over a
            # NCCL_DEBUG=VERSION from /etc/nccl.conf (the file only fills variables that are unset)
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=self.dev)
        self.dist = dist
        self.torch = torch

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()


class Run:
    """One workload resident on this rank's GPU: graph, features (local / sharded / replicated), model, batch workspace."""

    def __init__(self, env: Env, wl_name: str, features: str, halo: str, batch: int, fan, hot_rows: float, tag: str, streams: int = 1):
        import torch

        from gigl_b200 import Batch, Context, Graph, SageModel, synth

        self.env, self.wl_name, self.wl = env, wl_name, WORKLOADS[wl_name]
        wl, dev = self.wl, env.dev
        self.fan = fan
        self.B = min(batch, wl["nodes"] // env.world)
        self.sharded = features == "sharded"
        self.halo = halo if self.sharded else None
        self.ctx = Context.on_torch_stream(env.local)
        big = wl["nodes"] * wl["F"] * 4 > (8 << 30)
        src, dst, x, self.layers = build_inputs_torch(wl, dev, with_features=not (self.sharded and big))
        self.g = Graph.from_edges_dev(self.ctx, wl["nodes"], src, dst, is_graph_directed=wl["directed"])
        del src, dst
        self.table = None
        self.hot = None
        if self.sharded:
            from gigl_b200.sharding import ShardedFeatureTable

            # rows pitched to whole 128-byte lines: remote rows cross NVLink as full lines (2x the link efficiency at F = 100)
            pitch = -(-wl["F"] // 32) * 32
            self.table = ShardedFeatureTable(self.ctx, wl["nodes"], pitch, env.rank, env.world, tag=tag)
            t = self.table
            if x is not None:
                t.local[: t.row_hi - t.row_lo, : wl["F"]].copy_(x[t.row_lo:t.row_hi])  # this rank keeps only its rows
            else:  # too large to generate whole: every rank fills its own shard
                gen = torch.Generator(device=dev).manual_seed(synth.GEN_SEED + 1 + env.rank)
                rows = t.row_hi - t.row_lo
                for r0 in range(0, rows, 1 << 22):
                    r1 = min(rows, r0 + (1 << 22))
                    t.local[r0:r1, : wl["F"]].copy_(torch.randn(r1 - r0, wl["F"], device=dev, dtype=torch.float32, generator=gen))
            del x
            torch.cuda.synchronize()
            if env.world > 1:
                env.dist.barrier()
            x = t.table[: wl["nodes"], : wl["F"]]
        self.x = x
        self.g.set_features(x)
        self.model = SageModel(self.ctx, self.layers)
        self.batch = Batch(self.ctx, wl["nodes"])
        self.hot_rows = 0
        self.stage_in_sampler = self.sharded and self.halo == "staged" and os.environ.get("GIGL_HALO_EARLY", "1") not in ("0", "collate")
        if self.sharded and self.halo == "staged":
            self.batch.set_halo_staging(True, x if os.environ.get("GIGL_HALO_EARLY", "1") != "0" else None)
            if hot_rows > 0 and env.world > 1 and hasattr(self.batch, "set_hot_rows"):
                self.hot_rows = self.batch.set_hot_rows(self.g, x, hot_rows)
        self.ctx.sync()
        torch.cuda.empty_cache()
        self.out = torch.empty((self.B, wl["O"]), dtype=torch.float32, device=dev)
        self._batches = {}
        self.nbr = self.cnt = None
        # pipes: pipe 0 = the objects above; every further pipe has its own stream, context (scratch, sampler index), graph
        # handle over the SAME resident CSR / feature table, model copy and batch workspace
        self.pipes = [dict(ctx=self.ctx, g=self.g, model=self.model, batch=self.batch, out=self.out, nbr=None, cnt=None, stream=None)]
        for _ in range(1, max(1, streams)):
            st = torch.cuda.Stream(device=dev)
            c = Context(env.local, stream=st.cuda_stream)
            rowptr_t, col_t = self.g.csr_tensors()
            g2 = Graph.wrap_dev(c, rowptr_t, col_t)
            g2.set_features(x)
            b2 = Batch(c, wl["nodes"])
            if self.sharded and self.halo == "staged":
                b2.set_halo_staging(True, x if os.environ.get("GIGL_HALO_EARLY", "1") != "0" else None)
                if self.hot_rows and getattr(self.batch, "_hot", None) is not None:
                    b2.share_hot_rows(self.batch, x.shape[1])
            self.pipes.append(dict(ctx=c, g=g2, model=SageModel(c, self.layers), batch=b2,
                                   out=torch.empty((self.B, wl["O"]), dtype=torch.float32, device=dev), nbr=None, cnt=None, stream=st))
            c.sync()

    def roots(self, n_steps):
        key = n_steps
        if key not in self._batches:
            host = root_batches(self.wl["nodes"], self.env.rank, self.env.world, self.B, n_steps)
            self._batches[key] = (host, [self.env.torch.from_numpy(b).to(self.env.dev) for b in host])
        return self._batches[key]

    def step(self, roots_dev, pipe: int = 0):
        p = self.pipes[pipe]
        if p["nbr"] is None:
            p["nbr"], p["cnt"] = p["g"].sample_khop(roots_dev, self.fan)
            if pipe == 0:
                self.nbr, self.cnt = p["nbr"], p["cnt"]
        p["g"].sample_khop(roots_dev, self.fan, out=(p["nbr"], p["cnt"]), stage_into=p["batch"] if self.stage_in_sampler else None)
        p["batch"].collate(roots_dev, self.fan, p["nbr"], 2)
        p["batch"].sage_forward(p["model"], self.x, out=p["out"])
        return p["batch"].n_edges

    # ---- device-resident measurement (value) --------------------------------------------------
    def measure_device(self, K, W, streams=1):
        """K timed steps after W warm-up steps, step i on pipe i % streams.  streams = 1 also collects the per-phase device
        times (with several batches in flight the phases of different batches overlap, so they are taken from this form)."""
        torch, env = self.env.torch, self.env
        S = max(1, min(streams, len(self.pipes)))
        pipes = self.pipes[:S]
        _, roots_dev = self.roots(K + W)
        for i in range(max(W, S)):
            self.step(roots_dev[i % (K + W)], i % S)
        if S == 1:
            self.ctx.set_timing(True)
            self.ctx.reset_timing()
        l0 = sum(p["ctx"].launch_count for p in pipes)
        clocks = ClockSampler(env.local)
        main = torch.cuda.current_stream(env.dev)
        ev0 = torch.cuda.Event(enable_timing=True)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        env.barrier()
        clocks.start()
        ev0.record(main)
        for p in pipes[1:]:
            p["stream"].wait_event(ev0)
        e1_total = 0
        for i in range(W, W + K):
            k = (i - W) % S
            e1_total += self.step(roots_dev[i], k)
            evs[i - W].record(pipes[k]["stream"] or main)
        for p in pipes[1:]:
            main.wait_stream(p["stream"])
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record(main)
        env.barrier()
        ms = ev0.elapsed_time(ev1)
        done = sorted(ev0.elapsed_time(e) for e in evs)
        per_step = [done[0]] + [done[i] - done[i - 1] for i in range(1, K)]  # time between consecutive batch completions
        clk = clocks.stop()
        launches = sum(p["ctx"].launch_count for p in pipes) - l0
        timings = self.ctx.timings() if S == 1 else {}
        if S == 1:
            self.ctx.set_timing(False)
        ms = max_over_ranks(ms, env.dev)
        return {"ms": ms, "per_step": per_step, "clocks": clk, "launches": int(launches), "timings": timings, "e1_total": e1_total,
                "streams": S}

    def counts(self, K, W):
        """what a step touches (untimed recount): unique edges, rows per layer, batch nodes, sampled edges, frontier rows"""
        _, roots_dev = self.roots(K + W)
        tot = dict(e1=0, e2=0, n1=0, nodes=0, sampled=0, frontier=0, slots=0)
        for i in range(W, W + K):
            self.g.sample_khop(roots_dev[i], self.fan, out=(self.nbr, self.cnt))
            sizes = self.batch.collate(roots_dev[i], self.fan, self.nbr, 2)
            tot["e1"] += self.batch.n_edges
            tot["n1"] += sizes[1]
            node_ids, ei = self.batch.export()
            tot["e2"] += int((ei[1] < self.B).sum().item())
            tot["nodes"] += int(node_ids.numel())
            valid = [int((t >= 0).sum().item()) for t in self.nbr]
            tot["sampled"] += sum(valid)
            tot["frontier"] += self.B + sum(valid[:-1])  # rows the sampler resolves: the roots + every filled slot above the last hop
            tot["slots"] += sum(int(t.numel()) for t in self.nbr)
        return {k: v / K for k, v in tot.items()}

    # ---- end to end through the host entry points ----------------------------------------------
    def measure_e2e(self, K, W, padded=False, embeddings_only=False, streams=1, ids="bits"):
        """The same K batches through the blocking host entry points: one caller thread per pipe (the call is synchronous -
        roots in, results in host memory out - so several batches are in flight only if several callers are)."""
        torch, env = self.env.torch, self.env
        S = max(1, min(streams, len(self.pipes)))
        host, _ = self.roots(K + W)
        B, O_dim, fan = self.B, self.wl["O"], self.fan
        roots_pin = [torch.from_numpy(b).pin_memory() for b in host]
        d2h_steps = []
        steps_of = []
        for k in range(S):
            p = self.pipes[k]
            out_pin = torch.empty((B, O_dim), dtype=torch.float32).pin_memory()
            if embeddings_only:
                def host_step(i, p=p, out_pin=out_pin):
                    p["g"].infer_khop_sage_host(p["batch"], p["model"], roots_pin[i].numpy(), fan, out=out_pin.numpy())
                    d2h_steps.append(B * O_dim * 4)
                api = "gigl_infer_khop_sage_host(nbr_out = NULL): roots in pinned host memory -> root embeddings back in pinned host memory"
            else:
                nbr_pin, cnt_pin, cnt8_pin, width = [], [], [], 1
                for f in fan:
                    cnt_pin.append(torch.empty(B * width, dtype=torch.int32).pin_memory())
                    cnt8_pin.append(torch.empty(B * width, dtype=torch.uint8).pin_memory())
                    width *= f
                    nbr_pin.append(torch.empty(B * width, dtype=torch.int32).pin_memory())
                if padded:
                    s_out = ([t.numpy() for t in nbr_pin], [t.numpy() for t in cnt_pin])

                    def host_step(i, p=p, out_pin=out_pin, s_out=s_out, n_ints=sum(t.numel() for t in nbr_pin + cnt_pin)):
                        p["g"].infer_khop_sage_host(p["batch"], p["model"], roots_pin[i].numpy(), fan, return_samples=True,
                                                    out=out_pin.numpy(), samples_out=s_out)
                        d2h_steps.append(B * O_dim * 4 + n_ints * 4)
                    api = "gigl_infer_khop_sage_host: roots in pinned host memory -> padded-tree index sets + root embeddings in pinned host memory"
                elif ids == "bits":
                    words_pin = torch.empty(sum(t.numel() for t in nbr_pin) + 1, dtype=torch.int32).pin_memory()
                    w_out = (words_pin.numpy().view(np.uint32), [t.numpy() for t in cnt8_pin])

                    def host_step(i, p=p, out_pin=out_pin, w_out=w_out, n_cnt=sum(t.numel() for t in cnt8_pin)):
                        _, words, _, _, _ = p["g"].infer_khop_sage_bitpacked_host(p["batch"], p["model"], roots_pin[i].numpy(), fan,
                                                                                  out=out_pin.numpy(), packed_out=w_out)
                        d2h_steps.append(B * O_dim * 4 + words.size * 4 + n_cnt)
                    api = ("gigl_infer_khop_sage_bitpacked_host: roots in pinned host memory -> packed index sets [one-byte counts + filled "
                           "slots as a ceil(log2 n_nodes)-bit stream] + root embeddings in pinned host memory")
                else:
                    packed_pin = torch.empty(sum(t.numel() for t in nbr_pin), dtype=torch.int32).pin_memory()
                    p_out = (packed_pin.numpy(), [t.numpy() for t in cnt8_pin])

                    def host_step(i, p=p, out_pin=out_pin, p_out=p_out, n_cnt=sum(t.numel() for t in cnt8_pin)):
                        _, packed, _ = p["g"].infer_khop_sage_packed_host(p["batch"], p["model"], roots_pin[i].numpy(), fan,
                                                                          out=out_pin.numpy(), packed_out=p_out)
                        d2h_steps.append(B * O_dim * 4 + packed.size * 4 + n_cnt)
                    api = ("gigl_infer_khop_sage_packed_host: roots in pinned host memory -> packed index sets [one-byte counts + filled "
                           "slots] + root embeddings in pinned host memory")
            steps_of.append(host_step)
        for i in range(max(W, S)):
            steps_of[i % S](i % (K + W))
        d2h_steps.clear()

        def caller(k):
            for i in range(W + k, W + K, S):
                steps_of[k](i)

        main = torch.cuda.current_stream(env.dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env.barrier()
        ev0.record(main)
        for p in self.pipes[1:S]:
            p["stream"].wait_event(ev0)
        t0 = time.perf_counter()
        if S == 1:
            caller(0)
        else:
            ths = [threading.Thread(target=caller, args=(k,)) for k in range(S)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        for p in self.pipes[1:S]:
            main.wait_stream(p["stream"])
        ev1.record(main)
        env.barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms_e = max_over_ranks(ev0.elapsed_time(ev1), env.dev)  # device time from the first call's start to the last result's arrival
        d2h = int(sum(d2h_steps) / max(len(d2h_steps), 1))
        return {"value": env.world * B * K / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 4, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e / K, "host_wall_ms_per_step": wall_ms / K, "host_cpus_bound": env.numa_cpus, "api": api,
                "callers": S}

    def full_graph(self):
        """the aggregate over the WHOLE graph (every node a row, every CSR edge reduced once): the form the Trainer's nn
        modules and layer-wise inference use; reported beside the batch numbers, outside the timed step"""
        torch, env, wl = self.env.torch, self.env, self.wl
        F = wl["F"]
        peak, _, _ = measured_peaks()
        rowptr_t, col_t = self.g.csr_tensors()
        agg = torch.empty((wl["nodes"], F), dtype=torch.float32, device=env.dev)
        self.ctx.gather_mean(self.x, rowptr_t, col_t, out=agg)
        fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env.barrier()
        fe0.record()
        for _ in range(5):
            self.ctx.gather_mean(self.x, rowptr_t, col_t, out=agg)
        fe1.record()
        env.barrier()
        f_ms = max_over_ranks(fe0.elapsed_time(fe1) / 5, env.dev)
        f_bytes = self.g.n_edges * (4 * F + 4) + (wl["nodes"] + 1) * 8 + wl["nodes"] * 4 * F  # SURVEY 8(d) bytes_A without the projection
        return {"op": "gigl_gather_mean_dev over the whole CSR (SAGEConv mean aggregate of every node)", "edges": int(self.g.n_edges),
                "ms": f_ms, "aggregated_edges_per_sec": env.world * self.g.n_edges / (f_ms * 1e-3),
                "algorithmic_GBps_per_gpu": f_bytes / f_ms / 1e6, "frac_of_hbm_peak": f_bytes / f_ms / 1e6 / peak}

    def cpu_baseline(self, steps):
        torch, wl = self.env.torch, self.wl
        rowptr_t, col_t = self.g.csr_tensors()
        x = self.x
        if x.numel() * 4 > (8 << 30) or self.sharded:
            # a table this large (or sharded) is not copied to the host: the baseline's feature lookup reads the device copy
            x_host = lambda ids: x[torch.from_numpy(ids).to(self.env.dev)].cpu().numpy()  # noqa: E731
        else:
            x_host = x.cpu()
        cpu = run_cpu_baseline(rowptr_t.cpu().numpy(), col_t.cpu().numpy(), x_host, self.fan, self.layers, wl["nodes"], self.B, steps, 1)
        if callable(x_host):
            cpu["sample"] += "; batch feature rows fetched from the device-resident table"
        cpu["cpu_model"] = cpu_model()
        return cpu

    def residency(self):
        if not self.sharded:
            return "whole CSR + feature table resident in HBM on every GPU (replicated); roots sharded by contiguous id range"
        s = ("CSR replicated; feature table sharded by contiguous node range over the GPUs and mapped as one flat array "
             "(cuMemMap of peer shards, rows pitched to 128-byte multiples): ")
        if self.halo == "staged":
            s += ("the row of every unique batch node is copied over NVLink into a per-batch table once per step (halo staging), "
                  "layer 1 gathers from the copy")
            if self.hot_rows:
                s += f"; the {self.hot_rows} highest in-degree rows ({self.hot_rows / self.wl['nodes']:.3f} of the table) are replicated on every GPU"
        else:
            s += "remote neighbour rows are loaded over NVLink inside the gather kernel, one per unique edge"
        return s

    def close(self):
        self.ctx.sync()
        if self.env.world > 1:
            self.env.dist.barrier()
        for p in self.pipes[1:]:
            p["ctx"].sync()
            p["batch"].close()
            p["model"].close()
            p["g"].close()
            p["ctx"].close()
        self.pipes = []
        self.batch.close()
        self.model.close()
        self.g.close()
        if self.table is not None:
            self.table.close()
        self.nbr = self.cnt = self.out = self.x = None
        self._batches = {}
        self.ctx.close()
        self.env.torch.cuda.empty_cache()


def roofline_blocks(run: Run, dev_res, cnt, K):
    """One block per hot kernel: SURVEY.md 8(d) algorithmic bytes (no cache credit, no private byte model) / the kernel's
    CUDA-event time inside the timed region, against the measured peak; `traffic` = ncu DRAM bytes per launch of the same
    kernels, taken this round on the same sources (else null)."""
    wl = run.wl
    F, H = wl["F"], wl["H"]
    hbm, bf16, src = measured_peaks()
    tr = measured_traffic() if (run.wl_name == "products-like" and run.B == 65536 and not run.sharded) else None
    timings, step_ms = dev_res["timings"], dev_res["ms"] / K

    def phase(*names):
        ms = sum(timings.get(n, (0.0, 0))[0] for n in names) / K
        return ms if ms > 0 else None

    def hbm_block(kernel, names, bytes_, formula, key):
        ms = phase(*names)
        ach = bytes_ / (ms * 1e-3) / 1e9 if ms else None
        t = tr.get(key) if tr else None
        return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm if ach else None,
                "traffic": t.get("dram_bytes") if t else None,
                "dram_read_frac_of_peak": (t["dram_read_bytes"] / (ms * 1e-3) / 1e9 / hbm) if (t and ms and t.get("dram_read_bytes")) else None,
                "algorithmic_bytes_per_launch": bytes_, "formula": formula, "ms_per_launch": ms, "share_of_step": ms / step_ms if ms else None,
                "peak_source": src}

    e1, n1, nE, rows_s, nodes, slots = cnt["e1"], cnt["n1"], cnt["sampled"], cnt["frontier"], cnt["nodes"], cnt["slots"]
    blocks = []
    blocks.append(hbm_block("khop_tile_kernel<16> (all hops of the sampler)", ["sample"], 16 * rows_s + 12 * nE + 8 * run.B,
                            "8(d) bytes_S = sum over frontier rows (16 + 4 min(deg, f)) + 8 nE_sampled + 8 per root = 16 rows + 12 nE + 8 B",
                            "sample"))
    blocks.append(hbm_block("batch_gather_async_kernel<1, 8> + split-row parts / finish (layer-1 gather over the coalesced batch graph)",
                            ["gather_l1"], e1 * (4 * F + 4) + (n1 + 1) * 8 + n1 * 4 * F,
                            "8(d) bytes_A gather terms: e (4 F + 4) + (n + 1) 8 + n 4 F", "gather_l1"))
    ms_g = phase("gemm_l1")
    if ms_g:
        flops = 2.0 * n1 * (2 * F) * H
        peak_tf32 = bf16 / 2.0
        blocks.append({"kernel": "linear_tf32x3_kernel (layer-1 projection, 3 tcgen05 kind::tf32 passes per product)", "bound": "tensor",
                       "achieved": 3 * flops / (ms_g * 1e-3) / 1e12, "peak": peak_tf32, "unit": "TFLOP/s",
                       "frac": 3 * flops / (ms_g * 1e-3) / 1e12 / peak_tf32, "traffic": (tr or {}).get("gemm_l1", {}).get("dram_bytes"),
                       "algorithmic_flops_per_launch": flops, "executed_flops_per_launch": 3 * flops,
                       "formula": "8(d) flops_A = 2 n (2 F) H; executed = 3 x (hi*hi + hi*lo + lo*hi)", "ms_per_launch": ms_g,
                       "share_of_step": ms_g / step_ms, "peak_source": src + "; tf32 dense peak taken as bf16 / 2"})
    blocks.append(hbm_block("layer-1 gather-SpMM as a whole (gather + projection)", ["gather_l1", "gemm_l1"],
                            e1 * (4 * F + 4) + (n1 + 1) * 8 + n1 * 4 * F + n1 * 4 * H + 4 * (2 * F * H + H),
                            "8(d) bytes_A = e (4 F + 4) + (n + 1) 8 + n 4 F + n 4 F_out + 4 (2 F F_out + F_out)", "gather_spmm_l1"))
    blocks.append(hbm_block("batch collation (tree_rows / rows_alloc / rows_sort_* / expand_level kernels)",
                            ["collate_keys", "collate_sort", "collate_maps"], 4 * slots + 8 * nE + 12 * nodes,
                            "not in 8(d); minimum traffic: 4 B per tree slot read + 8 B per sorted key written + 12 B of map per batch node",
                            "collate"))
    if run.sharded and phase("halo_stage"):
        blocks.append(hbm_block("halo_stage_kernel (remote-neighbour feature halo: one row per unique batch node, NVLink for the remote ones)",
                                ["halo_stage"], nodes * (4 * F + 4) + nodes * 4 * F,
                                "one row read (local HBM or NVLink) + one row written per unique batch node; bound is NVLink, HBM peak shown",
                                "halo_stage"))
    return blocks


def measure(env: Env, args, wl_name, features, K, W, want_e2e, want_cpu, want_full, tag):
    fan = fanout_of(args, WORKLOADS[wl_name])
    hot = args.hot_rows if args.hot_rows is not None else float(os.environ.get("GIGL_HOT_ROWS", "0"))
    run = Run(env, wl_name, features, args.halo, args.batch, fan, hot, tag, streams=args.streams)
    wl, B, world = run.wl, run.B, env.world
    one = run.measure_device(K, W, streams=1)  # one batch at a time: the per-phase device times (rooflines) come from this form
    dev_res = run.measure_device(K, W, streams=args.streams) if args.streams > 1 else one
    ms = dev_res["ms"]
    value = world * B * K / (ms * 1e-3)
    cnt = run.counts(min(K, 8), W)
    agg_edges = (cnt["e1"] + cnt["e2"])
    blocks = roofline_blocks(run, one, cnt, K)
    timed = [b for b in blocks if b.get("ms_per_launch") and not b["kernel"].startswith("layer-1 gather-SpMM as a whole")]
    head = max(timed, key=lambda b: b["ms_per_launch"]) if timed else None
    phase_ms = {k: v[0] / K for k, v in one["timings"].items()}
    rec = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
           "ms_per_step_min": float(np.min(dev_res["per_step"])), "ms_per_step_median": float(np.median(dev_res["per_step"])),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(wl_name, wl, fan, B, world), "residency": run.residency(),
           "aggregated_edges_per_sec": world * agg_edges / (ms / K * 1e-3),
           "aggregate_only_edges_per_sec": agg_edges / max(1e-9, sum(phase_ms.get(k, 0.0) for k in
                                                                     ("gather_l1", "gather_deep", "gemm_l1", "gemm_deep", "halo_stage")) * 1e-3),
           "sample_only_subgraphs_per_sec": B / max(1e-9, phase_ms.get("sample", 0.0) * 1e-3),
           "phase_ms_per_step": phase_ms, "unique_edges_per_step": cnt["e1"], "layer1_rows_per_step": cnt["n1"],
           "batch_nodes_per_step": cnt["nodes"], "sampled_edges_per_step": cnt["sampled"],
           "batches_in_flight": dev_res["streams"],
           "one_batch_at_a_time": {"ms_per_step": one["ms"] / K, "value": world * B * K / (one["ms"] * 1e-3),
                                   "ms_per_step_median": float(np.median(one["per_step"]))},
           "roofline": head, "rooflines": blocks, "gpu_launches": dev_res["launches"], "clocks": dev_res["clocks"]}
    if run.sharded and phase_ms.get("halo_stage"):
        F = wl["F"]
        moved = -(-4 * F // 32) * 32  # bytes of a row that cross the link: whole 32-byte sectors of its F floats
        remote = None if run.hot_rows else cnt["nodes"] * (1.0 - 1.0 / world)  # with a hot set the share depends on the batch
        rec["halo"] = {"ms_per_step_on_the_side_stream": phase_ms["halo_stage"], "exposed_wait_ms_per_step": phase_ms.get("halo_wait"),
                       "rows_per_step": cnt["nodes"], "row_bytes_moved": moved, "remote_rows_per_step_if_uniform": cnt["nodes"] * (1.0 - 1.0 / world),
                       "nvlink_GBps_if_uniform": (remote * moved / (phase_ms["halo_stage"] * 1e-3) / 1e9) if remote else None,
                       "hot_rows_replicated": run.hot_rows,
                       "staging": ("forked level by level inside the sampling call" if run.stage_in_sampler else
                                   "forked inside the collation" if os.environ.get("GIGL_HALO_EARLY", "1") != "0" else "inside the forward"),
                       "note": "the copy runs on a side stream beside the sampler / collation kernels, so its own duration is stretched by "
                               "them; what the step pays is the difference of one_batch_at_a_time to the replicated variant"}
    if want_full and not run.sharded:
        rec["full_graph_aggregate"] = run.full_graph()
    if want_e2e:
        rec["e2e"] = run.measure_e2e(K, W, padded=args.e2e_padded, streams=args.streams, ids=args.e2e_ids)
        rec["e2e_embeddings_only"] = run.measure_e2e(K, W, embeddings_only=True, streams=args.streams)
    if want_cpu:
        if env.rank == 0:
            rec["cpu_baseline"] = run.cpu_baseline(max(1, args.cpu_steps))
        env.barrier()
    run.close()
    return rec


def measure_layerwise(env: Env, args, wl_name, K, W, tag):
    """BASELINE.json configs[4], the Inferencer's full-graph embedding export as layer-wise inference: a
    2-layer GraphSAGE over EVERY node, one layer at a time over the whole CSR - rank r computes the rows of its contiguous
    node range, reads its neighbours' rows from the flat sharded table (input features for layer 1, the hidden rows every
    rank just wrote for layer 2: the hidden-row halo, peer loads over NVLink) and writes its rows of the next table.  The
    only collective is the barrier between the layers (a rank may not read hidden rows before their owner wrote them).
    step = one pass over the graph; metric = aggregated edges / s (every CSR edge reduced once per layer)."""
    import torch

    from gigl_b200 import Context, Graph, synth
    from gigl_b200.sharding import ShardedFeatureTable, root_range

    wl, dev, world, rank = WORKLOADS[wl_name], env.dev, env.world, env.rank
    N, F, H, O = wl["nodes"], wl["F"], wl["H"], wl["O"]
    ctx = Context.on_torch_stream(env.local)
    src, dst, _, layers = build_inputs_torch(wl, dev, with_features=False)
    g = Graph.from_edges_dev(ctx, N, src, dst, is_graph_directed=wl["directed"])
    del src, dst
    rowptr, col = g.csr_tensors()
    lo, hi = root_range(N, rank, world)
    rows = hi - lo
    gen = torch.Generator(device=dev).manual_seed(synth.GEN_SEED + 1 + rank)
    tables = []
    if world > 1:
        tx = ShardedFeatureTable(ctx, N, F, rank, world, tag=tag + "x")
        th = ShardedFeatureTable(ctx, N, H, rank, world, tag=tag + "h")
        tables = [tx, th]
        assert (tx.row_lo, tx.row_hi) != (None, None)
        x_flat, h_flat = tx.table[:N], th.table[:N]
        x_mine, h_mine = tx.local, th.local
        own_lo, own_hi = tx.row_lo, tx.row_hi  # the table's shards follow its mapping granule, not root_range
        lo, hi, rows = own_lo, own_hi, own_hi - own_lo
    else:
        x_flat = torch.empty((N, F), dtype=torch.float32, device=dev)
        h_flat = torch.empty((N, H), dtype=torch.float32, device=dev)
        x_mine, h_mine = x_flat, h_flat
    for r0 in range(0, rows, 1 << 22):
        r1 = min(rows, r0 + (1 << 22))
        x_mine[r0:r1].copy_(torch.randn(r1 - r0, F, device=dev, dtype=torch.float32, generator=gen))
    W1 = torch.from_numpy(np.concatenate([layers[0][0], layers[0][2]], axis=1)).to(dev)  # [H, 2F] = [lin_l | lin_r]
    W2 = torch.from_numpy(np.concatenate([layers[1][0], layers[1][2]], axis=1)).to(dev)
    b1, b2 = torch.from_numpy(layers[0][1]).to(dev), torch.from_numpy(layers[1][1]).to(dev)
    out = torch.empty((rows, O), dtype=torch.float32, device=dev)
    A = torch.empty((rows, 2 * max(F, H)), dtype=torch.float32, device=dev)
    my_rowptr = rowptr[lo:hi + 1]
    my_edges = int((rowptr[hi] - rowptr[lo]).item())
    # N > 1: a rank's rows have in-neighbours all over the graph, so gathering straight from the sharded table moves one
    # (mostly remote) row per EDGE over NVLink.  With room in HBM the layer's input table is pulled into a transient local
    # replica first - one row per NODE, the all-gather volume - and the gather runs from local HBM.  GIGL_LAYERWISE=direct
    # keeps the per-edge peer loads (and is what runs when the replica does not fit).
    replica, ag_ms, parts = None, [], []
    if world > 1 and os.environ.get("GIGL_LAYERWISE", "replica") != "direct":  # replica | replica_memcpy | direct
        need = N * max(tx.table.shape[1], th.table.shape[1]) * 4
        torch.cuda.empty_cache()  # the graph build's temporaries
        free, _ = torch.cuda.mem_get_info(dev)
        if need + (12 << 30) < free:
            replica = torch.empty(need // 4, dtype=torch.float32, device=dev)
    ctx.sync()
    env.barrier()

    def layer(table_flat, mine, Fin, Wc, b, relu, dst_rows):
        if replica is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            rep = replica[: N * table_flat.shape[1]].view(N, table_flat.shape[1])
            e0.record()
            # peer reads through the flat mapping: (world - 1) / world of it crosses NVLink
            if os.environ.get("GIGL_LAYERWISE", "replica") == "replica_memcpy":
                rep.copy_(table_flat)   # cudaMemcpy over the peer mapping: measured 108 GB/s, the copy kernel below 540-650
            else:
                # ring order - rank r pulls shards r, r + 1, ... - so every owner serves ONE reader at a time; all ranks walking
                # the table front to back queue up on one GPU's NVLink egress (measured: the step 2x longer at N = 8)
                for k in range(world):
                    s_lo = ((rank + k) % world) * tx.rows_per_shard
                    s_hi = min(N, s_lo + tx.rows_per_shard)
                    if s_hi > s_lo:
                        torch.mul(table_flat[s_lo:s_hi], 1.0, out=rep[s_lo:s_hi])
            e1.record()
            ag_ms.append((e0, e1))
            table_flat = rep
        t0, t1, t2, t3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        t0.record()
        agg = ctx.gather_mean(table_flat, my_rowptr, col, n_rows_out=rows)   # mean of the in-neighbours' rows, local or peer
        t1.record()
        a = A[:, : 2 * Fin]
        a[:, :Fin].copy_(agg)
        a[:, Fin:].copy_(mine[:rows])
        t2.record()
        ctx.linear(a, Wc, b, relu=relu, out=dst_rows)
        t3.record()
        parts.append((t0, t1, t2, t3))

    def step():
        layer(x_flat, x_mine, F, W1, b1, True, h_mine[:rows])
        if world > 1:
            env.dist.barrier()  # every rank's hidden rows are written before anyone gathers them
        layer(h_flat, h_mine, H, W2, b2, False, out)
        if world > 1:
            env.dist.barrier()  # ... and read before the next pass overwrites them

    for _ in range(W):
        step()
    ag_ms.clear()
    parts.clear()
    ctx.set_timing(True)
    ctx.reset_timing()
    l0 = ctx.launch_count
    clocks = ClockSampler(env.local)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    env.barrier()
    clocks.start()
    evs[0].record()
    for i in range(K):
        step()
        evs[i + 1].record()
    env.barrier()
    ms = max_over_ranks(evs[0].elapsed_time(evs[K]), dev)
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
    clk = clocks.stop()
    timings = ctx.timings()
    ctx.set_timing(False)
    launches = ctx.launch_count - l0
    total_edges = int(g.n_edges)
    hbm, _, src_peak = measured_peaks()
    g_ms = timings.get("gather_full", (0.0, 0))[0] / K / 2  # per layer
    torch.cuda.synchronize()
    layer_ms = {"gather": float(np.mean([p[0].elapsed_time(p[1]) for p in parts])),
                "assemble_mean_self": float(np.mean([p[1].elapsed_time(p[2]) for p in parts])),
                "projection": float(np.mean([p[2].elapsed_time(p[3]) for p in parts]))}
    allgather = None
    if replica is not None:
        torch.cuda.synchronize()
        per_layer = float(np.mean([a.elapsed_time(b) for a, b in ag_ms]))
        remote = (N - rows) * F * 4
        allgather = {"ms_per_layer": per_layer, "remote_bytes_per_layer": remote, "nvlink_GBps_in": remote / (per_layer * 1e-3) / 1e9,
                     "replica_bytes": int(replica.numel()) * 4,
                     "note": "the layer's input table copied into a transient local replica (one row per node) before the gather"}
    bytes_layer = my_edges * (4 * F + 4) + (rows + 1) * 8  # SURVEY 8(d) gather terms of bytes_A (the self rows are read by the copy)
    ach = bytes_layer / (g_ms * 1e-3) / 1e9 if g_ms else None
    rec = {"metric": "aggregated-edges/sec", "value": 2.0 * total_edges * K / (ms * 1e-3), "unit": "edges/s", "n_gpus": world, "steps": K,
           "warmup": W, "ms_per_step": ms / K, "ms_per_step_min": float(np.min(per_step)), "ms_per_step_median": float(np.median(per_step)),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"BASELINE.json {wl['cfg']} shape: {wl_name} synthetic RMAT graph, N={N}, {wl['pairs']} directed edges, "
                                  f"layer-wise full-graph GraphSAGE {F}->{H}->{O} (every node's embedding)",
                      "rows_per_gpu": rows, "edges_per_gpu": my_edges,
                      "residency": ("CSR replicated; input and hidden tables sharded by node range and mapped flat; "
                                    + ("each layer's input pulled into a transient local replica over NVLink, then gathered locally"
                                       if replica is not None else "per-edge peer loads over NVLink"))
                                   if world > 1 else "CSR, input, hidden and output tables resident on the one GPU"},
           "allgather": allgather, "ms_per_layer": layer_ms,
           "phase_ms_per_step": {k: v[0] / K for k, v in timings.items()},
           "roofline": {"kernel": "gather_rows_kernel / gather_heavy_kernel (full-graph mean aggregate of one layer)", "bound": "hbm" if world == 1 or replica is not None else "nvlink (remote rows) / hbm",
                        "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm if ach else None, "traffic": None,
                        "algorithmic_bytes_per_launch": bytes_layer, "formula": "8(d) bytes_A gather terms: e (4 F + 4) + (n + 1) 8",
                        "ms_per_launch": g_ms, "peak_source": src_peak},
           "gpu_launches": int(launches), "clocks": clk}
    for t in tables:
        ctx.sync()
        env.barrier()
    g.close()
    for t in tables:
        t.close()
    ctx.close()
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    env = Env()
    if WORKLOADS[args.workload].get("layerwise"):
        rec = measure_layerwise(env, args, args.workload, args.steps, args.warmup, os.environ.get("MASTER_PORT", "0") + "l")
        if env.rank == 0:
            print(json.dumps(rec), flush=True)
        if env.world > 1:
            env.dist.destroy_process_group()
        return
    K, W = args.steps, args.warmup
    features = args.features
    if features == "auto":
        features = "sharded" if env.world > 1 else "replicated"
    port = os.environ.get("MASTER_PORT", "0")
    line = measure(env, args, args.workload, features, K, W, want_e2e=not args.no_e2e,
                   want_cpu=(env.world == 1 and not args.no_cpu_baseline), want_full=not args.no_full_graph, tag=port + "a")
    if env.world > 1 and not args.no_extras and features == "sharded":
        rep = measure(env, args, args.workload, "replicated", K, W, want_e2e=False, want_cpu=False, want_full=False, tag=port + "b")
        line["replicated"] = {k: rep[k] for k in ("value", "ms_per_step", "ms_per_step_median", "phase_ms_per_step", "residency")}
        line["sharded_over_replicated"] = line["value"] / rep["value"]
    if (env.world == 8 and not args.no_extras and args.workload == "products-like") or args.with_g1b:
        wl_name = "g1b"
        g1b = measure(env, args, wl_name, features if env.world > 1 else "replicated", max(5, K // 2), W, want_e2e=not args.no_e2e,
                      want_cpu=False, want_full=False, tag=port + "c")
        line["g1b"] = g1b
    if env.rank == 0:
        print(json.dumps(line), flush=True)
    if env.world > 1:
        env.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
