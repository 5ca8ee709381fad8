#!/usr/bin/env python
"""bench.py - the hot path of BASELINE.json on N B200s of one node.

A "step" = one batch of B roots through the whole path: k-hop rooted-neighbourhood sampling
(fanout [15, 10], GiGL's deterministic hash permutation) -> batch collation -> 2-layer GraphSAGE
over the coalesced batch graph -> root embeddings.  Workload at N = 1 is BASELINE.json configs[1]:
the ogbn-products-shaped synthetic graph (N = 2,449,029, 61.86M undirected pairs mirrored,
F = 100 fp32, 100 -> 256 -> 47), whole CSR + features HBM-resident.  At N > 1 every rank holds a
replica and samples its own contiguous range of roots (weak scaling, no data-path collective).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
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
    # N = 1e8, E = 1e9 directed RMAT, F = 128; every GPU holds the whole CSR (4.8 GB) + features (51 GB)
    "g1b": dict(nodes=100_000_000, pairs=1_000_000_000, F=128, H=128, O=128, directed=True, cfg="configs[3]"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gigl_b200", choices=["gigl_b200", "reference"])
    ap.add_argument("--workload", default="products-like", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=65536, help="roots per step per GPU")
    ap.add_argument("--fanout", default="15,10")
    ap.add_argument("--cpu-sample-roots", type=int, default=8192, help="roots per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-full-graph", action="store_true")
    ap.add_argument("--e2e-padded", action="store_true", help="e2e returns the padded-tree index sets instead of the packed form")
    ap.add_argument("--pitch-features", action="store_true", help="experiment: replicated feature rows pitched to 128-byte multiples")
    ap.add_argument("--shard-features", action="store_true",
                    help="features sharded by node range over the N GPUs and mapped as one flat table (remote rows over NVLink) "
                         "instead of replicated")
    ap.add_argument("--halo", default="staged", choices=["staged", "direct"],
                    help="--shard-features only: 'staged' copies each unique batch node's row into local HBM once per step and "
                         "gathers from the copy; 'direct' loads one (mostly remote) row per unique edge inside the gather kernel")
    return ap.parse_args()


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
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_numa_node(device_index: int):
    """N > 1 only: pins this rank's threads to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer is
    allocated, so the e2e path's host buffers live on the socket the GPU's PCIe link hangs off (with 8 ranks placed by
    the scheduler, half of the device-to-host copies otherwise cross the socket interconnect).  Best effort."""
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
# CPU pipeline = the reference algorithm restated (oracle/): sampler in C + OpenMP, collation in
# numpy, aggregate with torch's own CPU kernels.  Timed as the baseline; never the product path.
# ----------------------------------------------------------------------------------------------
def cpu_step(O, rowptr, col, x, roots, fan, layers, threads):
    nbr, _ = O.c_sample_khop(rowptr, col, roots, fan, n_threads=threads)
    t1 = time.perf_counter()
    node_ids, ei, root_idx = O.np_collate_fast(roots, nbr, fan)
    xb = x(node_ids) if callable(x) else x[node_ids]
    t2 = time.perf_counter()
    out = O.torch_sage_forward(xb, ei, layers, n_threads=threads)[root_idx]
    return out, ei.shape[1], t1, t2


def run_cpu_baseline(rowptr, col, x, fan, layers, n_nodes, n_roots, steps, warmup):
    from oracle import oracle as O

    threads = os.cpu_count() or 1
    batches = root_batches(n_nodes, 0, 1, n_roots, steps + warmup)
    for b in batches[:warmup]:
        cpu_step(O, rowptr, col, x, b, fan, layers, threads)
    t_s = t_c = t_a = 0.0
    edges = 0
    t0 = time.perf_counter()
    for b in batches[warmup:]:
        ta = time.perf_counter()
        _, e, t1, t2 = cpu_step(O, rowptr, col, x, b, fan, layers, threads)
        tb = time.perf_counter()
        t_s += t1 - ta
        t_c += t2 - t1
        t_a += tb - t2
        edges += e * len(layers)
    dt = time.perf_counter() - t0
    return {"value": n_roots * steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} step(s) x {n_roots} roots (ids in order) of the same graph/fanout/model; C+OpenMP sampler, numpy collate, "
                      f"torch-CPU SAGE on the whole batch graph as the reference does",
            "seconds": dt, "phase_seconds": {"sample": t_s, "collate": t_c, "aggregate": t_a},
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
def build_inputs_torch(wl, dev):
    """Synthetic graph + features + weights (seeded; identical on every box)."""
    from gigl_b200 import synth

    src, dst = synth.rmat_edges_torch(wl["nodes"], wl["pairs"], dev)
    x = synth.features_torch(wl["nodes"], wl["F"], dev)
    layers = synth.sage_weights(np.random.default_rng(synth.GEN_SEED), [wl["F"], wl["H"], wl["O"]])
    return src, dst, x, layers


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; Spark / PyG are not runnable
    here) on the host cores, same workload / metric."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    wl = WORKLOADS[args.workload]
    fan = [int(v) for v in args.fanout.split(",")]
    dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    src, dst, x, layers = build_inputs_torch(wl, dev)  # input preparation only (torch ops, not timed)
    rowptr, col = O.torch_build_in_csr(src, dst, wl["nodes"], wl["directed"])
    rowptr, col, x = rowptr.cpu().numpy(), col.cpu().numpy(), x.cpu().numpy()
    del src, dst
    n_roots = args.cpu_sample_roots
    res = run_cpu_baseline(rowptr, col, x, fan, layers, wl["nodes"], n_roots, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["seconds"] / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, wl, fan, n_roots, "each step is a bounded sample of the workload"),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "phase_seconds": res["phase_seconds"], "aggregated_edges_per_s": res["aggregated_edges_per_s"],
            "cpu_model": cpu_model()}
    print(json.dumps(line), flush=True)


def workload_config(args, wl, fan, batch, note):
    residency = ("whole CSR + feature table resident in HBM on every GPU (replicated); roots sharded by contiguous id range"
                 if not getattr(args, "shard_features", False) else
                 "CSR replicated; feature table sharded by contiguous node range over the GPUs and mapped as one flat array "
                 "(cuMemMap of peer shards, rows pitched to 128-byte multiples): " +
                 ("the row of every unique batch node is copied over NVLink into a per-batch table once per step (halo staging), "
                  "layer 1 gathers from the copy" if getattr(args, "halo", "staged") == "staged" else
                  "remote neighbour rows are loaded over NVLink inside the gather kernel, one per unique edge"))
    edges = (f"{wl['pairs']} directed edges (duplicates kept)" if wl["directed"] else
             f"{wl['pairs']} undirected pairs de-duplicated+mirrored")
    return {"workload": f"BASELINE.json {wl['cfg']} shape: {args.workload} synthetic RMAT(0.57,0.19,0.19,0.05) graph, "
                        f"N={wl['nodes']}, {edges}, F={wl['F']} fp32, "
                        f"2-hop fanout {fan}, GraphSAGE {wl['F']}->{wl['H']}->{wl['O']} inference on the coalesced batch graph",
            "roots_per_step_per_gpu": batch, "fanout": fan, "seed": {"generator": 20260101, "sampler_base_seed": 42, "first_call_no": 1},
            "residency": residency,
            "l2": f"every step reads different roots from a {wl['nodes'] * wl['F'] * 4 / 1e9:.2f} GB feature table + the CSR "
                  "(>> 126 MB L2); no flush needed",
            "note": note}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from gigl_b200 import Batch, Context, Graph, SageModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else None  # before CUDA / pinned allocations
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":  # NCCL's banner must not land on stdout beside the JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    fan = [int(v) for v in args.fanout.split(",")]
    B = min(args.batch, wl["nodes"] // world)
    K, W = args.steps, args.warmup

    ctx = Context.on_torch_stream(local)
    src, dst, x, layers = build_inputs_torch(wl, dev)
    g = Graph.from_edges_dev(ctx, wl["nodes"], src, dst, is_graph_directed=wl["directed"])
    del src, dst
    table = None
    if args.shard_features:
        from gigl_b200.sharding import ShardedFeatureTable

        # rows pitched to whole 128-byte lines: remote rows cross NVLink as full lines (2x the link efficiency at F = 100)
        pitch = -(-wl["F"] // 32) * 32
        table = ShardedFeatureTable(ctx, wl["nodes"], pitch, rank, world, tag=os.environ.get("MASTER_PORT", "0"))
        table.local[: table.row_hi - table.row_lo, : wl["F"]].copy_(x[table.row_lo:table.row_hi])  # this rank keeps only its rows
        del x
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        x = table.table[: wl["nodes"], : wl["F"]]
    if args.pitch_features and not args.shard_features:
        xp = torch.zeros((wl["nodes"], -(-wl["F"] // 32) * 32), dtype=torch.float32, device=dev)
        xp[:, : wl["F"]] = x
        x = xp[:, : wl["F"]]
    g.set_features(x)
    model = SageModel(ctx, layers)
    batch = Batch(ctx, wl["nodes"])
    if args.shard_features and args.halo == "staged":
        batch.set_halo_staging(True)
    ctx.sync()
    torch.cuda.empty_cache()

    batches = root_batches(wl["nodes"], rank, world, B, K + W)
    roots_dev = [torch.from_numpy(b).to(dev) for b in batches]
    O_dim = wl["O"]
    out = torch.empty((B, O_dim), dtype=torch.float32, device=dev)
    nbr, cnt = g.sample_khop(roots_dev[0], fan)

    def step(i):
        g.sample_khop(roots_dev[i], fan, out=(nbr, cnt))
        batch.collate(roots_dev[i], fan, nbr, 2)
        batch.sage_forward(model, x, out=out)
        return batch.n_edges

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident measurement (value) ---------------------------------------------------
    for i in range(W):
        step(i)
    ctx.set_timing(True)
    ctx.reset_timing()
    l0 = ctx.launch_count
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    e1_total = 0
    for i in range(W, W + K):
        e1_total += step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    launches = ctx.launch_count - l0
    timings = ctx.timings()
    ctx.set_timing(False)
    ms = max_over_ranks(ms, dev)
    value = world * B * K / (ms * 1e-3)

    # edges aggregated per step: layer 1 reduces every unique batch edge, layer 2 those into roots (untimed recount)
    e2_total = 0
    n1_total = 0
    nodes_total = 0
    sampled_total = 0
    for i in range(W, W + K):
        g.sample_khop(roots_dev[i], fan, out=(nbr, cnt))
        sizes = batch.collate(roots_dev[i], fan, nbr, 2)
        n1_total += sizes[1]
        node_ids, ei = batch.export()
        e2_total += int((ei[1] < B).sum().item())
        nodes_total += int(node_ids.numel())
        sampled_total += sum(int((t >= 0).sum().item()) for t in nbr)
    agg_edges = e1_total + e2_total

    # ---- roofline of the dominant kernel: the layer-1 gather ------------------------------------
    peak, peak_src = measured_peaks()
    F = wl["F"]
    g_ms, g_n = timings.get("gather_l1", (0.0, 0))
    # algorithmic bytes per launch: per unique edge one source row + its sorted key, per output row the
    # segment descriptor + node id + self row read + [mean | self] row written
    alg_bytes = (e1_total * (4 * F + 8) + n1_total * (8 + 4 + 4 * F + 8 * F)) / max(K, 1)
    achieved = alg_bytes / (g_ms / max(g_n, 1) * 1e-3) / 1e9 if g_n else None
    phase_ms = {k: v[0] / K for k, v in timings.items()}
    roofline = {"kernel": "batch_gather_async_kernel<1, 8> (layer-1 gather over the coalesced batch graph; + split-row parts/finish launches)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "ms_per_launch": g_ms / max(g_n, 1), "share_of_step": (g_ms / K) / (ms / K)}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and args.workload == "products-like" and B == 65536 and not args.shard_features:
        try:
            roofline["traffic"] = json.load(open(prof)).get("gather_l1_dram_bytes_per_launch")
        except Exception:
            pass

    # ---- the aggregate over the WHOLE graph (every node a row, every CSR edge reduced once): the full-graph form the
    # Trainer's nn modules and layer-wise inference use; reported beside the batch numbers, outside the timed step
    full = None
    if not args.no_full_graph and not args.shard_features:
        rowptr_t, col_t = g.csr_tensors()
        agg = torch.empty((wl["nodes"], F), dtype=torch.float32, device=dev)
        ctx.gather_mean(x, rowptr_t, col_t, out=agg)
        fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        fe0.record()
        for _ in range(5):
            ctx.gather_mean(x, rowptr_t, col_t, out=agg)
        fe1.record()
        barrier()
        f_ms = max_over_ranks(fe0.elapsed_time(fe1) / 5, dev)
        f_bytes = g.n_edges * (4 * F + 4) + (wl["nodes"] + 1) * 8 + wl["nodes"] * 4 * F  # SURVEY 8(d) bytes_A without the projection
        full = {"op": "gigl_gather_mean_dev over the whole CSR (SAGEConv mean aggregate of every node)", "edges": int(g.n_edges),
                "ms": f_ms, "aggregated_edges_per_sec": world * g.n_edges / (f_ms * 1e-3), "algorithmic_GBps_per_gpu": f_bytes / f_ms / 1e6,
                "frac_of_hbm_peak": f_bytes / f_ms / 1e6 / peak}
        del agg

    # ---- end to end through the host entry point (pinned host buffers, copies inside the timed region) ----
    # index sets come back packed (one-byte counts + the filled slots only: same edges, ~2/3 of the bytes; --e2e-padded
    # returns the padded tree instead)
    e2e = None
    if not args.no_e2e:
        roots_pin = [torch.from_numpy(b).pin_memory() for b in batches]
        out_pin = torch.empty((B, O_dim), dtype=torch.float32).pin_memory()
        nbr_pin, cnt_pin, cnt8_pin, width = [], [], [], 1
        for f in fan:
            cnt_pin.append(torch.empty(B * width, dtype=torch.int32).pin_memory())
            cnt8_pin.append(torch.empty(B * width, dtype=torch.uint8).pin_memory())
            width *= f
            nbr_pin.append(torch.empty(B * width, dtype=torch.int32).pin_memory())
        d2h_steps = []
        if args.e2e_padded:
            s_out = ([t.numpy() for t in nbr_pin], [t.numpy() for t in cnt_pin])

            def host_step(i):
                g.infer_khop_sage_host(batch, model, roots_pin[i].numpy(), fan, return_samples=True, out=out_pin.numpy(), samples_out=s_out)
                d2h_steps.append(B * O_dim * 4 + sum(t.numel() * 4 for t in nbr_pin + cnt_pin))
        else:
            packed_pin = torch.empty(sum(t.numel() for t in nbr_pin), dtype=torch.int32).pin_memory()
            p_out = (packed_pin.numpy(), [t.numpy() for t in cnt8_pin])

            def host_step(i):
                _, packed, _ = g.infer_khop_sage_packed_host(batch, model, roots_pin[i].numpy(), fan, out=out_pin.numpy(), packed_out=p_out)
                d2h_steps.append(B * O_dim * 4 + packed.size * 4 + sum(t.numel() for t in cnt8_pin))
        for i in range(W):
            host_step(i)
        d2h_steps.clear()
        barrier()
        ev0.record()
        for i in range(W, W + K):
            host_step(i)
        ev1.record()
        barrier()
        ms_e = ev0.elapsed_time(ev1)  # device time on the launching stream (the host call itself is synchronous)
        ms_e = max_over_ranks(ms_e, dev)
        d2h = int(sum(d2h_steps) / max(len(d2h_steps), 1))
        e2e = {"value": world * B * K / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 4, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_e / K, "host_cpus_bound": numa_cpus,
               "api": ("gigl_infer_khop_sage_host (roots in pinned host memory -> padded-tree index sets + root embeddings back in pinned "
                       "host memory)" if args.e2e_padded else
                       "gigl_infer_khop_sage_packed_host (roots in pinned host memory -> packed index sets [one-byte counts + filled "
                       "slots] + root embeddings back in pinned host memory)")}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rowptr_t, col_t = g.csr_tensors()
        if x.numel() * 4 > (8 << 30):
            # a table this large is not copied to the host: the baseline's feature lookup reads the device copy
            x_host = lambda ids: x[torch.from_numpy(ids).to(dev)].cpu().numpy()  # noqa: E731
        else:
            x_host = x.cpu().numpy()
        cpu = run_cpu_baseline(rowptr_t.cpu().numpy(), col_t.cpu().numpy(), x_host, fan, layers, wl["nodes"],
                               args.cpu_sample_roots, 3, 1)
        if callable(x_host):
            cpu["sample"] += "; batch feature rows fetched from the device-resident table"
        cpu["cpu_model"] = cpu_model()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, wl, fan, B, "value: roots already on the device; e2e: host buffers"),
                "aggregated_edges_per_sec": world * agg_edges / (ms * 1e-3),
                "aggregate_only_edges_per_sec": agg_edges / K / max(1e-9, sum(phase_ms.get(k, 0.0) for k in
                                                                              ("gather_l1", "gather_deep", "gemm_l1", "gemm_deep")) * 1e-3),
                "sample_only_subgraphs_per_sec": B / max(1e-9, phase_ms.get("sample", 0.0) * 1e-3),
                "phase_ms_per_step": phase_ms, "unique_edges_per_step": e1_total / K, "layer1_rows_per_step": n1_total / K,
                "batch_nodes_per_step": nodes_total / K, "sampled_edges_per_step": sampled_total / K,
                "roofline": roofline, "full_graph_aggregate": full, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk}
        print(json.dumps(line), flush=True)
    if table is not None:
        ctx.sync()
        if world > 1:
            dist.barrier()
        g.close()
        table.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
