"""Drop-in for the reference's Subgraph Sampler component on one B200:

    python -m gigl_b200.subgraph_sampler <frozen_task_config.yaml> <job_name> <resource_config.yaml>

the argv of `Main.main` (scala/subgraph_sampler/src/main/scala/Main.scala:12-16; submitted by
python/gigl/src/subgraph_sampler/subgraph_sampler.py:290-300).  It reads the node / edge tf.Example TFRecords named by
`sharedConfig.preprocessedMetadataUri`, keeps the graph as a sorted CSR in HBM, samples every node's k-hop
neighbourhood with the CUDA kernel (GiGL's deterministic hash permutation, `samplingSeed = 42`), hydrates node and edge
features and writes, picking the task as `TaskRunner.runTask` does (libs/TaskRunner.scala:16-82):

  * node-based task  - `RootedNodeNeighborhood` TFRecords to `...unlabeledTfrecordUriPrefix` FIRST, then
    `SupervisedNodeClassificationSample` TFRecords to `...labeledTfrecordUriPrefix`
    (SupervisedNodeClassificationTask.scala:29-124);
  * node-anchor-based link prediction - `RootedNodeNeighborhood` TFRecords of every node to
    `nodeTypeToRandomNegativeTfrecordUriPrefix[dstNodeType]` FIRST, then `NodeAnchorBasedLinkPredictionSample`
    TFRecords (positives = seeded out-edge samples, permutation call no. 3; merged neighbourhood of root and
    positives) to `nodeAnchorBasedLinkPredictionOutput.tfrecordUriPrefix` (NodeAnchorBasedLinkPredictionTask.scala:28-312).

Fanouts come from `numNeighborsToSample` (both hops, as the pure-Spark tasks do) or, when present, from
`subgraphSamplingStrategy` (`globalRandomUniform`, or a linear `messagePassingPaths` chain of INCOMING `randomUniform`
ops over the one edge type - the per-hop fanouts that only the reference's spark35 path can express,
scala_spark35/.../SamplingOpDAG.scala:19-51).

User-defined positive / negative label edges (`positiveEdgeInfo` / `negativeEdgeInfo` of the edge type,
UserDefinedLabelsNodeAnchorBasedLinkPredictionTask.scala:54-581) are sampled from and hydrated against their own tables
(`numUserDefinedPositiveSamples` / `numUserDefinedNegativeSamples`; negatives fill `hard_neg_edges`).

Graphs with several node / edge types go through `subgraphSamplingStrategy.messagePassingPaths` (one SamplingOp DAG per
root node type, INCOMING / OUTGOING ops, ops with several inputs; `gigl_b200.dag`) and emit typed, hydrated
`RootedNodeNeighborhood` TFRecords per node type and the typed `NodeAnchorBasedLinkPredictionSample` TFRecords of the
first supervision edge type, as `GraphDBNodeAnchorBasedLinkPredictionTask.scala:100-495` does.

Scope: local / file:// URIs; `gs://` is not implemented.  The reference's default
permutation strategy is the unseedable Spark shuffle; this implementation always uses the seeded hash permutation (a
valid uniform sample; bit-exact to the reference's `permutation_strategy: deterministic`).
"""
from __future__ import annotations

import os
import re
import sys
import time
from typing import List, Optional

import numpy as np
import yaml

from . import sample_io as sio
from .engine import Context, Graph

SAMPLING_SEED = 42  # SubgraphSamplerTask.samplingSeed (libs/task/SubgraphSamplerTask.scala:8-24)


def _load_yaml(uri: str, root: str) -> dict:
    with open(_resolve(uri, root)) as f:
        return yaml.safe_load(f) or {}


def _resolve(uri: str, root: str) -> str:
    p = uri[len("file://"):] if uri.startswith("file://") else uri
    if "://" in p:
        raise ValueError(f"only local / file:// URIs are supported, got {uri!r}")
    return p if os.path.isabs(p) else os.path.join(root, p)


def _first(d: dict):
    """`keysIterator.next()` of a proto map (the default condensed type, SupervisedNodeClassificationTask.scala:32-37)."""
    k = sorted(d.keys(), key=lambda s: int(s))[0]
    return int(k), d[k]


def fanouts_from_config(sgs: dict) -> List[int]:
    """Per-hop fanouts of `datasetConfig.subgraphSamplerConfig`."""
    strat = sgs.get("subgraphSamplingStrategy") or {}
    if "globalRandomUniform" in strat:
        gru = strat["globalRandomUniform"]
        f = int((gru.get("randomUniformSpec") or {}).get("numNodesToSample", 0))
        hops = int(gru.get("numHops", 0))
        if f < 1 or hops < 1:
            raise ValueError("globalRandomUniform needs numHops >= 1 and randomUniformSpec.numNodesToSample >= 1")
        return [f] * hops
    if "messagePassingPaths" in strat:
        paths = strat["messagePassingPaths"].get("paths") or []
        if len(paths) != 1:
            raise ValueError("exactly one messagePassingPath (one root node type) is supported")
        ops = paths[0].get("samplingOps") or []
        by_input = {}
        for op in ops:
            ins = op.get("inputOpNames") or []
            if len(ins) > 1 or "randomUniform" not in op or op.get("samplingDirection", "INCOMING") != "INCOMING":
                raise ValueError("only linear chains of INCOMING randomUniform sampling ops are supported")
            key = ins[0] if ins else None
            if key in by_input:
                raise ValueError("branching sampling DAGs are not supported")
            by_input[key] = op
        fan, cur = [], None
        while cur in by_input:
            op = by_input[cur]
            fan.append(int(op["randomUniform"].get("numNodesToSample", 0)))
            cur = op.get("opName")
        if len(fan) != len(ops) or not fan or min(fan) < 1:
            raise ValueError("sampling ops do not form one chain from the root with numNodesToSample >= 1")
        return fan
    fanout = int(sgs.get("numNeighborsToSample", 0))
    if fanout < 1:
        raise ValueError("datasetConfig.subgraphSamplerConfig.numNeighborsToSample must be >= 1")
    return [fanout, fanout]  # numHops is deprecated and fixed to 2 in the reference (scala/subgraph_sampler/README.md:39-42)


def _feature_matrix(table: "sio.ExampleTable", keys) -> Optional[np.ndarray]:
    """flatten(array(cast(col as array<float>) ...)) over the feature keys IN METADATA ORDER; scalars become
    1-element arrays (SGSPureSparkV1Task.scala:90-104, 176-193)."""
    cols = [table.column(k, "float32", table.width(k)) for k in (keys or [])]
    return np.concatenate(cols, axis=1) if cols else None


_SHARD = (0, 1)  # (rank, world) of this process, set by run()


def _roots_to_device(roots: np.ndarray, device: int):
    """int32 root ids -> the CUDA tensor the device-side SamplingOp entry points take."""
    import torch

    return torch.from_numpy(np.ascontiguousarray(roots, dtype=np.int32)).to(torch.device("cuda", device))


def _floats_to_device(values: np.ndarray, device: int):
    """float32 values (an edge feature per CSR position) -> the CUDA tensor the weighted SamplingOp entry point takes."""
    import torch

    return torch.from_numpy(np.ascontiguousarray(values, dtype=np.float32)).to(torch.device("cuda", device))


_PART_RE = re.compile(r"^part-(?:r(\d+)-)?\d+\.tfrecord$")


def _prepare_dir(dir_: str) -> None:
    """mode("overwrite") of the reference's writer (TFRecordIO.scala:62): part files of an earlier run under this prefix
    are removed before anything is written, so a re-run with another world size / batch size / sample limit leaves no
    stale samples for downstream globs.  No collective: every stale file has exactly one remover - rank r its own
    `part-r<r>-*`, rank 0 the un-ranked files and those of ranks >= world - and each rank removes before it writes."""
    rank, world = _SHARD
    os.makedirs(dir_, exist_ok=True)
    for name in os.listdir(dir_):
        m = _PART_RE.match(name)
        if not m:
            continue
        owner = None if m.group(1) is None else int(m.group(1))
        if world == 1:
            mine = True
        elif owner is None or owner >= world:
            mine = rank == 0
        else:
            mine = owner == rank
        if mine:
            try:
                os.remove(os.path.join(dir_, name))
            except FileNotFoundError:
                pass


def _write(dir_: str, part: int, data, kind: str = "rnn", edge_node_types=None) -> None:
    """data: bytes, or the encoder's NativeBuffer (written straight from the native memory, then released).  Before anything
    reaches the file the samples are held to the reference's TaskOutputValidator (TaskOutputValidator.scala:29-108, run by
    every task before `writeDatasetToTfrecord`): a failure raises and the part is not written."""
    sio.validate_samples(data, kind, edge_node_types)
    rank, world = _SHARD
    name = f"part-{part:05d}.tfrecord" if world == 1 else f"part-r{rank:03d}-{part:05d}.tfrecord"
    with open(os.path.join(dir_, name), "wb") as f:
        if isinstance(data, sio.NativeBuffer):
            f.write(data.view)
            data.close()
        else:
            f.write(data)


def _edge_node_types(hyd: dict) -> dict:
    """Homogeneous graph: the validator's map condensed edge type -> (src, dst) condensed node types."""
    nt = max(int(hyd["condensed_node_type"]), 0)
    return {max(int(hyd["condensed_edge_type"]), 0): (nt, nt)}


def _my_share(ids: np.ndarray) -> np.ndarray:
    """This rank's contiguous slice of the (sorted) root ids: the units are independent, so ranks share nothing and no
    collective is needed (the rule of gigl_b200.sharding.root_range / distributed_neighborloader.py:195-216)."""
    rank, world = _SHARD
    if world == 1:
        return ids
    lo, hi = (len(ids) * rank) // world, (len(ids) * (rank + 1)) // world
    return ids[lo:hi]


def _my_quota(limit: int) -> int:
    """numMaxTrainingSamplesToOutput split over the ranks (the reference keeps an arbitrary subset of that size)."""
    rank, world = _SHARD
    if limit <= 0 or world == 1:
        return limit
    return limit // world + (1 if rank < limit % world else 0)


def run(task_config_uri: str, job_name: str, resource_config_uri: Optional[str] = None, root: Optional[str] = None,
        device: Optional[int] = None, batch_roots: int = 1 << 20, log=print, rank: Optional[int] = None,
        world: Optional[int] = None) -> dict:
    """One process per GPU.  Launched under torchrun (`python -m torch.distributed.run --nproc-per-node N -m
    gigl_b200.subgraph_sampler ...`) every rank loads the graph onto its own GPU (LOCAL_RANK), samples a contiguous share
    of the roots and writes its own `part-r<rank>-*.tfrecord` files into the same output prefixes; rank / world default to
    RANK / WORLD_SIZE of the environment (0 / 1 outside torchrun)."""
    global _SHARD
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    _SHARD = (rank, world)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    root = root or os.getcwd()
    t0 = time.time()
    cfg = _load_yaml(task_config_uri, root)
    shared = cfg.get("sharedConfig", {})
    meta = _load_yaml(shared["preprocessedMetadataUri"], root)
    sgs = cfg.get("datasetConfig", {}).get("subgraphSamplerConfig", {})
    directed = bool(shared.get("isGraphDirected", False))
    skip_main = bool(shared.get("shouldSkipTraining", False)) and bool(shared.get("shouldSkipModelEvaluation", False))
    max_train = _my_quota(int(sgs.get("numMaxTrainingSamplesToOutput", 0) or 0))
    task_meta = cfg.get("taskMetadata", {})
    is_nablp = "nodeAnchorBasedLinkPredictionTaskMetadata" in task_meta
    flat = shared["flattenedGraphMetadata"]

    if len(meta["condensedNodeTypeToPreprocessedMetadata"]) > 1 or len(meta["condensedEdgeTypeToPreprocessedMetadata"]) > 1:
        return _run_typed(cfg, meta, sgs, root, device, batch_roots, job_name, log, t0)
    fanouts = fanouts_from_config(sgs)
    ntype, nmeta = _first(meta["condensedNodeTypeToPreprocessedMetadata"])
    etype, emeta = _first(meta["condensedEdgeTypeToPreprocessedMetadata"])
    # ---- node table: ids, features in featureKeys order, labels
    nodes = sio.ExampleTable.from_files(sio.list_tfrecord_files(_resolve(nmeta["tfrecordUriPrefix"], root)))
    node_id = nodes.column(nmeta["nodeIdKey"], "int64").astype(np.int64)
    feat = _feature_matrix(nodes, nmeta.get("featureKeys"))
    # ---- edge table (+ features)
    main = emeta["mainEdgeInfo"]
    edges = sio.ExampleTable.from_files(sio.list_tfrecord_files(_resolve(main["tfrecordUriPrefix"], root)))
    src = edges.column(emeta["srcNodeIdKey"], "int64")
    dst = edges.column(emeta["dstNodeIdKey"], "int64")
    ef = _feature_matrix(edges, main.get("featureKeys"))
    n_nodes = int(max(node_id.max(initial=-1), src.max(initial=-1), dst.max(initial=-1)) + 1)
    # user-defined label edges (UserDefinedLabelsNodeAnchorBasedLinkPredictionTask): their own tf.Example tables
    label_tables = {}
    if is_nablp:
        for usage, key in (("pos", "positiveEdgeInfo"), ("neg", "negativeEdgeInfo")):
            info = emeta.get(key)
            if info and info.get("tfrecordUriPrefix"):
                t = sio.ExampleTable.from_files(sio.list_tfrecord_files(_resolve(info["tfrecordUriPrefix"], root)))
                ls, ld = t.column(emeta["srcNodeIdKey"], "int64"), t.column(emeta["dstNodeIdKey"], "int64")
                label_tables[usage] = (ls.astype(np.int32), ld.astype(np.int32), _feature_matrix(t, info.get("featureKeys")))
                n_nodes = int(max(n_nodes, ls.max(initial=-1) + 1, ld.max(initial=-1) + 1))
    x = None
    if feat is not None:
        x = np.zeros((n_nodes, feat.shape[1]), dtype=np.float32)
        x[node_id] = feat
    log(f"[{job_name}] loaded {len(node_id)} nodes (F={0 if x is None else x.shape[1]}), {len(src)} edges "
        f"(Fe={0 if ef is None else ef.shape[1]}) in {time.time() - t0:.2f}s; fanouts {fanouts}")

    ctx = Context(device)
    src32, dst32 = src.astype(np.int32), dst.astype(np.int32)
    g = Graph.from_edges_host(ctx, n_nodes, src32, dst32, is_graph_directed=directed)
    # the hydration join needs the CSR on the host when edges carry features or (directed) may carry duplicate records
    csr = g.csr_host() if (ef is not None or directed) else None
    edge_rows = ctx.edge_rows_host(n_nodes, src32, dst32, directed) if ef is not None else None
    hyd = dict(condensed_node_type=ntype, condensed_edge_type=etype, csr=csr, edge_rows=edge_rows, edge_feat=ef)
    roots_all = _my_share(np.sort(node_id).astype(np.int32))  # every node of the node table gets exactly one RootedNodeNeighborhood
    stats = {"rank": rank, "world": world, "n_nodes": n_nodes, "n_edges_csr": g.n_edges, "rnn": 0, "snc": 0, "nablp": 0, "fanouts": fanouts}
    t1 = time.time()
    if is_nablp:
        _run_nablp(g, ctx, cfg, flat, root, roots_all, fanouts, x, hyd, src32, dst32, n_nodes, directed, sgs, skip_main, max_train,
                   batch_roots, stats, label_tables)
    else:
        _run_snc(g, flat, root, roots_all, fanouts, x, hyd, nodes, node_id, nmeta, n_nodes, skip_main, max_train, batch_roots, stats)
    stats["seconds_sample_and_write"] = time.time() - t1
    stats["seconds_total"] = time.time() - t0
    log(f"[{job_name}] wrote {stats['rnn']} RootedNodeNeighborhood + {stats['snc']} SupervisedNodeClassificationSample + "
        f"{stats['nablp']} NodeAnchorBasedLinkPredictionSample records in {stats['seconds_sample_and_write']:.2f}s")
    return stats


def _run_snc(g, flat, root, roots_all, fanouts, x, hyd, nodes, node_id, nmeta, n_nodes, skip_main, max_train, batch_roots, stats):
    out = flat["supervisedNodeClassificationOutput"]
    ent = _edge_node_types(hyd)
    labels = None
    label_key = (nmeta.get("labelKeys") or [None])[0]
    if label_key and not skip_main:
        labels = np.full(n_nodes, sio.INT32_MIN, dtype=np.int32)
        labels[node_id] = nodes.column(label_key, "int64").astype(np.int32)
    unl_dir = _resolve(out["unlabeledTfrecordUriPrefix"], root)
    lab_dir = _resolve(out["labeledTfrecordUriPrefix"], root)
    _prepare_dir(unl_dir)
    if labels is not None:
        _prepare_dir(lab_dir)
    for part, s in enumerate(range(0, len(roots_all), batch_roots)):
        roots = roots_all[s:s + batch_roots]
        nbr, cnt = g.sample_khop_host(roots, fanouts, base_seed=SAMPLING_SEED, first_call_no=1)
        data, offs = sio.encode_samples(roots, fanouts, nbr, x, kind="rnn", zero_copy=True, **hyd)
        _write(unl_dir, part, data, "rnn", ent)  # RootedNodeNeighborhood first
        stats["rnn"] += len(roots)
        if labels is not None:
            # isolated nodes (no sampled in-edge) are NOT training samples (includeIsolatedNodesInTrainingSamples = false, :43-44)
            lab = labels.copy()
            lab[roots[cnt[0] == 0]] = sio.INT32_MIN
            if max_train > 0:
                keep = roots[(cnt[0] > 0) & (labels[roots] != sio.INT32_MIN)][max(0, max_train - stats["snc"]):]
                lab[keep] = sio.INT32_MIN
            data, offs = sio.encode_samples(roots, fanouts, nbr, x, kind="snc", labels=lab, label_type=label_key, zero_copy=True, **hyd)
            _write(lab_dir, part, data, "snc", ent)
            stats["snc"] += int((np.diff(offs) > 0).sum())


def _label_table(ctx, n_nodes, ls, ld, feat):
    """A user-defined label table (never bidirectionalised: only MAIN edges are, SGSPureSparkV1Task.scala:262-273):
    the out-CSR the labels are sampled from + the host table their edges are hydrated against."""
    g_out = Graph.from_edges_host(ctx, n_nodes, ls, ld, is_graph_directed=True, by_source=True)
    g_in = Graph.from_edges_host(ctx, n_nodes, ls, ld, is_graph_directed=True)
    tab = sio.HostEdgeTable(g_in.csr_host(), ctx.edge_rows_host(n_nodes, ls, ld, True) if feat is not None else None, feat)
    g_in.close()
    return g_out, tab


def _run_nablp(g, ctx, cfg, flat, root, roots_all, fanouts, x, hyd, src32, dst32, n_nodes, directed, sgs, skip_main, max_train,
               batch_roots, stats, label_tables):
    out = flat["nodeAnchorBasedLinkPredictionOutput"]
    ent = _edge_node_types(hyd)
    sup = cfg["taskMetadata"]["nodeAnchorBasedLinkPredictionTaskMetadata"]["supervisionEdgeTypes"]
    dst_type = sup[0]["dstNodeType"]
    neg_map = out.get("nodeTypeToRandomNegativeTfrecordUriPrefix") or {}
    if dst_type not in neg_map:
        raise KeyError(f"nodeTypeToRandomNegativeTfrecordUriPrefix is missing the dstNodeType {dst_type!r} of the first supervision "
                       "edge type")  # the reference throws here too (NodeAnchorBasedLinkPredictionTask.scala:101-108)
    rnn_dir = _resolve(neg_map[dst_type], root)
    main_dir = _resolve(out["tfrecordUriPrefix"], root)
    _prepare_dir(rnn_dir)
    g_pos = g_neg = pos_tab = neg_tab = None
    num_pos = num_neg = 0
    main_tab = sio.HostEdgeTable(hyd["csr"], hyd["edge_rows"], hyd["edge_feat"])
    if not skip_main:
        _prepare_dir(main_dir)
        if "pos" in label_tables:
            num_pos = int(sgs.get("numUserDefinedPositiveSamples", 0))
            if num_pos < 1:  # the reference asserts this (UserDefinedLabelsNodeAnchorBasedLinkPredictionTask.scala:81-88)
                raise ValueError("numUserDefinedPositiveSamples must be > 0 when user-defined positive edges are provided")
            g_pos, pos_tab = _label_table(ctx, n_nodes, *label_tables["pos"])
        else:
            num_pos = int(sgs.get("numPositiveSamples", 0))
            if num_pos < 1:
                raise ValueError("datasetConfig.subgraphSamplerConfig.numPositiveSamples must be >= 1")
            # positives walk the out-CSR (row u = sorted destinations of u); undirected graphs are symmetric
            g_pos = g if not directed else Graph.from_edges_host(ctx, n_nodes, src32, dst32, is_graph_directed=True, by_source=True)
        if "neg" in label_tables:
            num_neg = int(sgs.get("numUserDefinedNegativeSamples", 0))
            if num_neg < 1:
                raise ValueError("numUserDefinedNegativeSamples must be > 0 when user-defined negative edges are provided")
            g_neg, neg_tab = _label_table(ctx, n_nodes, *label_tables["neg"])
    for part, s in enumerate(range(0, len(roots_all), batch_roots)):
        roots = roots_all[s:s + batch_roots]
        n = len(roots)
        pos = neg = None
        sample_roots = roots
        if g_pos is not None:
            # permutation call numbers: hop 1 = 1, hop 2 = 2, positives = 3, user-defined negatives = 4 (the JVM-global
            # counter of SamplingStrategy.scala:14,36,79 in the order the task calls it)
            pos, pcnt = g_pos.sample_positives_host(roots, num_pos, base_seed=SAMPLING_SEED, call_no=3)
            pos = pos.reshape(n, num_pos)
            if max_train > 0:
                # numMaxTrainingSamplesToOutput: the reference keeps an arbitrary `LIMIT n` of the anchors
                # (downsampleNumberOfNodes, SGSPureSparkV1Task.scala:1042-1081); this keeps the first n by node id
                anchors = np.flatnonzero(pcnt > 0)
                pos[anchors[max(0, max_train - stats["nablp"]):]] = -1
            labels = pos[pos >= 0]
            if g_neg is not None:
                neg, _ = g_neg.sample_positives_host(roots, num_neg, base_seed=SAMPLING_SEED, call_no=4)
                neg = neg.reshape(n, num_neg)
                labels = np.concatenate([labels, neg[neg >= 0]])
            extra = np.setdiff1d(labels, roots)  # label nodes whose trees are not in this batch
            sample_roots = np.concatenate([roots, extra.astype(np.int32)])
        nbr, cnt = g.sample_khop_host(sample_roots, fanouts, base_seed=SAMPLING_SEED, first_call_no=1)
        width, own = 1, []
        for f in fanouts:
            width *= f
            own.append(nbr[len(own)][: n * width])
        data, _ = sio.encode_samples(roots, fanouts, own, x, kind="rnn", zero_copy=True, **hyd)
        _write(rnn_dir, part, data, "rnn", ent)  # RootedNodeNeighborhood (random negatives) first
        stats["rnn"] += n
        if pos is not None:
            order = np.argsort(sample_roots, kind="stable")
            sorted_roots = sample_roots[order]

            def tree_of(ids):
                where = order[np.searchsorted(sorted_roots, np.where(ids >= 0, ids, sorted_roots[0]))]
                return np.where(ids >= 0, where, -1).astype(np.int64)

            data, offs = sio.encode_link_samples(sample_roots, fanouts, nbr, x, n, pos, tree_of(pos), main_tab, pos_tab, neg,
                                                 tree_of(neg) if neg is not None else None, neg_tab,
                                                 condensed_node_type=hyd["condensed_node_type"], condensed_edge_type=hyd["condensed_edge_type"],
                                                 zero_copy=True)
            _write(main_dir, part, data, "nablp", ent)
            stats["nablp"] += int((np.diff(offs) > 0).sum())


def _run_typed(cfg, meta, sgs, root, device, batch_roots, job_name, log, t0) -> dict:
    """Several node / edge types: one SamplingOp DAG per root node type (`subgraphSamplingStrategy.messagePassingPaths`, the
    only strategy the reference's typed path accepts - SubgraphSamplingStrategyWrapper.scala:10-22), one kernel launch per
    op instance over that edge type's CSR.  Outputs as GraphDBNodeAnchorBasedLinkPredictionTask.run writes them
    (scala_spark35/.../libs/task/graphdb/GraphDBNodeAnchorBasedLinkPredictionTask.scala:100-495): typed, hydrated
    RootedNodeNeighborhood TFRecords per anchor / target node type FIRST, then - for a link-prediction task that trains or
    evaluates - NodeAnchorBasedLinkPredictionSample TFRecords for the first supervision edge type: positives = an OUTGOING
    uniform sample of `numPositiveSamples` over that edge type, neighbourhood = the anchor's merged with its positives'."""
    from . import dag

    shared = cfg["sharedConfig"]
    gm = cfg["graphMetadata"]
    node_type_of = {int(k): v for k, v in gm["condensedNodeTypeMap"].items()}
    cnt_of = {v: k for k, v in node_type_of.items()}
    cet_of = {(e["srcNodeType"], e["relation"], e["dstNodeType"]): int(k) for k, e in gm["condensedEdgeTypeMap"].items()}
    strat = (sgs.get("subgraphSamplingStrategy") or {}).get("messagePassingPaths")
    if not strat:
        raise ValueError("graphs with several node / edge types need subgraphSamplingStrategy.messagePassingPaths")
    flat = shared["flattenedGraphMetadata"]
    task_meta = cfg.get("taskMetadata", {})
    # what the reference's config validation rejects before any component runs (subgraph_sampling_strategy.py:160-283)
    dag.validate_strategy(strat.get("paths") or [], gm, task_meta)
    sup_et = None
    if "nodeAnchorBasedLinkPredictionOutput" in flat:
        out_dirs = dict(flat["nodeAnchorBasedLinkPredictionOutput"].get("nodeTypeToRandomNegativeTfrecordUriPrefix") or {})
        sup = (task_meta.get("nodeAnchorBasedLinkPredictionTaskMetadata") or {}).get("supervisionEdgeTypes") or []
        if sup:
            sup_et = (sup[0]["srcNodeType"], sup[0]["relation"], sup[0]["dstNodeType"])  # phase 1: one supervision edge type (:139)
    else:
        sup = task_meta["nodeBasedTaskMetadata"]["supervisionNodeTypes"]
        out_dirs = {sup[0]: flat["supervisedNodeClassificationOutput"]["unlabeledTfrecordUriPrefix"]}
    skip_main = bool(shared.get("shouldSkipTraining", False)) and bool(shared.get("shouldSkipModelEvaluation", False))
    include_isolated = bool(shared.get("shouldIncludeIsolatedNodesInTraining", False))
    max_train = _my_quota(int(sgs.get("numMaxTrainingSamplesToOutput", 0) or 0))
    # ---- node tables per condensed node type
    ids, tables, n_max = {}, [None] * (max(node_type_of) + 1), 0
    for k, nmeta in meta["condensedNodeTypeToPreprocessedMetadata"].items():
        t = sio.ExampleTable.from_files(sio.list_tfrecord_files(_resolve(nmeta["tfrecordUriPrefix"], root)))
        nid = t.column(nmeta["nodeIdKey"], "int64").astype(np.int64)
        feat = _feature_matrix(t, nmeta.get("featureKeys"))
        ids[int(k)] = np.sort(nid).astype(np.int32)
        if feat is not None:
            x = np.zeros((int(nid.max(initial=-1)) + 1, feat.shape[1]), dtype=np.float32)
            x[nid] = feat
            tables[int(k)] = x
        n_max = max(n_max, int(nid.max(initial=-1)) + 1)
    # ---- edge tables per condensed edge type (+ features)
    edges, edge_feat, edge_feat_cols = {}, {}, {}
    for k, emeta in meta["condensedEdgeTypeToPreprocessedMetadata"].items():
        main = emeta["mainEdgeInfo"]
        t = sio.ExampleTable.from_files(sio.list_tfrecord_files(_resolve(main["tfrecordUriPrefix"], root)))
        es, ed = t.column(emeta["srcNodeIdKey"], "int64"), t.column(emeta["dstNodeIdKey"], "int64")
        edges[int(k)] = (es.astype(np.int32), ed.astype(np.int32))
        edge_feat[int(k)] = _feature_matrix(t, main.get("featureKeys"))
        off, cols = 0, {}
        for fk in main.get("featureKeys") or []:  # feature name -> (first column, width) in the matrix above
            cols[fk] = (off, t.width(fk))
            off += t.width(fk)
        edge_feat_cols[int(k)] = cols
        n_max = max(n_max, int(es.max(initial=-1)) + 1, int(ed.max(initial=-1)) + 1)
    for tt, x in enumerate(tables):  # ids above a type's own table (seen only as edge endpoints) hydrate as zeros
        if x is not None and x.shape[0] < n_max:
            tables[tt] = np.concatenate([x, np.zeros((n_max - x.shape[0], x.shape[1]), np.float32)])
    log(f"[{job_name}] loaded {len(ids)} node types, {len(edges)} edge types in {time.time() - t0:.2f}s")
    ctx = Context.on_torch_stream(device)
    graphs = {}

    def graph_of(edge_type, direction):
        key = (edge_type, direction)  # CSR per (edge type, direction), built on first use
        if key not in graphs:
            es, ed = edges[cet_of[edge_type]]
            graphs[key] = Graph.from_edges_host(ctx, n_max, es, ed, is_graph_directed=True, by_source=direction == dag.OUTGOING)
        return graphs[key]

    edge_tabs = None

    def edge_tables():
        """Host copies of every edge type's records as the hydration join reads them (built once, on first need)."""
        nonlocal edge_tabs
        if edge_tabs is None:
            edge_tabs = [None] * (max(edges) + 1)
            for k, (es, ed) in edges.items():
                g_in = Graph.from_edges_host(ctx, n_max, es, ed, is_graph_directed=True)
                ef = edge_feat[k]
                edge_tabs[k] = sio.HostEdgeTable(g_in.csr_host(), ctx.edge_rows_host(n_max, es, ed, True) if ef is not None else None, ef)
                g_in.close()
        return edge_tabs

    op_weights = {}

    def weights_of(edge_type, direction, name):
        """The edge feature `name` of every edge of `edge_type`, laid out by CSR position of graph_of(edge_type, direction):
        what a TopK / RandomWeighted op orders by (NebulaQueryResponseTranslator.scala:73-105)."""
        key = (edge_type, direction, name)
        if key not in op_weights:
            k = cet_of[edge_type]
            if edge_feat[k] is None or name not in edge_feat_cols[k]:
                raise ValueError(f"sampling op orders {edge_type} by {name!r}, which is not among the edge type's featureKeys")
            col0, width = edge_feat_cols[k][name]
            if width != 1:
                raise ValueError(f"edge feature {name!r} of {edge_type} has {width} values per edge; a weighted sampling op needs a scalar")
            es, ed = edges[k]
            rows = ctx.edge_rows_host(n_max, es, ed, True) if direction == dag.INCOMING else ctx.edge_rows_host(n_max, ed, es, True)
            op_weights[key] = _floats_to_device(edge_feat[k][rows, col0], device)
        return op_weights[key]

    dags = {}
    for path in strat.get("paths") or []:
        ops = dag.ops_from_config(path)
        planned = dag.plan(ops, path["rootNodeType"])
        for p in planned:
            graph_of(p.op.edge_type, p.op.sampling_direction)
            if p.op.sampling_method != "random_uniform":
                weights_of(p.op.edge_type, p.op.sampling_direction, p.op.edge_feat_name)
        # hydrateRnn joins the edges only if an edge type of the DAG's ROOT ops carries features (:186-193, 283-291)
        hydrate = any(edge_feat[cet_of[p.op.edge_type]] is not None for p in planned if p.parent is None)
        dags[path["rootNodeType"]] = (ops, planned, hydrate)

    def sample(rtype, roots):
        ops, planned, _ = dags[rtype]
        res = dag.sample_dag(graphs, _roots_to_device(roots, device), ops, rtype, base_seed=SAMPLING_SEED, weights=op_weights)
        ctx.sync()
        return dag.encoder_ops(planned, res, cet_of, cnt_of)

    stats = {"rnn": 0, "snc": 0, "nablp": 0, "rnn_per_node_type": {}, "n_nodes": n_max}
    ent = {c: (cnt_of[t[0]], cnt_of[t[2]]) for t, c in cet_of.items()}  # the validator's edge type -> endpoint node types
    t1 = time.time()
    for rtype in dags:
        if rtype not in out_dirs:
            continue
        out_dir = _resolve(out_dirs[rtype], root)
        _prepare_dir(out_dir)
        roots_all = _my_share(ids[cnt_of[rtype]])
        hydrate = dags[rtype][2]
        for part, s in enumerate(range(0, len(roots_all), batch_roots)):
            roots = roots_all[s:s + batch_roots]
            data, _ = sio.encode_typed_samples(roots, cnt_of[rtype], sample(rtype, roots), tables, edge_tables() if hydrate else None,
                                               kind="rnn", hydrate_edges=hydrate, zero_copy=True)
            _write(out_dir, part, data, "rnn", ent)
            stats["rnn"] += len(roots)
            stats["rnn_per_node_type"][rtype] = stats["rnn_per_node_type"].get(rtype, 0) + len(roots)
    # ---- main samples of the link-prediction task
    if sup_et is not None and not skip_main and "tfrecordUriPrefix" in flat["nodeAnchorBasedLinkPredictionOutput"]:
        a_type, t_type = sup_et[0], sup_et[2]
        if a_type not in dags or t_type not in dags:
            raise KeyError(f"messagePassingPaths needs a path for the anchor type {a_type!r} and the target type {t_type!r}")
        num_pos = int(sgs.get("numPositiveSamples", 0))
        if num_pos < 1:
            raise ValueError("datasetConfig.subgraphSamplerConfig.numPositiveSamples must be >= 1")
        g_pos = graph_of(sup_et, dag.OUTGOING)
        pos_cet = cet_of[sup_et]
        hyd_graph = dags[a_type][2] or dags[t_type][2]
        hyd_pos = edge_feat[pos_cet] is not None  # posEdgeHasEdgeFeatures (:344-346)
        pos_call = len(dags[a_type][0]) + 1  # the positives are drawn after the anchor's own ops (= call 3 after a 2-hop chain)
        main_dir = _resolve(flat["nodeAnchorBasedLinkPredictionOutput"]["tfrecordUriPrefix"], root)
        _prepare_dir(main_dir)
        anchors_all = _my_share(ids[cnt_of[a_type]])
        if max_train > 0:
            # numMaxTrainingSamplesToOutput: the reference draws a random `.sample(fraction)` of the anchors' RNNs before the
            # join with the positives (:232-262); this keeps the first n by node id
            anchors_all = anchors_all[:max_train]
        for part, s in enumerate(range(0, len(anchors_all), batch_roots)):
            roots = anchors_all[s:s + batch_roots]
            pos, _ = g_pos.sample_op(_roots_to_device(roots, device), [num_pos], [], pos_call, SAMPLING_SEED)
            ctx.sync()
            pos = pos.cpu().numpy().reshape(len(roots), num_pos)
            t_roots = np.unique(pos[pos >= 0]).astype(np.int32)
            tree = np.where(pos >= 0, np.searchsorted(t_roots, np.maximum(pos, 0)), -1).astype(np.int64)
            need = hyd_graph or hyd_pos
            data, offs = sio.encode_typed_samples(roots, cnt_of[a_type], sample(a_type, roots), tables, edge_tables() if need else None,
                                                  kind="nablp", pos=pos, pos_tree=tree, pos_condensed_edge_type=pos_cet,
                                                  target_roots=t_roots, target_node_type=cnt_of[t_type],
                                                  target_ops=sample(t_type, t_roots) if len(t_roots) else [],
                                                  include_isolated=include_isolated, hydrate_edges=hyd_graph, hydrate_pos_edges=hyd_pos,
                                                  zero_copy=True)
            _write(main_dir, part, data, "nablp", ent)
            stats["nablp"] += int((np.diff(offs) > 0).sum())
    stats["seconds_sample_and_write"] = time.time() - t1
    stats["seconds_total"] = time.time() - t0
    log(f"[{job_name}] wrote typed RootedNodeNeighborhood records {stats['rnn_per_node_type']} + {stats['nablp']} "
        f"NodeAnchorBasedLinkPredictionSample records in {stats['seconds_sample_and_write']:.2f}s")
    return stats


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        print(__doc__)
        return 2
    run(argv[0], argv[1], argv[2] if len(argv) > 2 else None)
    return 0


if __name__ == "__main__":
    sys.exit(main())
