"""Drop-in for the reference's Subgraph Sampler component on one B200:

    python -m gigl_b200.subgraph_sampler <frozen_task_config.yaml> <job_name> <resource_config.yaml>

the argv of `Main.main` (scala/subgraph_sampler/src/main/scala/Main.scala:12-16; submitted by
python/gigl/src/subgraph_sampler/subgraph_sampler.py:290-300).  It reads the node / edge tf.Example TFRecords named by
`sharedConfig.preprocessedMetadataUri`, keeps the graph as a sorted CSR in HBM, samples every node's 2-hop
neighbourhood with the CUDA kernel (GiGL's deterministic hash permutation, `samplingSeed = 42`), hydrates and writes
`RootedNodeNeighborhood` TFRecords to `...unlabeledTfrecordUriPrefix` FIRST and then
`SupervisedNodeClassificationSample` TFRecords to `...labeledTfrecordUriPrefix`
(SupervisedNodeClassificationTask.scala:29-124).

Scope: homogeneous graphs, node-classification task output, local / file:// URIs, node features;
edge features and the link-prediction sample types are not emitted yet.  The reference's default permutation strategy is
the unseedable Spark shuffle; this implementation always uses the seeded hash permutation (a valid uniform sample;
bit-exact to the reference's `permutation_strategy: deterministic`).
"""
from __future__ import annotations

import os
import sys
import time
from typing import Optional

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


def run(task_config_uri: str, job_name: str, resource_config_uri: Optional[str] = None, root: Optional[str] = None,
        device: int = 0, batch_roots: int = 1 << 20, log=print) -> dict:
    root = root or os.getcwd()
    t0 = time.time()
    cfg = _load_yaml(task_config_uri, root)
    shared = cfg.get("sharedConfig", {})
    meta = _load_yaml(shared["preprocessedMetadataUri"], root)
    sgs = cfg.get("datasetConfig", {}).get("subgraphSamplerConfig", {})
    fanout = int(sgs.get("numNeighborsToSample", 0))
    if fanout < 1:
        raise ValueError("datasetConfig.subgraphSamplerConfig.numNeighborsToSample must be >= 1")
    fanouts = [fanout, fanout]  # numHops is deprecated and fixed to 2 in the reference (scala/subgraph_sampler/README.md:39-42)
    directed = bool(shared.get("isGraphDirected", False))
    skip_labeled = bool(shared.get("shouldSkipTraining", False)) and bool(shared.get("shouldSkipModelEvaluation", False))
    max_train = int(sgs.get("numMaxTrainingSamplesToOutput", 0) or 0)
    out = shared["flattenedGraphMetadata"]["supervisedNodeClassificationOutput"]

    ntype, nmeta = _first(meta["condensedNodeTypeToPreprocessedMetadata"])
    etype, emeta = _first(meta["condensedEdgeTypeToPreprocessedMetadata"])
    # ---- node table: ids, features in featureKeys order (scalars become 1-element arrays, :90-104), labels
    nodes = sio.ExampleTable.from_files(sio.list_tfrecord_files(_resolve(nmeta["tfrecordUriPrefix"], root)))
    node_id = nodes.column(nmeta["nodeIdKey"], "int64").astype(np.int64)
    cols = [nodes.column(k, "float32", nodes.width(k)) for k in (nmeta.get("featureKeys") or [])]
    # ---- edge table
    edges = sio.ExampleTable.from_files(sio.list_tfrecord_files(_resolve(emeta["mainEdgeInfo"]["tfrecordUriPrefix"], root)))
    src = edges.column(emeta["srcNodeIdKey"], "int64")
    dst = edges.column(emeta["dstNodeIdKey"], "int64")
    n_nodes = int(max(node_id.max(initial=-1), src.max(initial=-1), dst.max(initial=-1)) + 1)
    x = None
    if cols:
        feat = np.concatenate(cols, axis=1)
        x = np.zeros((n_nodes, feat.shape[1]), dtype=np.float32)
        x[node_id] = feat
    labels = None
    label_key = (nmeta.get("labelKeys") or [None])[0]
    if label_key and not skip_labeled:
        labels = np.full(n_nodes, sio.INT32_MIN, dtype=np.int32)
        labels[node_id] = nodes.column(label_key, "int64").astype(np.int32)
    log(f"[{job_name}] loaded {len(node_id)} nodes (F={0 if x is None else x.shape[1]}), {len(src)} edges in {time.time() - t0:.2f}s")

    ctx = Context(device)
    g = Graph.from_edges_host(ctx, n_nodes, src.astype(np.int32), dst.astype(np.int32), is_graph_directed=directed)
    roots_all = np.sort(node_id).astype(np.int32)  # every node of the node table gets exactly one RootedNodeNeighborhood
    stats = {"n_nodes": n_nodes, "n_edges_csr": g.n_edges, "rnn": 0, "snc": 0}
    unl_dir = _resolve(out["unlabeledTfrecordUriPrefix"], root)
    lab_dir = _resolve(out["labeledTfrecordUriPrefix"], root)
    os.makedirs(unl_dir, exist_ok=True)
    if labels is not None:
        os.makedirs(lab_dir, exist_ok=True)
    t1 = time.time()
    part = 0
    for s in range(0, len(roots_all), batch_roots):
        roots = roots_all[s:s + batch_roots]
        nbr, cnt = g.sample_khop_host(roots, fanouts, base_seed=SAMPLING_SEED, first_call_no=1)
        data, offs = sio.encode_samples(roots, fanouts, nbr, x, kind="rnn", condensed_node_type=ntype, condensed_edge_type=etype)
        with open(os.path.join(unl_dir, f"part-{part:05d}.tfrecord"), "wb") as f:  # RootedNodeNeighborhood first
            f.write(data)
        stats["rnn"] += len(roots)
        if labels is not None:
            # isolated nodes (no sampled in-edge) are NOT training samples (includeIsolatedNodesInTrainingSamples = false, :43-44)
            lab = labels.copy()
            lab[roots[cnt[0] == 0]] = sio.INT32_MIN
            if max_train > 0:
                keep = roots[(cnt[0] > 0) & (labels[roots] != sio.INT32_MIN)][max(0, max_train - stats["snc"]):]
                lab[keep] = sio.INT32_MIN
            data, offs = sio.encode_samples(roots, fanouts, nbr, x, kind="snc", condensed_node_type=ntype, condensed_edge_type=etype,
                                            labels=lab, label_type=label_key)
            with open(os.path.join(lab_dir, f"part-{part:05d}.tfrecord"), "wb") as f:
                f.write(data)
            stats["snc"] += int((np.diff(offs) > 0).sum())
        part += 1
    stats["seconds_sample_and_write"] = time.time() - t1
    stats["seconds_total"] = time.time() - t0
    log(f"[{job_name}] wrote {stats['rnn']} RootedNodeNeighborhood + {stats['snc']} SupervisedNodeClassificationSample "
        f"records in {stats['seconds_sample_and_write']:.2f}s")
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
