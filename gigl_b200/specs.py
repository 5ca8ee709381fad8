"""Trainer / Inferencer plug-ins with the reference's operator interface, running on the library's kernels.

Mirrors (duck-typed; ``import gigl`` itself needs tensorflow / PyG / GCP SDKs that this image does not have):

* ``BaseTrainer``:  ``init_model(gbml_config_pb_wrapper, state_dict=None)``, ``setup_for_training()``,
  ``train(gbml_config_pb_wrapper, device, profiler=None)``, ``eval(gbml_config_pb_wrapper, device)``,
  ``model`` property + setter, ``supports_distributed_training``
  (python/gigl/src/training/v1/lib/base_trainer.py:16-37, python/gigl/src/common/types/model.py:9-38);
* ``BaseInferencer``: ``infer_batch(batch, device) -> InferBatchResults(embeddings, predictions)``
  (python/gigl/src/inference/v1/lib/base_inferencer.py:23-57);
* behaviour of ``NodeClassificationModelingTaskSpec`` (python/gigl/src/common/modeling_task_specs/
  node_classification_modeling_task_spec.py:48-300): kwargs arrive as strings, Adam(lr, weight_decay), cross-entropy
  on ``out[root_node_indices]``, accuracy scoring; the model is GraphSAGE (BASELINE.json configs[0..1]) or the
  reference's TwoLayerGCN (``model=gcn``).

The batches come from :class:`ResidentGraphLoader` - the graph stays in HBM and a batch is
sample -> collate -> export on the device (what replaces TFRecord decoding + PygGraphBuilder + the collate functions
of python/gigl/src/training/v1/lib/data_loaders/) - or from :func:`batch_from_sample_protos` for samples the
reference's own Subgraph Sampler / Split Generator wrote.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, Iterator, List, NamedTuple, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import nn as gnn
from .engine import Batch, Context, Graph


@dataclass
class SampledNodeBatch:
    """SupervisedNodeClassificationBatch / RootedNodeNeighborhoodBatch of the reference
    (supervised_node_classification_data_loader.py:73-117): the coalesced batch graph in local ids."""

    x: torch.Tensor                      # [n, F] features of the batch nodes (row = local id)
    edge_index: torch.Tensor             # int64 [2, e], row 0 = src, row 1 = dst, local ids
    root_node_indices: torch.Tensor      # int64 [B] local ids of the roots
    root_nodes: torch.Tensor             # int32 [B] global ids of the roots
    root_node_labels: Optional[torch.Tensor] = None   # int64 [B]
    level_sizes: Optional[List[int]] = None           # dependency levels (roots first) for pruned execution
    node_ids: Optional[torch.Tensor] = None           # int32 [n] local -> global

    @property
    def graph(self):
        return self


class InferBatchResults(NamedTuple):
    embeddings: Optional[torch.Tensor]
    predictions: Optional[torch.Tensor]


class ResidentGraphLoader:
    """Yields :class:`SampledNodeBatch` for consecutive slices of ``root_ids``: k-hop sample, collate and export all on
    the device (gigl_sample_khop_dev -> gigl_batch_collate_dev -> gigl_batch_export_dev)."""

    def __init__(self, graph: Graph, x: torch.Tensor, root_ids, fanouts: Sequence[int], batch_size: int, num_layers: int,
                 labels: Optional[torch.Tensor] = None, shuffle: bool = False, seed: int = 0):
        self.g, self.x, self.fan, self.bs, self.L = graph, x, [int(f) for f in fanouts], int(batch_size), int(num_layers)
        self.roots = torch.as_tensor(np.asarray(root_ids), dtype=torch.int32, device=x.device)
        self.labels = labels
        self.shuffle, self.seed, self.epoch = shuffle, seed, 0
        self.batch = Batch(graph.ctx, graph.n_nodes)

    def __len__(self):
        return (self.roots.numel() + self.bs - 1) // self.bs

    def __iter__(self) -> Iterator[SampledNodeBatch]:
        roots = self.roots
        if self.shuffle:
            gen = torch.Generator(device="cpu").manual_seed(self.seed + self.epoch)
            roots = roots[torch.randperm(roots.numel(), generator=gen).to(roots.device)]
            self.epoch += 1
        for i in range(0, roots.numel(), self.bs):
            r = roots[i:i + self.bs].contiguous()
            nbr, _ = self.g.sample_khop(r, self.fan)
            levels = self.batch.collate(r, self.fan, nbr, self.L)
            node_ids, ei = self.batch.export()
            xb = self.x.index_select(0, node_ids.to(torch.int64))
            lab = None if self.labels is None else self.labels.index_select(0, r.to(torch.int64))
            yield SampledNodeBatch(x=xb, edge_index=ei, root_node_indices=torch.arange(r.numel(), device=r.device), root_nodes=r,
                                   root_node_labels=lab, level_sizes=levels, node_ids=node_ids)


def batch_from_sample_protos(samples: Iterable[dict], device, label_dtype=torch.int64) -> SampledNodeBatch:
    """Builds the batch from parsed sample protos (``gigl_b200.sample_io.parse_sample`` dicts of
    RootedNodeNeighborhood / SupervisedNodeClassificationSample) with the reference's graph-builder rules
    (abstract_graph_builder.py:49-197, pyg_graph_builder.py:20-69): nodes de-duplicated by id in first-seen order,
    edges de-duplicated by (src, dst), features stacked in local-id order, roots' local ids collected."""
    lid: Dict[int, int] = {}
    feats: List[np.ndarray] = []
    edges: Dict[tuple, None] = {}
    roots, labels = [], []
    for s in samples:
        for nd in s["nodes"]:
            if nd["node_id"] not in lid:
                lid[nd["node_id"]] = len(lid)
                feats.append(np.asarray(nd.get("feature_values", []), dtype=np.float32))
        for ed in s["edges"]:
            edges.setdefault((ed["src_node_id"], ed["dst_node_id"]), None)
        roots.append(s["root_node"]["node_id"])
        if s.get("root_node_labels"):
            labels.append(int(s["root_node_labels"][0]["label"]))
    ei = np.array([[lid[a] for a, _ in edges], [lid[b] for _, b in edges]], dtype=np.int64).reshape(2, -1)
    x = np.stack(feats) if feats else np.zeros((0, 0), np.float32)
    inv = np.fromiter(lid.keys(), dtype=np.int32, count=len(lid))
    return SampledNodeBatch(
        x=torch.from_numpy(x).to(device), edge_index=torch.from_numpy(ei).to(device),
        root_node_indices=torch.tensor([lid[r] for r in roots], dtype=torch.int64, device=device),
        root_nodes=torch.tensor(roots, dtype=torch.int32, device=device),
        root_node_labels=torch.tensor(labels, dtype=label_dtype, device=device) if len(labels) == len(roots) and roots else None,
        node_ids=torch.from_numpy(inv).to(device))


@dataclass
class SupervisionEdges:
    """NodeAnchorBasedLinkPredictionBatch.BatchSupervisionEdgeData of the reference
    (node_anchor_based_link_prediction_data_loader.py:58-66) for ONE condensed edge type."""

    root_node_to_target_node_id: Dict[int, torch.Tensor]                    # local root id -> int64 local ids of its targets
    label_edge_features: Optional[Dict[int, torch.Tensor]] = None           # local root id -> [n_targets, Fe] (user-defined labels)


@dataclass
class SampledLinkBatch:
    """NodeAnchorBasedLinkPredictionBatch of the reference (node_anchor_based_link_prediction_data_loader.py:68-230): the
    coalesced batch graph in local ids + per condensed edge type the roots' positive and hard-negative targets."""

    graph: SampledNodeBatch
    root_nodes: torch.Tensor                                                 # int64 [B] local ids of the roots, in sample order
    pos_supervision_edge_data: Dict[int, SupervisionEdges]
    hard_neg_supervision_edge_data: Dict[int, SupervisionEdges]
    edge_attr: Optional[torch.Tensor] = None                                 # [e, Fe] features of the message-passing edges


def link_batch_from_sample_protos(samples: Iterable[dict], device) -> SampledLinkBatch:
    """collate_pyg_node_anchor_based_link_prediction_minibatch (node_anchor_based_link_prediction_data_loader.py:90-230) on
    parsed NodeAnchorBasedLinkPredictionSample protos (``sample_io.parse_nablp_sample`` dicts): the neighbourhoods are
    unioned with the graph-builder rules of :func:`batch_from_sample_protos` - an edge shared by k samples is message-passed
    once, a node of two samples keeps the union of its sampled edges - and every root is mapped to the local ids of the
    targets of its pos / hard-neg edges, per condensed edge type, with the label edges' features when they carry any."""
    samples = list(samples)
    g = batch_from_sample_protos(samples, device)
    local = {int(v): i for i, v in enumerate(g.node_ids.tolist())}
    pos: Dict[int, SupervisionEdges] = {}
    neg: Dict[int, SupervisionEdges] = {}

    def fill(store, root, edges):
        by_type: Dict[int, list] = {}
        for e in edges:
            by_type.setdefault(int(e.get("condensed_edge_type") or 0), []).append(e)
        for cet, es in by_type.items():
            d = store.setdefault(cet, SupervisionEdges({}))
            d.root_node_to_target_node_id[root] = torch.tensor([local[e["dst_node_id"]] for e in es], dtype=torch.int64, device=device)
            feats = [np.asarray(e.get("feature_values") or [], dtype=np.float32) for e in es]
            if feats and all(len(f) > 0 for f in feats):
                if d.label_edge_features is None:
                    d.label_edge_features = {}
                d.label_edge_features[root] = torch.from_numpy(np.stack(feats)).to(device)

    roots = []
    for s in samples:
        r = local[s["root_node"]["node_id"]]
        roots.append(r)
        fill(pos, r, s.get("pos_edges") or [])
        fill(neg, r, s.get("hard_neg_edges") or [])
    feat_of = {}
    for s in samples:
        for ed in s["edges"]:
            feat_of.setdefault((ed["src_node_id"], ed["dst_node_id"]), ed.get("feature_values") or [])
    ids = g.node_ids.tolist()
    ei = g.edge_index.cpu().numpy()
    attrs = [np.asarray(feat_of[(ids[a], ids[b])], dtype=np.float32) for a, b in zip(ei[0], ei[1])]
    edge_attr = torch.from_numpy(np.stack(attrs)).to(device) if attrs and all(len(a) > 0 for a in attrs) else None
    return SampledLinkBatch(graph=g, root_nodes=torch.tensor(roots, dtype=torch.int64, device=device), pos_supervision_edge_data=pos,
                            hard_neg_supervision_edge_data=neg, edge_attr=edge_attr)


def _feature_dim(gbml_config_pb_wrapper, default: Optional[int]) -> int:
    w = gbml_config_pb_wrapper
    try:
        m = w.preprocessed_metadata_pb_wrapper.condensed_node_type_to_feature_dim_map
        return int(m[sorted(m.keys())[0]])
    except AttributeError:
        pass
    if isinstance(w, dict) and "in_dim" in w:
        return int(w["in_dim"])
    if default is None:
        raise ValueError("cannot infer the input feature dim: pass in_dim=... or a config wrapper that carries it")
    return default


class GraphSageB200Spec:
    """Node-classification trainer + inferencer over :mod:`gigl_b200.nn` (BaseTrainer + BaseInferencer duck type)."""

    def __init__(self, is_training: bool = True, **kwargs) -> None:
        self._lr = float(kwargs.get("optim_lr", 0.01))
        self._wd = float(kwargs.get("optim_weight_decay", 5e-4))
        self._num_epochs = int(kwargs.get("num_epochs", 5))
        self._out_dim = int(kwargs.get("out_dim", 7))
        self._hid_dim = int(kwargs.get("hid_dim", 16))
        self._num_layers = int(kwargs.get("num_layers", 2))
        self._in_dim = int(kwargs["in_dim"]) if "in_dim" in kwargs else None
        self._model_kind = str(kwargs.get("model", "graphsage")).lower()
        self._prune = str(kwargs.get("prune_to_roots", "true")).lower() in ("1", "true", "yes")
        self._is_training = is_training
        self.main_sample_batch_size = int(kwargs.get("main_sample_batch_size", 16))
        self._model: Optional[torch.nn.Module] = None
        self._gbml_config_pb_wrapper = None
        self.loaders: Dict[str, Iterable[SampledNodeBatch]] = {}

    # ---- BaseModelOperationsProtocol ------------------------------------------------------------
    @property
    def model(self) -> torch.nn.Module:
        return self._model

    @model.setter
    def model(self, model: torch.nn.Module) -> None:
        self._model = model

    @property
    def supports_distributed_training(self) -> bool:
        return True

    def init_model(self, gbml_config_pb_wrapper=None, state_dict=None) -> torch.nn.Module:
        self._gbml_config_pb_wrapper = gbml_config_pb_wrapper
        in_dim = _feature_dim(gbml_config_pb_wrapper, self._in_dim)
        if self._model_kind == "gcn":
            model = gnn.TwoLayerGCN(in_dim=in_dim, out_dim=self._out_dim, hid_dim=self._hid_dim, is_training=self._is_training)
        else:
            model = gnn.GraphSAGE(in_dim, self._hid_dim, self._num_layers, self._out_dim)
        if state_dict is not None:
            model.load_state_dict(state_dict)
        self.model = model
        self._graph_backend = model.graph_backend
        return model

    # ---- BaseTrainer --------------------------------------------------------------------------
    def setup_for_training(self) -> None:
        self._optimizer = torch.optim.Adam(self.model.parameters(), lr=self._lr, weight_decay=self._wd)
        self._train_loss_fn = lambda input, target: F.cross_entropy(input=input, target=target)
        self.model.train()

    def _forward(self, batch: SampledNodeBatch, device) -> torch.Tensor:
        """model(x, edge_index)[root_node_indices] (node_classification_modeling_task_spec.py:160-166)."""
        x, ei = batch.x.to(device), batch.edge_index.to(device)
        module = self.model.module if hasattr(self.model, "module") else self.model  # DistributedDataParallel wrapper
        if isinstance(module, gnn.GraphSAGE) and self._prune and batch.level_sizes is not None:
            # roots are local ids 0..B-1 and only the rows they depend on are computed: identical root rows
            return self.model(x, ei, batch.level_sizes)
        return self.model(x, ei)[batch.root_node_indices.to(device)]

    def _train(self, data_loader: Iterable[SampledNodeBatch], device) -> Optional[torch.Tensor]:
        self.model.train()
        loss = None
        for batch in data_loader:
            assert batch.root_node_labels is not None, "Labels required for training."
            self._optimizer.zero_grad()
            out = self._forward(batch, device)
            loss = self._train_loss_fn(input=out, target=batch.root_node_labels.to(device))
            loss.backward()
            self._optimizer.step()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.barrier()
        return loss

    def train(self, gbml_config_pb_wrapper=None, device=torch.device("cuda", 0), profiler=None) -> Dict[str, float]:
        """Epoch loop of NodeClassificationModelingTaskSpec.train (:227-268); ``self.loaders`` must hold 'train_main'
        and optionally 'val_main'.  Returns the last loss / best validation accuracy instead of writing tf.summary."""
        best_val_acc, last = 0.0, None
        for _ in range(self._num_epochs):
            last = self._train(self.loaders["train_main"], device)
            if "val_main" in self.loaders:
                best_val_acc = max(best_val_acc, self.score(self.loaders["val_main"], device))
        return {"train_loss": float(last.detach()) if last is not None else float("nan"), "best_val_acc": best_val_acc}

    def eval(self, gbml_config_pb_wrapper=None, device=torch.device("cuda", 0)) -> Dict[str, float]:
        """EvalMetricsCollection(metrics=[acc]) of the reference (:270-300), as a plain dict."""
        return {"acc": self.score(self.loaders["test_main"], device)}

    # ---- BaseInferencer -----------------------------------------------------------------------
    @torch.no_grad()
    def infer_batch(self, batch: SampledNodeBatch, device=torch.device("cuda", 0)) -> InferBatchResults:
        was_training = self.model.training
        self.model.eval()
        try:
            embed = self._forward(batch, device)
        finally:
            self.model.train(was_training)
        return InferBatchResults(embeddings=embed, predictions=embed.argmax(dim=1))

    @torch.no_grad()
    def score(self, data_loader: Iterable[SampledNodeBatch], device) -> float:
        num_correct = num_evaluated = 0
        for batch in data_loader:
            assert batch.root_node_labels is not None, "Labels required for scoring."
            res = self.infer_batch(batch, device)
            num_correct += int((res.predictions == batch.root_node_labels.to(device)).sum())
            num_evaluated += int(batch.root_node_labels.numel())
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            t = torch.tensor([num_correct, num_evaluated], dtype=torch.float64, device=device)
            torch.distributed.all_reduce(t)
            num_correct, num_evaluated = int(t[0].item()), int(t[1].item())
        return num_correct / max(num_evaluated, 1)
