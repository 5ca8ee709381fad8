"""SamplingOp DAGs over typed edges (subgraph_sampling_strategy.proto:38-79): the per-hop, per-edge-type sampling that
only the reference's spark35 graph-DB path expresses (`SamplingOpDAG.from`, scala_spark35/common/src/main/scala/types/
SamplingOpDAG.scala:19-51; `GraphDBSampler.getKHopSubgraphForRootNode`, .../libs/sampler/GraphDBSampler.scala:40-148).

Every op expands the result nodes of its input ops (or the root) over ONE edge type: a uniform sample of `numNodesToSample`
edges per frontier node, or - RandomWeighted / TopK (proto :17-36; NebulaQueryResponseTranslator.scala:73-105) - the
`numNodesToSample` edges with the largest edge feature (times a seeded uniform draw for RandomWeighted).  Ops with at
most one input (tree-shaped DAGs: chains that branch) map one to one onto `gigl_sample_op_*` launches - one kernel
launch per op over that edge type's CSR, ancestors' padded trees as the frontier.  An op with several inputs (the union
of several parents' results, GraphDBSampler.scala:66-82) runs once per input - one launch per (op, input) instance -
and its result set is the union of the instances' outputs.  Like the reference, an op expands every DISTINCT frontier
node once per root: the parents' results are a HashSet[Node] there; here the parent level is passed through
`Context.frontier_distinct` (gigl_frontier_distinct_dev), which keeps the first slot of every node - also against the
op's earlier input instances - so a node reached along several paths contributes at most `numNodesToSample` edges per op.

The reference's graph-DB clients do not sample reproducibly (`LocalDbClient.scala:186,204` takes `Set.take(n)`; Nebula
samples server-side), so there is nothing bit-level to match: each op draws GiGL's seeded hash permutation with the op's
1-based position in the DAG as the permutation call number - a valid uniform sample without replacement, and exactly
`gigl_sample_khop_*` for a linear chain over one edge type.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

INCOMING, OUTGOING = "INCOMING", "OUTGOING"


@dataclass
class SamplingOp:
    op_name: str
    edge_type: Tuple[str, str, str]           # (src_node_type, relation, dst_node_type)
    num_nodes_to_sample: int
    input_op_names: List[str] = field(default_factory=list)
    sampling_direction: str = INCOMING
    sampling_method: str = "random_uniform"   # | "top_k" | "random_weighted" (SamplingOp.sampling_method, proto :38-58)
    edge_feat_name: str = ""                   # the edge feature a weighted method orders by (RandomWeighted / TopK .edge_feat_name)

    @property
    def frontier_node_type(self) -> str:
        """Type of the nodes this op expands: the dst type for INCOMING, the src type for OUTGOING."""
        return self.edge_type[2] if self.sampling_direction == INCOMING else self.edge_type[0]

    @property
    def result_node_type(self) -> str:
        return self.edge_type[0] if self.sampling_direction == INCOMING else self.edge_type[2]


class SamplingValidationError(ValueError):
    """A SamplingOp DAG the reference's config validation rejects.  `error_type` carries the name of the reference's
    SubgraphSamplingValidationErrorType member (python/gigl/src/common/types/exception.py:4-16)."""

    def __init__(self, message: str, error_type: str):
        super().__init__(message)
        self.error_type = error_type


_METHOD_KEYS = {"randomUniform": "random_uniform", "randomWeighted": "random_weighted", "topK": "top_k"}


def ops_from_config(path: dict) -> List[SamplingOp]:
    """`MessagePassingPath.samplingOps` (YAML / JSON form of the proto) -> SamplingOp list."""
    out = []
    for o in path.get("samplingOps") or []:
        given = [k for k in _METHOD_KEYS if k in o]
        if "userDefined" in o or len(given) != 1:  # the reference throws NotImplementedError for both (NebulaQueryResponseTranslator.scala:106-111)
            raise ValueError(f"sampling op {o.get('opName')!r}: exactly one of randomUniform / randomWeighted / topK is supported")
        m = o[given[0]]
        feat = str(m.get("edgeFeatName") or "")
        if given[0] != "randomUniform" and not feat:
            raise ValueError(f"sampling op {o.get('opName')!r}: {given[0]} needs an edgeFeatName")
        et = o["edgeType"]
        out.append(SamplingOp(o["opName"], (et["srcNodeType"], et["relation"], et["dstNodeType"]), int(m["numNodesToSample"]),
                              list(o.get("inputOpNames") or []), o.get("samplingDirection", INCOMING), _METHOD_KEYS[given[0]], feat))
    return out


@dataclass
class PlannedOp:
    op: SamplingOp
    call_no: int                  # 1-based position of the op in DAG order = the permutation call number
    parent: Optional[str]         # key of the input instance, None = expands the root
    chain: List[str]              # instance keys from the root down to and including this one
    key: str = ""                 # unique instance key: the op name, or "op@parent-key" for an op with several inputs
    fanouts: List[int] = field(default_factory=list)   # num_nodes_to_sample along `chain`


def plan(ops: Sequence[SamplingOp], root_node_type: str) -> List[PlannedOp]:
    """Topological order with type checks: an op's frontier type must be the result type of each of its inputs (the root
    type for root ops).  An op with several inputs expands the union of its inputs' result nodes
    (GraphDBSampler.scala:66-82); here that is one INSTANCE per input - each expands that input's padded tree - and the
    op's result set is the union of its instances' outputs.  Instances of one op share its call number."""
    by_name = {o.op_name: o for o in ops}
    if len(by_name) != len(ops):
        raise SamplingValidationError("duplicate op names", "REPEATED_OP_NAME")
    instances: Dict[str, List[PlannedOp]] = {}   # op name -> its planned instances
    order: List[PlannedOp] = []
    pending = list(ops)
    n_done = 0
    while pending:
        progressed = False
        for o in list(pending):
            if o.num_nodes_to_sample < 1:
                raise ValueError(f"op {o.op_name!r}: numNodesToSample must be >= 1")
            if len(set(o.input_op_names)) != len(o.input_op_names):
                raise ValueError(f"op {o.op_name!r} names an input op twice")
            for par in o.input_op_names:
                if par not in by_name:
                    raise SamplingValidationError(f"op {o.op_name!r} names an unknown input op {par!r}", "BAD_INPUT_OP_NAME")
            if any(par not in instances for par in o.input_op_names):
                continue
            n_done += 1
            made = []
            if not o.input_op_names:
                if o.frontier_node_type != root_node_type:
                    raise SamplingValidationError(f"op {o.op_name!r} expands {o.frontier_node_type!r} nodes but its input yields {root_node_type!r}",
                                                  "CONTAINS_INVALID_EDGE_IN_DAG")
                made.append(PlannedOp(o, n_done, None, [o.op_name], o.op_name, [o.num_nodes_to_sample]))
            for par in o.input_op_names:
                want = by_name[par].result_node_type
                if o.frontier_node_type != want:
                    raise SamplingValidationError(f"op {o.op_name!r} expands {o.frontier_node_type!r} nodes but its input yields {want!r}",
                                                  "CONTAINS_INVALID_EDGE_IN_DAG")
                for pi in instances[par]:
                    single = len(o.input_op_names) == 1 and len(instances[par]) == 1
                    key = o.op_name if single else f"{o.op_name}@{pi.key}"
                    made.append(PlannedOp(o, n_done, pi.key, pi.chain + [key], key, pi.fanouts + [o.num_nodes_to_sample]))
            instances[o.op_name] = made
            order += made
            pending.remove(o)
            progressed = True
        if not progressed:
            if not any(not o.input_op_names for o in ops):
                raise SamplingValidationError("no sampling op expands the root node", "MISSING_ROOT_SAMPLING_OP")
            raise SamplingValidationError("sampling ops contain a cycle", "DAG_CONTAINS_CYCLE")
    return order


def task_root_node_types(task_metadata: dict) -> set:
    """The node types a task roots its samples at (TaskMetadataPbWrapper.get_task_root_node_types,
    python/gigl/src/common/types/pb_wrappers/task_metadata.py:123-145): both endpoints of every supervision edge type of a
    link task, the supervision node types of a node task."""
    for key in ("nodeAnchorBasedLinkPredictionTaskMetadata", "linkBasedTaskMetadata"):
        if key in task_metadata:
            out = set()
            for e in task_metadata[key].get("supervisionEdgeTypes") or []:
                out |= {e["srcNodeType"], e["dstNodeType"]}
            return out
    if "nodeBasedTaskMetadata" in task_metadata:
        return set(task_metadata["nodeBasedTaskMetadata"].get("supervisionNodeTypes") or [])
    raise ValueError("taskMetadata names no task")


def validate_strategy(paths: Sequence[dict], graph_metadata: dict, task_metadata: dict) -> Dict[str, List[SamplingOp]]:
    """The checks of the reference's config validation on `subgraphSamplingStrategy.messagePassingPaths.paths`
    (SubgraphSamplingStrategyPbWrapper / MessagePassingPathPbWrapper, python/gigl/src/common/types/pb_wrappers/
    subgraph_sampling_strategy.py:22-283; SamplingOpPbWrapper.check_sampling_op_edge_type_validity), in its order and with
    its error types: per path unique op names and known input ops, one path per root node type; then per path the root type in
    the graph metadata and among the task's root types, a root op unless the path has no ops at all (a 0-hop neighbourhood),
    every op's edge type in the graph metadata and aligned with its inputs (an op expands the nodes its inputs yield - the
    four INCOMING / OUTGOING parent / child rules in one), no cycle; finally every root type of the task has a path.
    Returns {root node type: ops}."""
    by_root: Dict[str, List[SamplingOp]] = {}
    for path in paths:
        ops = ops_from_config(path)
        names = set()
        for o in ops:
            if o.op_name in names:
                raise SamplingValidationError(f"repeated op name {o.op_name!r} in one path", "REPEATED_OP_NAME")
            names.add(o.op_name)
        for o in ops:
            for par in o.input_op_names:
                if par not in names:
                    raise SamplingValidationError(f"op {o.op_name!r} names an unknown input op {par!r}", "BAD_INPUT_OP_NAME")
        root = path["rootNodeType"]
        if root in by_root:
            raise SamplingValidationError(f"two paths for root node type {root!r}", "REPEATED_ROOT_NODE_TYPE")
        by_root[root] = ops
    node_types = set(graph_metadata.get("nodeTypes") or [])
    edge_types = {(e["srcNodeType"], e["relation"], e["dstNodeType"]) for e in graph_metadata.get("edgeTypes") or []}
    expected = task_root_node_types(task_metadata)
    for root, ops in by_root.items():
        if root not in node_types:
            raise SamplingValidationError(f"root node type {root!r} is not in the graph metadata", "ROOT_NODE_TYPE_NOT_IN_GRAPH_METADATA")
        if root not in expected:
            raise SamplingValidationError(f"root node type {root!r} is neither in a supervision edge type nor a supervision node type",
                                          "ROOT_NODE_TYPE_NOT_IN_TASK_METADATA")
        expected.discard(root)
        if ops and not any(not o.input_op_names for o in ops):
            raise SamplingValidationError(f"the path of {root!r} has no sampling op from the root node", "MISSING_ROOT_SAMPLING_OP")
        by_name = {o.op_name: o for o in ops}
        for o in ops:
            if o.edge_type not in edge_types:
                raise SamplingValidationError(f"op {o.op_name!r}: edge type {o.edge_type} is not in the graph metadata",
                                              "SAMPLING_OP_EDGE_TYPE_NOT_IN_GRAPH_METADATA")
            for want in ([root] if not o.input_op_names else [by_name[par].result_node_type for par in o.input_op_names]):
                if o.frontier_node_type != want:
                    raise SamplingValidationError(f"op {o.op_name!r} expands {o.frontier_node_type!r} nodes but its input yields {want!r}",
                                                  "CONTAINS_INVALID_EDGE_IN_DAG")
        plan(ops, root)  # DAG_CONTAINS_CYCLE
    if expected:
        raise SamplingValidationError(f"no path for the task's root node types {sorted(expected)}", "MISSING_EXPECTED_ROOT_NODE_TYPE")
    return by_root


def frontier_of(p: PlannedOp, res: dict, frontiers: dict, distinct):
    """The level instance `p` expands: its parent's padded tree with, per root, only the first slot of every distinct node
    kept - against the parent's own lower slots and against the frontiers of the same op's earlier instances
    (GraphDBSampler.scala:66-82: one HashSet[Node] over all the parents' results).  `distinct(cur, slots, prev)` is
    `Context.frontier_distinct` on the device or `oracle.np_frontier_distinct` in the CPU tests."""
    slots = 1
    for f in p.fanouts[:-1]:
        slots *= f
    prev = frontiers.setdefault(p.op.op_name, [])
    level = distinct(res[p.parent][0], slots, list(prev))
    prev.append((level, slots))
    return level


def sample_dag(graphs: Dict[Tuple[Tuple[str, str, str], str], "object"], roots, ops: Sequence[SamplingOp], root_node_type: str,
               base_seed: int = 42, call_no_offset: int = 0, distinct_frontier: bool = True, weights=None):
    """Runs every op instance on the device.  `graphs[(edge_type, direction)]` = the :class:`gigl_b200.Graph` of that edge
    type, built by destination for INCOMING ops and `by_source=True` for OUTGOING ones.  `roots`: int32 CUDA tensor.
    Returns {instance key: (nbr, cnt, chain_fanouts)} with the padded-tree layout of `Graph.sample_khop` (the key is the op
    name unless the op has several inputs, see :func:`plan`).  distinct_frontier = False expands every slot of the parent
    level (one expansion per PATH, the pure-Spark GROUP BY semantics) instead of every distinct node.
    weights[(edge_type, direction, edge_feat_name)] = float32 CUDA tensor with that edge feature per CSR position of
    graphs[(edge_type, direction)] - needed by the top_k / random_weighted ops only."""
    res = {}
    frontiers: dict = {}
    for p in plan(ops, root_node_type):
        chain_nbr = [res[k][0] for k in p.chain[:-1]]
        g = graphs[(p.op.edge_type, p.op.sampling_direction)]
        if distinct_frontier and p.parent is not None:
            chain_nbr[-1] = frontier_of(p, res, frontiers, g.ctx.frontier_distinct)
        w = None
        if p.op.sampling_method != "random_uniform":
            w = (weights or {}).get((p.op.edge_type, p.op.sampling_direction, p.op.edge_feat_name))
            if w is None:
                raise ValueError(f"op {p.op.op_name!r}: no weights for edge feature {p.op.edge_feat_name!r} of {p.op.edge_type}")
        nbr, cnt = g.sample_op(roots, p.fanouts, chain_nbr, p.call_no + call_no_offset, base_seed, weights=w, method=p.op.sampling_method)
        res[p.key] = (nbr, cnt, p.fanouts)
    return res


def encoder_ops(planned: Sequence[PlannedOp], res: dict, cet_of: dict, cnt_of: dict) -> List[dict]:
    """The planned instances + their sampled padded trees as the op dicts `sample_io.encode_typed_samples` takes.
    cet_of: edge type triple -> condensed edge type; cnt_of: node type name -> condensed node type."""
    index = {p.key: i for i, p in enumerate(planned)}
    out = []
    for p in planned:
        nbr = res[p.key][0]
        out.append(dict(parent=-1 if p.parent is None else index[p.parent], fanout=p.op.num_nodes_to_sample,
                        condensed_edge_type=cet_of[p.op.edge_type], result_node_type=cnt_of[p.op.result_node_type],
                        outgoing=p.op.sampling_direction == OUTGOING, nbr=nbr.cpu().numpy() if hasattr(nbr, "cpu") else nbr))
    return out
