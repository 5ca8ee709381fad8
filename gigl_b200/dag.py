"""SamplingOp DAGs over typed edges (subgraph_sampling_strategy.proto:38-79): the per-hop, per-edge-type sampling that
only the reference's spark35 graph-DB path expresses (`SamplingOpDAG.from`, scala_spark35/common/src/main/scala/types/
SamplingOpDAG.scala:19-51; `GraphDBSampler.getKHopSubgraphForRootNode`, .../libs/sampler/GraphDBSampler.scala:40-148).

Every op expands the result nodes of its parent op (or the root) over ONE edge type with a uniform fanout.  Ops with at
most one input (tree-shaped DAGs: chains that branch) map one to one onto `gigl_sample_op_*` launches - one kernel
launch per op over that edge type's CSR, ancestors' padded trees as the frontier.  Ops with several inputs (the union
of several parents' results) are not supported.

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

    @property
    def frontier_node_type(self) -> str:
        """Type of the nodes this op expands: the dst type for INCOMING, the src type for OUTGOING."""
        return self.edge_type[2] if self.sampling_direction == INCOMING else self.edge_type[0]

    @property
    def result_node_type(self) -> str:
        return self.edge_type[0] if self.sampling_direction == INCOMING else self.edge_type[2]


def ops_from_config(path: dict) -> List[SamplingOp]:
    """`MessagePassingPath.samplingOps` (YAML / JSON form of the proto) -> SamplingOp list."""
    out = []
    for o in path.get("samplingOps") or []:
        if "randomUniform" not in o:
            raise ValueError(f"sampling op {o.get('opName')!r}: only randomUniform sampling is supported")
        et = o["edgeType"]
        out.append(SamplingOp(o["opName"], (et["srcNodeType"], et["relation"], et["dstNodeType"]),
                              int(o["randomUniform"]["numNodesToSample"]), list(o.get("inputOpNames") or []),
                              o.get("samplingDirection", INCOMING)))
    return out


@dataclass
class PlannedOp:
    op: SamplingOp
    call_no: int                  # 1-based position in DAG order = the permutation call number
    parent: Optional[str]         # op_name of the input op, None = expands the root
    chain: List[str]              # op names from the root down to and including this op


def plan(ops: Sequence[SamplingOp], root_node_type: str) -> List[PlannedOp]:
    """Topological order with type checks: an op's frontier type must be its parent's result type (the root type for
    root ops)."""
    by_name = {o.op_name: o for o in ops}
    if len(by_name) != len(ops):
        raise ValueError("duplicate op names")
    planned: Dict[str, PlannedOp] = {}
    order: List[PlannedOp] = []
    pending = list(ops)
    while pending:
        progressed = False
        for o in list(pending):
            if len(o.input_op_names) > 1:
                raise ValueError(f"op {o.op_name!r} has several input ops: only tree-shaped DAGs are supported")
            if o.num_nodes_to_sample < 1:
                raise ValueError(f"op {o.op_name!r}: numNodesToSample must be >= 1")
            par = o.input_op_names[0] if o.input_op_names else None
            if par is not None and par not in by_name:
                raise ValueError(f"op {o.op_name!r} names an unknown input op {par!r}")
            if par is not None and par not in planned:
                continue
            want = root_node_type if par is None else planned[par].op.result_node_type
            if o.frontier_node_type != want:
                raise ValueError(f"op {o.op_name!r} expands {o.frontier_node_type!r} nodes but its input yields {want!r}")
            p = PlannedOp(o, len(order) + 1, par, (planned[par].chain if par else []) + [o.op_name])
            planned[o.op_name] = p
            order.append(p)
            pending.remove(o)
            progressed = True
        if not progressed:
            raise ValueError("sampling ops contain a cycle")
    return order


def sample_dag(graphs: Dict[Tuple[Tuple[str, str, str], str], "object"], roots, ops: Sequence[SamplingOp], root_node_type: str,
               base_seed: int = 42):
    """Runs every op on the device.  `graphs[(edge_type, direction)]` = the :class:`gigl_b200.Graph` of that edge type,
    built by destination for INCOMING ops and `by_source=True` for OUTGOING ones.  `roots`: int32 CUDA tensor.
    Returns {op_name: (nbr, cnt, chain_fanouts)} with the padded-tree layout of `Graph.sample_khop`."""
    res = {}
    for p in plan(ops, root_node_type):
        fan = [next(q for q in ops if q.op_name == name).num_nodes_to_sample for name in p.chain]
        chain_nbr = [res[name][0] for name in p.chain[:-1]]
        g = graphs[(p.op.edge_type, p.op.sampling_direction)]
        nbr, cnt = g.sample_op(roots, fan, chain_nbr, p.call_no, base_seed)
        res[p.op.op_name] = (nbr, cnt, fan)
    return res
