"""CPU oracle for the GiGL hot path - TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and there only as the checker or the
timed CPU baseline.  Nothing under ``gigl_b200/`` imports it.

Two independent restatements of the same spec:

* ``c_*``  - ctypes bindings to ``oracle/gigl_oracle.c`` (plain C, OpenMP over roots);
* ``np_*`` - numpy / pure-python versions used to cross-check the C one on small cases.

Reference files followed (relative to /root/reference): see the header of gigl_oracle.c.
Pinning status: **parity unpinned** for the deterministic permutation (the reference's tests
only ever run the non-deterministic shuffle); pinned for the hash (Spark xxhash64 KAT) and
for the structural rules (reference sgs_output fixtures).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgigl_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/gigl_oracle.c (gcc) if the .so is missing or stale."""
    src = os.path.join(_HERE, "gigl_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgigl_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        i32, i64, u64, p = C.c_int32, C.c_int64, C.c_uint64, C.c_void_p
        L.oracle_xxh64_int.restype = i64
        L.oracle_xxh64_int.argtypes = [i32, u64]
        L.oracle_perm_full.restype = C.c_int
        L.oracle_perm_full.argtypes = [i64, i32, i32, p]
        L.oracle_perm_topk.restype = i64
        L.oracle_perm_topk.argtypes = [i64, i32, i32, i32, p]
        L.oracle_sample_khop.restype = C.c_int
        L.oracle_sample_khop.argtypes = [i64, p, p, p, i64, p, i32, i32, i32, p, p, i32]
        for nm in ("oracle_sage_conv_f32", "oracle_sage_conv_f64"):
            f = getattr(L, nm)
            f.restype = C.c_int
            f.argtypes = [i64, i64, i32, i32, p, p, p, p, p, p, p, i32, i32]
        for nm in ("oracle_gcn_conv_f32", "oracle_gcn_conv_f64"):
            f = getattr(L, nm)
            f.restype = C.c_int
            f.argtypes = [i64, i64, i32, i32, p, p, p, p, p, p, i32]
        L.oracle_collate.restype = C.c_int
        L.oracle_collate.argtypes = [i64, p, i64, p, i32, p, p, p, p, p, p, p, i32]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------------------
# hash + permutation
# ----------------------------------------------------------------------------------------
_P1 = np.uint64(0x9E3779B185EBCA87)
_P2 = np.uint64(0xC2B2AE3D27D4EB4F)
_P3 = np.uint64(0x165667B19E3779F9)
_P5 = np.uint64(0x27D4EB2F165667C5)


def np_xxh64_int(x, seed: int = 42) -> np.ndarray:
    """Spark ``XXH64.hashInt(int, seed)`` on an int32 array -> signed int64 array."""
    x = np.asarray(x).astype(np.int64).astype(np.int32)  # wrap to int32
    with np.errstate(over="ignore"):
        h = np.uint64(seed) + _P5 + np.uint64(4)
        h = h ^ (x.astype(np.uint32).astype(np.uint64) * _P1)
        h = ((h << np.uint64(23)) | (h >> np.uint64(41))) * _P2 + _P3
        h = h ^ (h >> np.uint64(33))
        h = h * _P2
        h = h ^ (h >> np.uint64(29))
        h = h * _P3
        h = h ^ (h >> np.uint64(32))
    return h.astype(np.int64)


def c_xxh64_int(x: int, seed: int = 42) -> int:
    x = ((int(x) + 2**31) % 2**32) - 2**31
    return int(lib().oracle_xxh64_int(x, seed))


def _wrap32(v: int) -> int:
    return ((int(v) + 2**31) % 2**32) - 2**31


def np_perm(size: int, internal_seed: int, current_seed: int) -> np.ndarray:
    """0-based permutation of range(size) per SamplingStrategy.scala:47-76."""
    i = np.arange(1, size + 1, dtype=np.int64)
    x = i + _wrap32(internal_seed) + _wrap32(current_seed)  # exact in int64, wrapped in np_xxh64_int
    keys = np_xxh64_int(x)
    order = np.lexsort((i, keys))  # ascending by (key, idx)
    return order.astype(np.int64)


def c_perm_full(size: int, internal_seed: int, current_seed: int) -> np.ndarray:
    out = np.empty(max(size, 1), dtype=np.int64)
    rc = lib().oracle_perm_full(size, _wrap32(internal_seed), _wrap32(current_seed), _ptr(out))
    assert rc == 0
    return out[:size]


def c_perm_topk(size: int, internal_seed: int, current_seed: int, f: int) -> np.ndarray:
    out = np.empty(max(f, 1), dtype=np.int64)
    n = lib().oracle_perm_topk(size, _wrap32(internal_seed), _wrap32(current_seed), f, _ptr(out))
    return out[:n]


# ----------------------------------------------------------------------------------------
# graph loading rules (SGSPureSparkV1Task.scala:120-286) -> CSR by dst, rows sorted ascending
# ----------------------------------------------------------------------------------------
def np_build_in_csr(src, dst, n_nodes: int, is_graph_directed: bool):
    """Edge list -> (rowptr int64 [n+1], col int32 [E]) of sorted in-neighbour lists.

    Undirected (``enforceBidirectionalization`` :218-258): DISTINCT (least, greatest) then union
    with the reverse.  Directed: duplicates kept.
    """
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    if not is_graph_directed:
        lo = np.minimum(src, dst)
        hi = np.maximum(src, dst)
        pairs = np.unique(np.stack([lo, hi], 1), axis=0) if len(lo) else np.zeros((0, 2), np.int64)
        # Spark SQL `UNION` is UNION DISTINCT (:238-252): (lo,hi) U (hi,lo) de-duplicated, so a
        # self loop (u,u) survives exactly once.
        both = np.unique(np.concatenate([pairs, pairs[:, ::-1]]), axis=0)
        src, dst = both[:, 0], both[:, 1]
    order = np.lexsort((src, dst))
    src, dst = src[order], dst[order]
    rowptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, src.astype(np.int32)


# ----------------------------------------------------------------------------------------
# k-hop sampling
# ----------------------------------------------------------------------------------------
def c_sample_khop(rowptr, col, roots, fanouts, base_seed=42, first_call_no=1, n_threads=None):
    """-> (nbr, cnt): lists per hop of the padded-tree arrays (see gigl_oracle.c)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    roots = np.ascontiguousarray(roots, dtype=np.int32)
    fan = np.ascontiguousarray(fanouts, dtype=np.int32)
    n_roots = len(roots)
    nbr, cnt = [], []
    width = 1
    for f in fan:
        cnt.append(np.zeros(n_roots * width, dtype=np.int32))
        width *= int(f)
        nbr.append(np.full(n_roots * width, -1, dtype=np.int32))
    pn = (C.c_void_p * len(fan))(*[a.ctypes.data for a in nbr])
    pc = (C.c_void_p * len(fan))(*[a.ctypes.data for a in cnt])
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    rc = lib().oracle_sample_khop(
        len(rowptr) - 1, _ptr(rowptr), _ptr(col), _ptr(roots), n_roots, _ptr(fan), len(fan), base_seed,
        first_call_no, pn, pc, n_threads,
    )
    if rc != 0:
        raise RuntimeError(f"oracle_sample_khop failed rc={rc}")
    return nbr, cnt


def np_sample_khop(rowptr, col, roots, fanouts, base_seed=42, first_call_no=1):
    """Pure-python restatement of the same padded-tree sampler (small inputs only)."""
    n_roots = len(roots)
    nbr, cnt = [], []
    width_prev = 1
    prev_vals = [int(r) for r in roots]
    prev_sums = [int(r) for r in roots]
    for h, f in enumerate(fanouts, start=1):
        f = int(f)
        cur_seed = _wrap32(base_seed * _wrap32(first_call_no + h - 1))
        out = np.full(n_roots * width_prev * f, -1, dtype=np.int32)
        oc = np.zeros(n_roots * width_prev, dtype=np.int32)
        sums = [0] * (n_roots * width_prev * f)
        for ps in range(n_roots * width_prev):
            v = prev_vals[ps]
            if v < 0:
                continue
            m = 1
            if h > 1:
                fp = int(fanouts[h - 2])
                sib0 = (ps // fp) * fp
                sib = prev_vals[sib0 : sib0 + fp]
                if sib.index(v) + sib0 != ps:
                    continue
                m = sib.count(v)
            row = col[rowptr[v] : rowptr[v + 1]]
            arr = np.repeat(row, m)  # sorted(m copies of a sorted list)
            order = np_perm(len(arr), prev_sums[ps], cur_seed)[:f]
            sel = arr[order]
            out[ps * f : ps * f + len(sel)] = sel
            oc[ps] = len(sel)
            for j, s in enumerate(sel):
                sums[ps * f + j] = _wrap32(prev_sums[ps] + int(s))
        nbr.append(out)
        cnt.append(oc)
        prev_vals = [int(x) for x in out]
        prev_sums = sums
        width_prev *= f
    return nbr, cnt


def np_sample_chain(csrs, roots, fanouts, call_nos, base_seed=42):
    """np_sample_khop with a CSR and a permutation call number PER HOP: one root-to-op chain of a SamplingOp DAG
    (subgraph_sampling_strategy.proto:38-58), hop h expanding over csrs[h] = (rowptr, col) of that op's edge type."""
    n_roots = len(roots)
    nbr, cnt = [], []
    width_prev = 1
    prev_vals = [int(r) for r in roots]
    prev_sums = [int(r) for r in roots]
    for h, f in enumerate(fanouts):
        f = int(f)
        rowptr, col = csrs[h]
        cur_seed = _wrap32(base_seed * _wrap32(call_nos[h]))
        out = np.full(n_roots * width_prev * f, -1, dtype=np.int32)
        oc = np.zeros(n_roots * width_prev, dtype=np.int32)
        sums = [0] * (n_roots * width_prev * f)
        for ps in range(n_roots * width_prev):
            v = prev_vals[ps]
            if v < 0:
                continue
            m = 1
            if h > 0:
                fp = int(fanouts[h - 1])
                sib0 = (ps // fp) * fp
                sib = prev_vals[sib0 : sib0 + fp]
                if sib.index(v) + sib0 != ps:
                    continue
                m = sib.count(v)
            arr = np.repeat(col[rowptr[v] : rowptr[v + 1]], m)
            sel = arr[np_perm(len(arr), prev_sums[ps], cur_seed)[:f]]
            out[ps * f : ps * f + len(sel)] = sel
            oc[ps] = len(sel)
            for j, s_ in enumerate(sel):
                sums[ps * f + j] = _wrap32(prev_sums[ps] + int(s_))
        nbr.append(out)
        cnt.append(oc)
        prev_vals = [int(x) for x in out]
        prev_sums = sums
        width_prev *= f
    return nbr, cnt


def np_sample_op(csr, roots, fanouts, chain_nbr, call_no, base_seed=42):
    """ONE op of a SamplingOp DAG given its ancestors' padded trees (the inputs of gigl_sample_op_*): the last hop of
    np_sample_chain.  fanouts = the ancestors' fanouts followed by this op's; chain_nbr = the ancestors' outputs."""
    n_roots = len(roots)
    depth = len(fanouts)
    f = int(fanouts[-1])
    rowptr, col = csr
    width_prev = 1
    for g in fanouts[:-1]:
        width_prev *= int(g)
    prev_vals = [int(r) for r in roots] if depth == 1 else [int(v) for v in chain_nbr[-1]]
    prev_sums = [0] * (n_roots * width_prev)
    for ps in range(n_roots * width_prev):  # wrapping int32 sum of the path ids root .. parent
        tot, s_ = 0, ps
        for h in range(depth - 1, 0, -1):
            tot = _wrap32(tot + int(chain_nbr[h - 1][s_]))
            s_ //= int(fanouts[h - 1])
        prev_sums[ps] = _wrap32(tot + int(roots[s_]))
    cur_seed = _wrap32(base_seed * _wrap32(call_no))
    out = np.full(n_roots * width_prev * f, -1, dtype=np.int32)
    oc = np.zeros(n_roots * width_prev, dtype=np.int32)
    for ps in range(n_roots * width_prev):
        v = prev_vals[ps]
        if v < 0:
            continue
        m = 1
        if depth > 1:
            fp = int(fanouts[-2])
            sib0 = (ps // fp) * fp
            sib = prev_vals[sib0 : sib0 + fp]
            if sib.index(v) + sib0 != ps:
                continue
            m = sib.count(v)
        arr = np.repeat(col[rowptr[v] : rowptr[v + 1]], m)
        sel = arr[np_perm(len(arr), prev_sums[ps], cur_seed)[:f]]
        out[ps * f : ps * f + len(sel)] = sel
        oc[ps] = len(sel)
    return out, oc


def np_sample_op_weighted(csr, weights, roots, fanouts, chain_nbr, call_no, method, base_seed=42):
    """ONE TopK / RandomWeighted op (subgraph_sampling_strategy.proto:17-36) as NebulaQueryResponseTranslator.scala:73-105
    words it - ORDER BY score DESC LIMIT numNodesToSample over the frontier node's edges of one type, score = the edge feature
    (method "top_k") or the feature times a uniform draw ("random_weighted").  The reference's rand() is unseeded; the draw
    here is u = (top 52 bits of the op's permutation key of the window position + 1/2) * 2^-52 (the key of np_perm with its
    sign bit flipped = the unsigned order of the signed hash).  Ties: lower CSR position first; NaN scores last; a frontier
    node repeated among its siblings is expanded at its first slot only.  weights: float32 per CSR position."""
    assert method in ("top_k", "random_weighted")
    n_roots = len(roots)
    depth = len(fanouts)
    f = int(fanouts[-1])
    rowptr, col = csr
    weights = np.asarray(weights, dtype=np.float32)
    width_prev = 1
    for g in fanouts[:-1]:
        width_prev *= int(g)
    prev_vals = [int(r) for r in roots] if depth == 1 else [int(v) for v in chain_nbr[-1]]
    cur_seed = _wrap32(base_seed * _wrap32(call_no))
    out = np.full(n_roots * width_prev * f, -1, dtype=np.int32)
    oc = np.zeros(n_roots * width_prev, dtype=np.int32)
    for ps in range(n_roots * width_prev):
        v = prev_vals[ps]
        if v < 0:
            continue
        if depth > 1:
            fp = int(fanouts[-2])
            sib0 = (ps // fp) * fp
            if prev_vals[sib0 : sib0 + fp].index(v) + sib0 != ps:
                continue
        tot, s_ = 0, ps
        for h in range(depth - 1, 0, -1):
            tot = _wrap32(tot + int(chain_nbr[h - 1][s_]))
            s_ //= int(fanouts[h - 1])
        path_sum = _wrap32(tot + int(roots[s_]))
        lo, hi = int(rowptr[v]), int(rowptr[v + 1])
        score = weights[lo:hi].astype(np.float64)
        if method == "random_weighted":
            i = np.arange(1, hi - lo + 1, dtype=np.int64)
            key = np_xxh64_int(i + path_sum + cur_seed).view(np.uint64) ^ np.uint64(1 << 63)
            score = score * (((key >> np.uint64(12)).astype(np.float64) + 0.5) * 2.0 ** -52)
        score = np.where(np.isnan(score), -np.inf, score)
        sel = col[lo:hi][np.argsort(-score, kind="stable")[:f]]
        out[ps * f : ps * f + len(sel)] = sel
        oc[ps] = len(sel)
    return out, oc


def np_frontier_distinct(cur, cur_slots: int, prev=()):
    """The frontier a SamplingOp expands: per root, the SET of its parents' result nodes
    (GraphDBSampler.scala:66-82 collects them into a HashSet[Node]).  cur: [n_roots * cur_slots] parent level; prev:
    [(level, slots), ...] of the op's earlier input instances.  Returns cur with every later occurrence of a node (per
    root; -1 = empty) replaced by -1."""
    cur = np.asarray(cur, dtype=np.int32).reshape(-1, cur_slots)
    out = cur.copy()
    prev = [(np.asarray(p, dtype=np.int32).reshape(-1, s)) for p, s in prev]
    for r in range(cur.shape[0]):
        seen = set()
        for p in prev:
            seen.update(int(v) for v in p[r] if v >= 0)
        for s in range(cur_slots):
            v = int(cur[r, s])
            if v < 0:
                continue
            if v in seen:
                out[r, s] = -1
            else:
                seen.add(v)
    return out.reshape(-1)


def np_sample_dag(planned, csr_of, roots, base_seed=42, call_no_offset=0, distinct=True, weights_of=None):
    """A whole SamplingOp DAG on the CPU, the way gigl_b200.dag.sample_dag runs it on the device: the planned op instances
    in order, each ONE np_sample_op over csr_of(instance) = (rowptr, col) of its edge type and direction, expanding its
    parent's level reduced to the distinct nodes per root (np_frontier_distinct: the HashSet[Node] of
    GraphDBSampler.scala:66-82), also against the earlier instances of the same op.  `planned`: objects with .key,
    .parent, .chain, .fanouts, .call_no and .op.op_name (gigl_b200.dag.plan output).  -> {key: (nbr, cnt, fanouts)}"""
    res, frontiers = {}, {}
    for p in planned:
        chain_nbr = [res[k][0] for k in p.chain[:-1]]
        if distinct and p.parent is not None:
            slots = int(np.prod(p.fanouts[:-1], dtype=np.int64))
            prev = frontiers.setdefault(p.op.op_name, [])
            level = np_frontier_distinct(res[p.parent][0], slots, list(prev))
            prev.append((level, slots))
            chain_nbr[-1] = level
        method = getattr(p.op, "sampling_method", "random_uniform")
        if method == "random_uniform":
            nbr, cnt = np_sample_op(csr_of(p), roots, p.fanouts, chain_nbr, p.call_no + call_no_offset, base_seed)
        else:  # weights_of(instance) = the op's edge feature per CSR position of csr_of(instance)
            nbr, cnt = np_sample_op_weighted(csr_of(p), weights_of(p), roots, p.fanouts, chain_nbr, p.call_no + call_no_offset, method, base_seed)
        res[p.key] = (nbr, cnt, list(p.fanouts))
    return res


def tree_to_edges(roots, nbr, fanouts):
    """Padded tree -> per-root list of (src, dst) index pairs, src = hop-k node, dst = hop-(k-1)
    node (SGSPureSparkV1Task.scala:615-629), one pair per sampled slot (explode semantics)."""
    n_roots = len(roots)
    res = [[] for _ in range(n_roots)]
    prev = np.asarray(roots, dtype=np.int64)
    width_prev = 1
    for h, f in enumerate(fanouts):
        f = int(f)
        cur = np.asarray(nbr[h], dtype=np.int64)
        for r in range(n_roots):
            for ps in range(width_prev):
                p = prev[r * width_prev + ps]
                if p < 0:
                    continue
                for j in range(f):
                    c = cur[(r * width_prev + ps) * f + j]
                    if c >= 0:
                        res[r].append((int(c), int(p)))
        prev = cur
        width_prev *= f
    return res


# ----------------------------------------------------------------------------------------
# hydration + sample assembly (pure python, small inputs): the SQL of SGSPureSparkV1Task.scala:496-820,
# NodeAnchorBasedLinkPredictionBaseTask.scala:19-334 and NodeAnchorBasedLinkPredictionTask.scala:146-312
# restated over python lists / dicts.  Arrays that Spark builds with collect_list have no defined
# order, so samples are returned in a canonical (sorted) form and compared as multisets.
# ----------------------------------------------------------------------------------------
def np_hydrated_edge_table(src, dst, is_graph_directed: bool, edge_feat=None):
    """loadEdgeDataframeIntoSparkSql (:120-286): {(from, to): [feature tuple per record]}.
    Directed: every input record, duplicates kept.  Undirected: one record per (least, greatest) pair (the
    reference's dropDuplicates keeps an arbitrary one; this restatement keeps the FIRST input record), unioned
    with its reverse (UNION DISTINCT: a self loop once)."""
    src = [int(v) for v in src]
    dst = [int(v) for v in dst]
    feat = [tuple(np.asarray(edge_feat[i], dtype=np.float32).tolist()) if edge_feat is not None else () for i in range(len(src))]
    table = {}
    if is_graph_directed:
        for s_, d_, f_ in zip(src, dst, feat):
            table.setdefault((s_, d_), []).append(f_)
        return table
    seen = {}
    for s_, d_, f_ in zip(src, dst, feat):
        seen.setdefault((min(s_, d_), max(s_, d_)), f_)
    for (lo, hi), f_ in seen.items():
        table[(lo, hi)] = [f_]
        table[(hi, lo)] = [f_]
    return table


def np_sample_positives(out_rowptr, out_col, srcs, num_pos, base_seed=42, call_no=3):
    """sampleDstNodesUniformly (NodeAnchorBasedLinkPredictionBaseTask.scala:19-104): per source with an out-edge the
    first num_pos of perm(sorted destinations, internal seed = the source id, seed x call_no)."""
    cur_seed = _wrap32(base_seed * _wrap32(call_no))
    res = {}
    for u in srcs:
        u = int(u)
        row = np.asarray(out_col[out_rowptr[u] : out_rowptr[u + 1]])
        if len(row):
            res[u] = [int(v) for v in row[np_perm(len(row), u, cur_seed)[:num_pos]]]
    return res


def _neighborhood(root, tree_edges, table):
    """createSubgraph: edges = CONCAT(hop edges) with every sampled pair joined to its edge records (hydrateEdges),
    nodes = array_distinct(hop nodes ++ root)."""
    edges = [(c, p_, f_) for c, p_ in tree_edges for f_ in table.get((c, p_), [])]
    nodes = {root} | {c for c, _ in tree_edges}
    return edges, nodes


def np_assemble_rnn(roots, nbr, fanouts, table):
    """{root: (sorted edge multiset [(src, dst, feat)], sorted node ids)} for every root (isolated roots: no edges,
    nodes = [root]; createRootedNodeNeighborhoodSubgraph :847-1017)."""
    te = tree_to_edges(roots, nbr, fanouts)
    out = {}
    for i, r in enumerate(roots):
        e, n = _neighborhood(int(r), te[i], table)
        out[int(r)] = (sorted(e), sorted(n))
    return out


def np_assemble_nablp(roots, nbr, fanouts, table, positives, pos_table=None, negatives=None, neg_table=None):
    """{anchor: (sorted distinct edges, sorted node ids, sorted pos_edges[, sorted hard_neg_edges])}; `roots` must
    cover every anchor, positive and negative.  Anchor = any node with a sampled positive: for nodes with in-edges the
    merge of NodeAnchorBasedLinkPredictionTask.scala:186-209, for directed source-only nodes
    formNeighborhoodForSrcOnlyNodes (NodeAnchorBasedLinkPredictionBaseTask.scala:200-278); both reduce to
    array_distinct(own neighbourhood ++ positives' neighbourhoods) plus the root node.
    User-defined labels (UserDefinedLabelsNodeAnchorBasedLinkPredictionTask.scala:54-581): positives / negatives were
    sampled from their own edge tables (`pos_table` / `neg_table`, directed, duplicates kept) and are hydrated against
    them; negatives are LEFT JOINed (an anchor needs a positive only) and their neighbourhoods merge in too; with
    `negatives` given every value carries a 4th element, the hard_neg_edges."""
    te = tree_to_edges(roots, nbr, fanouts)
    idx = {int(r): i for i, r in enumerate(roots)}
    pos_table = table if pos_table is None else pos_table
    out = {}
    for u, ps in positives.items():
        if not ps:
            continue
        e, n = _neighborhood(u, te[idx[u]], table)
        for p_ in ps:
            pe, pn = _neighborhood(p_, te[idx[p_]], table)
            e += pe
            n |= pn
        pos_edges = [(u, p_, f_) for p_ in ps for f_ in pos_table.get((u, p_), [])]  # hydrateTaskBasedEdges :280-334
        if not pos_edges:
            continue
        if negatives is None:
            out[u] = (sorted(set(e)), sorted(n), sorted(pos_edges))
            continue
        neg_edges = []
        for q_ in negatives.get(u, []):
            qe, qn = _neighborhood(q_, te[idx[q_]], table)
            e += qe
            n |= qn
            neg_edges += [(u, q_, f_) for f_ in neg_table.get((u, q_), [])]
        out[u] = (sorted(set(e)), sorted(n), sorted(pos_edges), sorted(neg_edges))
    return out


def np_assemble_dag_rnn(roots, root_node_type, ops):
    """{root: (sorted distinct typed edges [(edge_type, src, dst)], sorted distinct typed nodes [(node_type, id)])}:
    the union of the ops' edge / node sets plus the root (GraphDBSampler.scala:129-148).  ops: dicts with parent,
    fanout, condensed_edge_type, result_node_type, outgoing, nbr (padded tree), in topological order."""
    out = {}
    width = []
    for o in ops:
        width.append((1 if o["parent"] < 0 else width[o["parent"]]) * int(o["fanout"]))
    for r, root in enumerate(roots):
        nodes, edges = {(int(root_node_type), int(root))}, set()
        for o, w in zip(ops, width):
            nbr = np.asarray(o["nbr"])
            for s_ in range(r * w, (r + 1) * w):
                c = int(nbr[s_])
                if c < 0:
                    continue
                par = int(root) if o["parent"] < 0 else int(np.asarray(ops[o["parent"]]["nbr"])[s_ // int(o["fanout"])])
                nodes.add((int(o["result_node_type"]), c))
                edges.add((int(o["condensed_edge_type"]), par, c) if o.get("outgoing") else (int(o["condensed_edge_type"]), c, par))
        out[int(root)] = (sorted(edges), sorted(nodes))
    return out


def np_typed_edge_records(tables):
    """{condensed edge type: (src, dst, feat or None)} -> {type: {(src, dst): [feature tuple (or None) per record, input order]}}:
    the hydrated edge view keyed the way SGSTask.hydrateRnn joins it (_from, _to, _condensed_edge_type)."""
    out = {}
    for t, (src, dst, feat) in tables.items():
        d = {}
        for i, (a, b) in enumerate(zip(src, dst)):
            d.setdefault((int(a), int(b)), []).append(None if feat is None else tuple(np.float32(feat[i]).tolist()))
        out[int(t)] = d
    return out


def _typed_join(key, records, all_records):
    """LEFT JOIN of one (type, src, dst) key with the records of its type: [(type, src, dst, feat or None)]."""
    t, a, b = key
    recs = records[t].get((a, b)) if records is not None and t in records else None
    if not recs:
        return [(t, a, b, None)]
    return [(t, a, b, f) for f in (recs if all_records else recs[:1])]


def np_hydrate_typed_rnn(sets, records):
    """np_assemble_dag_rnn output -> {root: (hydrated edges [(type, src, dst, feat)], nodes)}: collect_list over the LEFT
    JOIN with the hydrated edge view = one Edge per matching record (SGSTask.scala:243-262)."""
    return {r: (sorted((e for k in edges for e in _typed_join(k, records, True)), key=lambda e: (e[0], e[1], e[2], e[3] or ())), nodes)
            for r, (edges, nodes) in sets.items()}


def np_assemble_typed_nablp(anchor_sets, target_sets, target_node_type, pos, pos_edge_type, records=None, hydrate_edges=True,
                            hydrate_pos=True, include_isolated=False):
    """The typed task's NodeAnchorBasedLinkPredictionSample (GraphDBNodeAnchorBasedLinkPredictionTask.scala:283-470):
    anchor_sets / target_sets = np_assemble_dag_rnn of the anchors / of the positives' node type (keyed by node id), pos
    {anchor: [positive ids]}.  pos_edges = the SET of (anchor -> positive) edges, LEFT JOINed with the edge records (one
    per record); neighbourhood = mergeGraphs(anchor's, positives') = distinct by key (GraphPbWrappers.scala:43-68), one
    record per key.  Anchors without a positive are dropped unless include_isolated.
    Returns {anchor: (pos_edges, edges, nodes)}."""
    out = {}
    for a, (edges, nodes) in anchor_sets.items():
        plist = sorted({int(p) for p in pos.get(a, []) if p >= 0})
        if not plist and not include_isolated:
            continue
        eset, nset = set(edges), set(nodes)
        for p_ in plist:
            if p_ in target_sets:
                eset |= set(target_sets[p_][0])
                nset |= set(target_sets[p_][1])
            else:
                nset.add((int(target_node_type), p_))
        pe = [e for p_ in plist for e in _typed_join((int(pos_edge_type), int(a), p_), records if hydrate_pos else None, True)]
        ee = [e for k in sorted(eset) for e in _typed_join(k, records if hydrate_edges else None, False)]
        out[int(a)] = (sorted(pe, key=lambda e: (e[0], e[1], e[2], e[3] or ())), ee, sorted(nset))
    return out


# ----------------------------------------------------------------------------------------
# aggregate
# ----------------------------------------------------------------------------------------
def _coo(edge_index):
    ei = np.asarray(edge_index)
    return np.ascontiguousarray(ei[0], dtype=np.int64), np.ascontiguousarray(ei[1], dtype=np.int64)


def c_sage_conv(x, edge_index, Wl, bl, Wr, relu=False, f64=False, n_threads=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    src, dst = _coo(edge_index)
    Wl = np.ascontiguousarray(Wl, dtype=np.float32)
    Wr = np.ascontiguousarray(Wr, dtype=np.float32)
    bl = None if bl is None else np.ascontiguousarray(bl, dtype=np.float32)
    n, F = x.shape
    O = Wl.shape[0]
    out = np.empty((n, O), dtype=np.float64 if f64 else np.float32)
    fn = lib().oracle_sage_conv_f64 if f64 else lib().oracle_sage_conv_f32
    rc = fn(n, len(src), F, O, _ptr(src), _ptr(dst), _ptr(x), _ptr(Wl), _ptr(bl), _ptr(Wr), _ptr(out),
            int(relu), n_threads or (os.cpu_count() or 1))
    if rc != 0:
        raise RuntimeError(f"oracle_sage_conv failed rc={rc}")
    return out


def np_sage_conv(x, edge_index, Wl, bl, Wr, relu=False):
    """fp64 numpy restatement (np.add.at == index_add_)."""
    x = np.asarray(x, dtype=np.float64)
    src, dst = _coo(edge_index)
    n = x.shape[0]
    agg = np.zeros_like(x)
    np.add.at(agg, dst, x[src])
    c = np.bincount(dst, minlength=n).clip(min=1)[:, None]
    out = (agg / c) @ np.asarray(Wl, np.float64).T + x @ np.asarray(Wr, np.float64).T
    if bl is not None:
        out = out + np.asarray(bl, np.float64)
    return np.maximum(out, 0) if relu else out


def c_gcn_conv(x, edge_index, W, b, relu=False, f64=False):
    x = np.ascontiguousarray(x, dtype=np.float32)
    src, dst = _coo(edge_index)
    W = np.ascontiguousarray(W, dtype=np.float32)
    b = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
    n, F = x.shape
    O = W.shape[0]
    out = np.empty((n, O), dtype=np.float64 if f64 else np.float32)
    fn = lib().oracle_gcn_conv_f64 if f64 else lib().oracle_gcn_conv_f32
    rc = fn(n, len(src), F, O, _ptr(src), _ptr(dst), _ptr(x), _ptr(W), _ptr(b), _ptr(out), int(relu))
    if rc != 0:
        raise RuntimeError(f"oracle_gcn_conv failed rc={rc}")
    return out


def np_gcn_conv(x, edge_index, W, b, relu=False):
    x = np.asarray(x, dtype=np.float64)
    src, dst = _coo(edge_index)
    n = x.shape[0]
    keep = src != dst
    src = np.concatenate([src[keep], np.arange(n)])
    dst = np.concatenate([dst[keep], np.arange(n)])
    deg = np.bincount(dst, minlength=n).astype(np.float64)
    dis = deg**-0.5
    xp = x @ np.asarray(W, np.float64).T
    out = np.zeros((n, xp.shape[1]))
    np.add.at(out, dst, (dis[src] * dis[dst])[:, None] * xp[src])
    if b is not None:
        out = out + np.asarray(b, np.float64)
    return np.maximum(out, 0) if relu else out


def sage_model(x, edge_index, layers, f64=False):
    """GraphSAGE (BasicGNN) forward: relu between layers, none after the last.
    ``layers`` = [(Wl, bl, Wr), ...]  (graphsage_template_modeling_spec.py:143-148)."""
    h = x
    for li, (Wl, bl, Wr) in enumerate(layers):
        last = li == len(layers) - 1
        h = c_sage_conv(np.asarray(h, np.float32), edge_index, Wl, bl, Wr, relu=not last, f64=f64)
    return h


# ----------------------------------------------------------------------------------------
# batch collation (what `x` / `edge_index` mean for the aggregate)
# ----------------------------------------------------------------------------------------
def np_collate(roots, nbr, fanouts):
    """Union of the B sampled subgraphs as the reference's GraphBuilder / collate fns build it
    (python/gigl/src/common/graph_builder/abstract_graph_builder.py:49-197, pyg_graph_builder.py:20-69,
    training/v1/lib/data_loaders/rooted_node_neighborhood_data_loader.py:78-): nodes de-duplicated
    by id, edges de-duplicated by (src, dst).

    -> (node_ids int64 [n] with the roots first (first occurrence order), edge_index int64 [2, e] in
    local ids sorted by (dst, src) global id, root_index int64 [B])."""
    roots = np.asarray(roots, dtype=np.int64)
    parents = roots
    keys = []
    for h, f in enumerate(fanouts):
        ch = np.asarray(nbr[h], dtype=np.int64)
        par = np.repeat(parents, int(f))
        ok = ch >= 0
        keys.append((par[ok] << 32) | ch[ok])
        parents = ch
    keys = np.unique(np.concatenate(keys)) if keys else np.zeros(0, np.int64)
    dst, src = keys >> 32, keys & 0xFFFFFFFF
    seen = {}
    for r in roots.tolist():
        seen.setdefault(r, len(seen))
    rest = np.setdiff1d(np.unique(np.concatenate([dst, src])), roots)
    node_ids = np.concatenate([np.fromiter(seen.keys(), dtype=np.int64, count=len(seen)), rest])
    lut = {int(v): i for i, v in enumerate(node_ids.tolist())}
    loc = np.vectorize(lut.__getitem__, otypes=[np.int64])
    ei = np.stack([loc(src), loc(dst)]) if len(keys) else np.zeros((2, 0), np.int64)
    root_index = loc(roots) if len(roots) else np.zeros(0, np.int64)
    return node_ids, ei, root_index


def batch_sage_embeddings(x_global, roots, nbr, fanouts, layers, f64=False, n_graph_nodes=None):
    """Reference inference on one batch: collate, run GraphSAGE on the WHOLE batch graph, select
    the root rows (graphsage_template_modeling_spec.py:565-577).  n_graph_nodes: collate with the C + OpenMP
    oracle_collate instead of the python dictionaries (same node / edge sets; batches of thousands of roots)."""
    node_ids, ei, root_index = np_collate(roots, nbr, fanouts) if n_graph_nodes is None else c_collate(n_graph_nodes, roots, nbr, fanouts)
    xb = np.ascontiguousarray(np.asarray(x_global)[node_ids], dtype=np.float32)
    out = sage_model(xb, ei, layers, f64=f64)
    return out[root_index]


def np_collate_fast(roots, nbr, fanouts):
    """Vectorised np_collate (same result up to the order of the non-root nodes, which here is
    ascending id); used by the timed CPU baseline where the python-dict version would dominate."""
    roots = np.asarray(roots, dtype=np.int64)
    parents = roots
    keys = []
    for h, f in enumerate(fanouts):
        ch = np.asarray(nbr[h], dtype=np.int64)
        par = np.repeat(parents, int(f))
        ok = ch >= 0
        keys.append((par[ok] << 32) | ch[ok])
        parents = ch
    keys = np.unique(np.concatenate(keys)) if keys else np.zeros(0, np.int64)
    dst, src = keys >> 32, keys & 0xFFFFFFFF
    uroots, first = np.unique(roots, return_index=True)
    uroots = uroots[np.argsort(first, kind="stable")]  # first-occurrence order
    rest = np.setdiff1d(np.unique(np.concatenate([dst, src])), uroots, assume_unique=True)
    node_ids = np.concatenate([uroots, rest])
    order = np.argsort(node_ids, kind="stable")
    sorted_ids = node_ids[order]

    def loc(v):
        return order[np.searchsorted(sorted_ids, v)]

    ei = np.stack([loc(src), loc(dst)]) if len(keys) else np.zeros((2, 0), np.int64)
    return node_ids, ei, loc(roots) if len(roots) else np.zeros(0, np.int64)


def c_collate(n_graph_nodes, roots, nbr, fanouts, n_threads=None):
    """np_collate_fast in C + OpenMP (oracle_collate): the collation of the TIMED CPU baseline, every host core."""
    roots = np.ascontiguousarray(roots, dtype=np.int32)
    fan = np.ascontiguousarray(fanouts, dtype=np.int32)
    levels = [np.ascontiguousarray(t, dtype=np.int32) for t in nbr]
    slots = int(sum(t.size for t in levels))
    node_ids = np.empty(len(roots) + slots, dtype=np.int64)
    es = np.empty(max(slots, 1), dtype=np.int64)
    ed = np.empty(max(slots, 1), dtype=np.int64)
    ridx = np.empty(max(len(roots), 1), dtype=np.int64)
    nn, ne = C.c_int64(), C.c_int64()
    ptrs = (C.c_void_p * len(levels))(*[t.ctypes.data for t in levels])
    rc = lib().oracle_collate(int(n_graph_nodes), _ptr(roots), len(roots), _ptr(fan), len(fan), ptrs, _ptr(node_ids), C.byref(nn),
                              _ptr(es), _ptr(ed), C.byref(ne), _ptr(ridx), int(n_threads or 0))
    if rc != 0:
        raise ValueError("oracle_collate: bad argument")
    return node_ids[:nn.value], np.stack([es[:ne.value], ed[:ne.value]]), ridx[:len(roots)]


def torch_sage_forward(x, edge_index, layers, n_threads=None):
    """The reference's aggregate on the CPU with torch itself: PyG 2.5.3 SAGEConv(mean) restated as
    index_add_ + F.linear, GraphSAGE/BasicGNN layer loop with ReLU between
    layers (homogeneous.py:107-153, graphsage_template_modeling_spec.py:143-148).  This is the
    timed CPU baseline of the aggregate half: torch's own multi-threaded
    CPU kernels, all host threads."""
    import torch
    import torch.nn.functional as F

    if n_threads:
        torch.set_num_threads(int(n_threads))
    with torch.no_grad():
        h = torch.as_tensor(x, dtype=torch.float32)
        ei = torch.as_tensor(edge_index, dtype=torch.int64)
        src, dst = ei[0], ei[1]
        n = h.shape[0]
        cnt = torch.zeros(n, dtype=torch.float32).index_add_(0, dst, torch.ones(dst.numel(), dtype=torch.float32))
        inv = 1.0 / cnt.clamp(min=1.0)
        for li, (Wl, bl, Wr) in enumerate(layers):
            agg = torch.zeros_like(h).index_add_(0, dst, h.index_select(0, src)) * inv[:, None]
            out = F.linear(agg, torch.as_tensor(Wl), None if bl is None else torch.as_tensor(bl)) + F.linear(h, torch.as_tensor(Wr))
            h = out.relu_() if li < len(layers) - 1 else out
        return h.numpy()


def torch_build_in_csr(src, dst, n_nodes: int, is_graph_directed: bool):
    """np_build_in_csr with torch tensor ops (any device): used to prepare BASELINE-size inputs for
    the CPU reference arm without going through the product library."""
    import torch

    src = src.to(torch.int64)
    dst = dst.to(torch.int64)
    if not is_graph_directed:
        lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
        pairs = torch.unique((lo << 32) | hi)
        lo, hi = pairs >> 32, pairs & 0xFFFFFFFF
        both = torch.unique(torch.cat([(hi << 32) | lo, (lo << 32) | hi]))  # key = dst << 32 | src, UNION DISTINCT
    else:
        both, _ = torch.sort((dst << 32) | src)
    d, s = both >> 32, both & 0xFFFFFFFF
    rowptr = torch.zeros(n_nodes + 1, dtype=torch.int64, device=src.device)
    rowptr[1:] = torch.cumsum(torch.bincount(d, minlength=n_nodes), 0)
    return rowptr, s.to(torch.int32)


# ----------------------------------------------------------------------------------------------
# Training reference: the restated layers under torch CPU autograd (what `loss.backward()` does in
# node_classification_modeling_task_spec.py:134-173 / graphsage_template_modeling_spec.py:299-367).
# ----------------------------------------------------------------------------------------------
def torch_sage_grads(x, edge_index, layers, grad_out=None, level_sizes=None, f64=True, x_requires_grad=True):
    """GraphSAGE forward + backward on the CPU in fp64 (or fp32).  ``grad_out`` is d(loss)/d(output) (defaults to ones).
    Returns (out, grad_x, [(gWl, gbl, gWr), ...]).  ``level_sizes`` prunes like gigl_b200.nn.GraphSAGE (same numbers on
    the kept rows)."""
    import torch
    import torch.nn.functional as F

    dt = torch.float64 if f64 else torch.float32
    h = torch.tensor(np.asarray(x), dtype=dt, requires_grad=x_requires_grad)
    x0 = h
    ei = torch.as_tensor(np.asarray(edge_index), dtype=torch.int64)
    src, dst = ei[0], ei[1]
    n = h.shape[0]
    cnt = torch.zeros(n, dtype=dt).index_add_(0, dst, torch.ones(dst.numel(), dtype=dt))
    inv = 1.0 / cnt.clamp(min=1.0)
    params = []
    L = len(layers)
    for li, (Wl, bl, Wr) in enumerate(layers):
        Wl_t = torch.tensor(np.asarray(Wl), dtype=dt, requires_grad=True)
        Wr_t = torch.tensor(np.asarray(Wr), dtype=dt, requires_grad=True)
        bl_t = None if bl is None else torch.tensor(np.asarray(bl), dtype=dt, requires_grad=True)
        params.append((Wl_t, bl_t, Wr_t))
        n_in = h.shape[0]
        keep = (src < n_in) & (dst < n_in)
        s, d = src[keep], dst[keep]
        agg = torch.zeros((n_in, h.shape[1]), dtype=dt).index_add(0, d, h.index_select(0, s)) * inv[:n_in, None]
        out = F.linear(agg, Wl_t, bl_t) + F.linear(h, Wr_t)
        if li < L - 1:
            out = out.relu()
        if level_sizes is not None:
            out = out[: int(level_sizes[L - 1 - li])]
        h = out
    g = torch.ones_like(h) if grad_out is None else torch.as_tensor(np.asarray(grad_out), dtype=dt)
    h.backward(g)
    grads = [(p[0].grad.numpy(), None if p[1] is None else p[1].grad.numpy(), p[2].grad.numpy()) for p in params]
    return h.detach().numpy(), (x0.grad.numpy() if x_requires_grad else None), grads


def torch_gcn_grads(x, edge_index, W, b, relu=False, grad_out=None, f64=True):
    """GCNConv forward + backward under torch CPU autograd: (out, grad_x, grad_W, grad_b)."""
    import torch

    dt = torch.float64 if f64 else torch.float32
    xt = torch.tensor(np.asarray(x), dtype=dt, requires_grad=True)
    Wt = torch.tensor(np.asarray(W), dtype=dt, requires_grad=True)
    bt = None if b is None else torch.tensor(np.asarray(b), dtype=dt, requires_grad=True)
    ei = torch.as_tensor(np.asarray(edge_index), dtype=torch.int64)
    n = xt.shape[0]
    keep = ei[0] != ei[1]
    loops = torch.arange(n, dtype=torch.int64)
    src = torch.cat([ei[0][keep], loops])
    dst = torch.cat([ei[1][keep], loops])
    deg = torch.zeros(n, dtype=dt).index_add_(0, dst, torch.ones(dst.numel(), dtype=dt))
    dinv = deg.pow(-0.5)
    w = dinv[src] * dinv[dst]
    xp = xt @ Wt.t()
    out = torch.zeros((n, Wt.shape[0]), dtype=dt).index_add(0, dst, xp.index_select(0, src) * w[:, None])
    if bt is not None:
        out = out + bt
    if relu:
        out = out.relu()
    g = torch.ones_like(out) if grad_out is None else torch.as_tensor(np.asarray(grad_out), dtype=dt)
    out.backward(g)
    return out.detach().numpy(), xt.grad.numpy(), Wt.grad.numpy(), None if bt is None else bt.grad.numpy()
