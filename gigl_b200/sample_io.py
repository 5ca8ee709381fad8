"""Host mirror of the sampler's file contract (include/gigl_b200.h, "the sampler's file contract"):
TFRecord files of tf.Example rows in, TFRecord files of serialized sample protos out.

Reference: scala/common/src/main/scala/utils/TFRecordIO.scala:21-69 (read / write),
SGSPureSparkV1Task.scala:52-311 (node / edge table loading), :496-820 and :1019-1040 (hydration, proto schema).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import GiglError

INT32_MIN = -(2 ** 31)


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise GiglError(rc, what)


def crc32c_masked(data: bytes) -> int:
    return int(_capi.lib().gigl_crc32c_masked(data, len(data)))


def list_tfrecord_files(uri_prefix: str) -> List[str]:
    """A `tfrecordUriPrefix` names a directory (or a file-name prefix) of TFRecord part files."""
    p = uri_prefix[len("file://"):] if uri_prefix.startswith("file://") else uri_prefix
    if os.path.isdir(p):
        files = sorted(f for f in glob.glob(os.path.join(p, "*")) if os.path.isfile(f) and not os.path.basename(f).startswith(("_", ".")))
    else:
        files = sorted(f for f in glob.glob(p + "*") if os.path.isfile(f))
    if not files:
        raise FileNotFoundError(f"no TFRecord files under {uri_prefix!r}")
    return files


def validate_samples(data, kind: str, edge_node_types=None) -> int:
    """The reference's TaskOutputValidator (scala/subgraph_sampler/src/main/scala/libs/task/TaskOutputValidator.scala:29-108)
    on a TFRecord stream of encoded samples, natively (gigl_validate_samples_host): both endpoints of every neighbourhood
    edge - for kind "nablp" also of every pos / neg / hard-neg edge - must be among the neighbourhood nodes as
    (node id, condensed node type).  edge_node_types: {condensed edge type: (src condensed node type, dst condensed node
    type)} or None (homogeneous).  data: bytes or the encoder's NativeBuffer.  Returns the number of records; raises
    GiglError naming the first offending record (the reference throws RuntimeException)."""
    L = _capi.lib()
    if isinstance(data, NativeBuffer):
        ptr, n = data._ptr, data.nbytes
        keep = data
    else:
        keep = np.frombuffer(data, dtype=np.uint8)
        ptr, n = keep.ctypes.data, len(keep)
    st = dt = None
    n_types = 0
    if edge_node_types:
        n_types = max(edge_node_types) + 1
        st = np.zeros(n_types, dtype=np.int32)
        dt = np.zeros(n_types, dtype=np.int32)
        for t, (a, b) in edge_node_types.items():
            st[t], dt[t] = a, b
    n_rec, bad, why = C.c_int64(), C.c_int64(-1), C.c_int32(0)
    rc = L.gigl_validate_samples_host(C.c_void_p(ptr), n, {"rnn": 0, "snc": 1, "nablp": 2}[kind], n_types,
                                      None if st is None else st.ctypes.data, None if dt is None else dt.ctypes.data,
                                      C.byref(n_rec), C.byref(bad), C.byref(why))
    del keep
    if rc != 0:
        reason = {1: "malformed sample bytes", 2: "neighborhood not present in sample",
                  3: "a node of an edge is not present in the neighborhood graph"}.get(why.value, "invalid stream")
        raise GiglError(int(rc), f"Output Validation failed: record {bad.value}: {reason}")
    return int(n_rec.value)


class ExampleTable:
    """All tf.Example records of a set of TFRecord files, decoded column by column in native code."""

    def __init__(self, data: bytes, verify: bool = True):
        self._L = _capi.lib()
        self.data = np.frombuffer(data, dtype=np.uint8)
        n = self._L.gigl_tfrecord_index_host(self.data.ctypes.data, len(self.data), int(verify), None, None, 0)
        if n < 0:
            raise GiglError(int(n), "malformed TFRecord stream (framing or crc32c)")
        self.n = int(n)
        self.offsets = np.zeros(max(self.n, 1), dtype=np.int64)
        self.lengths = np.zeros(max(self.n, 1), dtype=np.int64)
        if self.n:
            m = self._L.gigl_tfrecord_index_host(self.data.ctypes.data, len(self.data), 0, self.offsets.ctypes.data,
                                                 self.lengths.ctypes.data, self.n)
            assert m == self.n

    @classmethod
    def from_files(cls, files: Sequence[str], verify: bool = True) -> "ExampleTable":
        return cls(b"".join(open(f, "rb").read() for f in files), verify)

    def record(self, i: int) -> bytes:
        o, l = int(self.offsets[i]), int(self.lengths[i])
        return self.data[o:o + l].tobytes()

    def width(self, name: str) -> int:
        """Values per record of feature `name` (probed on the first record)."""
        if self.n == 0:
            return 1
        feats = parse_example(self.record(0))
        if name not in feats:
            raise KeyError(f"feature {name!r} not in the tf.Example records (have {sorted(feats)})")
        return max(len(feats[name]), 1)

    def column(self, name: str, dtype: str, width: Optional[int] = None) -> np.ndarray:
        """dtype 'int64' or 'float32' -> array [n] (width 1) or [n, width]."""
        w = self.width(name) if width is None else width
        if dtype == "int64":
            out = np.zeros((self.n, w), dtype=np.int64)
            rc = self._L.gigl_examples_column_host(self.data.ctypes.data, self.n, self.offsets.ctypes.data, self.lengths.ctypes.data,
                                                   name.encode(), 0, w, out.ctypes.data, None)
        else:
            out = np.zeros((self.n, w), dtype=np.float32)
            rc = self._L.gigl_examples_column_host(self.data.ctypes.data, self.n, self.offsets.ctypes.data, self.lengths.ctypes.data,
                                                   name.encode(), 1, w, None, out.ctypes.data)
        _check(rc, f"decoding tf.Example feature {name!r}")
        return out[:, 0] if w == 1 and width is None else out


_KINDS = {"rnn": 0, "snc": 1, "nablp": 2}


class NativeBuffer:
    """The encoder's malloc'd output, handed over WITHOUT a copy (`zero_copy=True` of the encode_* functions): a
    bytes-like object (`memoryview(buf)`, `file.write(buf.view)`, `len(buf)`) that frees the native memory when it is
    closed or collected.  Copying 1.2 GB into a Python `bytes` costs seven times the encoding itself."""

    def __init__(self, ptr: int, nbytes: int):
        self._ptr, self.nbytes = ptr, int(nbytes)
        self._arr = (C.c_ubyte * max(self.nbytes, 1)).from_address(ptr) if ptr else None

    @property
    def view(self) -> memoryview:
        if self._arr is None:
            raise ValueError("buffer is closed")
        return memoryview(self._arr).cast("B")[: self.nbytes]

    def __len__(self) -> int:
        return self.nbytes

    def __bytes__(self) -> bytes:
        return bytes(self.view)

    def close(self) -> None:
        if self._ptr:
            self._arr = None
            _capi.lib().gigl_free_host(C.c_void_p(self._ptr))
            self._ptr = 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _take(out: "C.c_void_p", nbytes: int, zero_copy: bool):
    """The encoder's output buffer as `bytes` (copied, then freed) or as a :class:`NativeBuffer` (no copy)."""
    if zero_copy:
        return NativeBuffer(out.value or 0, nbytes)
    try:
        return C.string_at(out.value, nbytes)
    finally:
        _capi.lib().gigl_free_host(out)


def encode_samples(roots, fanouts, nbr, x: Optional[np.ndarray] = None, kind: str = "rnn", condensed_node_type: int = 0,
                   condensed_edge_type: int = 0, labels: Optional[np.ndarray] = None, label_type: str = "",
                   tfrecord_framing: bool = True, csr: Optional[Tuple[np.ndarray, np.ndarray]] = None,
                   edge_rows: Optional[np.ndarray] = None, edge_feat: Optional[np.ndarray] = None, n_emit: Optional[int] = None,
                   pos: Optional[np.ndarray] = None, pos_tree: Optional[np.ndarray] = None, zero_copy: bool = False) -> Tuple[bytes, np.ndarray]:
    """Padded-tree index sets -> serialized RootedNodeNeighborhood ('rnn'), SupervisedNodeClassificationSample ('snc') or
    NodeAnchorBasedLinkPredictionSample ('nablp') messages, one per emitting root (include/gigl_b200.h,
    gigl_encode_samples_ex_host).  `csr` = host (rowptr, col) of the in-CSR turns every sampled pair into one Edge per
    matching edge record (the reference's join) and, with `edge_feat` [records, Fe] + `edge_rows`, hydrates edge
    features.  'nablp': the first `n_emit` roots are anchors with positives `pos` [n_emit, num_pos] (-1 = none) whose
    trees sit at `pos_tree` (indices into roots).  Returns (bytes, record_offsets [n_emit + 1])."""
    L = _capi.lib()
    roots = np.ascontiguousarray(roots, dtype=np.int32)
    fan = np.ascontiguousarray(fanouts, dtype=np.int32)
    nbr = [np.ascontiguousarray(a, dtype=np.int32) for a in nbr]
    n_emit = len(roots) if n_emit is None else int(n_emit)
    F = 0
    xp = None
    if x is not None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        F = x.shape[1]
        xp = x.ctypes.data
    lp = None
    if kind == "snc":
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        lp = labels.ctypes.data
    rp = cl = er = efp = None
    Fe = 0
    if csr is not None:
        rowptr = np.ascontiguousarray(csr[0], dtype=np.int64)
        col = np.ascontiguousarray(csr[1], dtype=np.int32)
        rp, cl = rowptr.ctypes.data, col.ctypes.data
        if edge_feat is not None and edge_feat.size:
            edge_feat = np.ascontiguousarray(edge_feat, dtype=np.float32)
            Fe = edge_feat.shape[1]
            efp = edge_feat.ctypes.data
            if edge_rows is not None:
                edge_rows = np.ascontiguousarray(edge_rows, dtype=np.int32)
                er = edge_rows.ctypes.data
    num_pos = 0
    pp = ptp = None
    if kind == "nablp":
        pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(n_emit, -1)
        pos_tree = np.ascontiguousarray(pos_tree, dtype=np.int64).reshape(n_emit, -1)
        num_pos = pos.shape[1]
        pp, ptp = pos.ctypes.data, pos_tree.ctypes.data
    pn = (C.c_void_p * len(fan))(*[a.ctypes.data for a in nbr])
    out = C.c_void_p()
    nbytes = C.c_int64()
    offs = np.zeros(n_emit + 1, dtype=np.int64)
    rc = L.gigl_encode_samples_ex_host(_KINDS[kind], len(roots), n_emit, roots.ctypes.data, fan.ctypes.data, len(fan), pn, xp, F,
                                       condensed_node_type, condensed_edge_type, rp, cl, er, efp, Fe, lp, label_type.encode(),
                                       num_pos, pp, ptp, int(tfrecord_framing), C.byref(out), C.byref(nbytes), offs.ctypes.data)
    _check(rc, "gigl_encode_samples_ex_host")
    return _take(out, nbytes.value, zero_copy), offs


class HostEdgeTable:
    """Host arrays of one hydrated edge table (main / user-defined positive / negative) for the sample encoder:
    in-CSR by destination + the input record behind every CSR slot + the records' feature rows."""

    def __init__(self, csr: Optional[Tuple[np.ndarray, np.ndarray]] = None, edge_rows: Optional[np.ndarray] = None,
                 edge_feat: Optional[np.ndarray] = None):
        self.rowptr = self.col = self.edge_rows = self.feat = None
        if csr is not None:
            self.rowptr = np.ascontiguousarray(csr[0], dtype=np.int64)
            self.col = np.ascontiguousarray(csr[1], dtype=np.int32)
            if edge_feat is not None and edge_feat.size:
                self.feat = np.ascontiguousarray(edge_feat, dtype=np.float32)
                if edge_rows is not None:
                    self.edge_rows = np.ascontiguousarray(edge_rows, dtype=np.int32)

    def struct(self) -> "_capi.EdgeTable":
        p = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        return _capi.EdgeTable(p(self.rowptr), p(self.col), p(self.edge_rows), p(self.feat), 0 if self.feat is None else self.feat.shape[1])


def encode_link_samples(roots, fanouts, nbr, x: Optional[np.ndarray], n_emit: int, pos, pos_tree, main: HostEdgeTable,
                        pos_table: Optional[HostEdgeTable] = None, neg=None, neg_tree=None, neg_table: Optional[HostEdgeTable] = None,
                        condensed_node_type: int = 0, condensed_edge_type: int = 0, tfrecord_framing: bool = True,
                        zero_copy: bool = False) -> Tuple[bytes, np.ndarray]:
    """NodeAnchorBasedLinkPredictionSample messages with positives (and optional hard negatives) that were sampled
    from - and are hydrated against - their own edge tables (gigl_encode_link_samples_host).  `pos_table` None = the
    main table.  Returns (bytes, record_offsets [n_emit + 1])."""
    L = _capi.lib()
    roots = np.ascontiguousarray(roots, dtype=np.int32)
    fan = np.ascontiguousarray(fanouts, dtype=np.int32)
    nbr = [np.ascontiguousarray(a, dtype=np.int32) for a in nbr]
    F, xp = 0, None
    if x is not None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        F, xp = x.shape[1], x.ctypes.data
    pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(n_emit, -1)
    pos_tree = np.ascontiguousarray(pos_tree, dtype=np.int64).reshape(n_emit, -1)
    num_neg, ngp, ntp = 0, None, None
    if neg is not None:
        neg = np.ascontiguousarray(neg, dtype=np.int32).reshape(n_emit, -1)
        neg_tree = np.ascontiguousarray(neg_tree, dtype=np.int64).reshape(n_emit, -1)
        num_neg, ngp, ntp = neg.shape[1], neg.ctypes.data, neg_tree.ctypes.data
    tm = main.struct()
    tp = pos_table.struct() if pos_table is not None else None
    tn = neg_table.struct() if neg_table is not None else None
    pn = (C.c_void_p * len(fan))(*[a.ctypes.data for a in nbr])
    out = C.c_void_p()
    nbytes = C.c_int64()
    offs = np.zeros(n_emit + 1, dtype=np.int64)
    rc = L.gigl_encode_link_samples_host(len(roots), n_emit, roots.ctypes.data, fan.ctypes.data, len(fan), pn, xp, F, condensed_node_type,
                                         condensed_edge_type, C.addressof(tm), C.addressof(tp) if tp is not None else None,
                                         C.addressof(tn) if tn is not None else None, pos.shape[1], pos.ctypes.data, pos_tree.ctypes.data,
                                         num_neg, ngp, ntp, int(tfrecord_framing), C.byref(out), C.byref(nbytes), offs.ctypes.data)
    _check(rc, "gigl_encode_link_samples_host")
    return _take(out, nbytes.value, zero_copy), offs


def encode_dag_samples(roots, root_node_type: int, ops: Sequence[dict], node_tables: Sequence[Optional[np.ndarray]],
                       tfrecord_framing: bool = True, zero_copy: bool = False) -> Tuple[bytes, np.ndarray]:
    """Typed RootedNodeNeighborhood messages from the ops of a SamplingOp DAG (gigl_encode_dag_samples_host).
    ops (topological order): dicts with parent (index or -1), fanout, condensed_edge_type, result_node_type, outgoing,
    nbr (the op's padded-tree output).  node_tables[t] = feature matrix of condensed node type t (or None)."""
    L = _capi.lib()
    roots = np.ascontiguousarray(roots, dtype=np.int32)
    keep = []
    c_ops = (_capi.DagOp * max(len(ops), 1))()
    for i, o in enumerate(ops):
        nbr = np.ascontiguousarray(o["nbr"], dtype=np.int32)
        keep.append(nbr)
        c_ops[i] = _capi.DagOp(int(o["parent"]), int(o["fanout"]), int(o["condensed_edge_type"]), int(o["result_node_type"]),
                               int(bool(o.get("outgoing", False))), nbr.ctypes.data)
    c_tabs = (_capi.NodeTable * max(len(node_tables), 1))()
    for t, x in enumerate(node_tables):
        if x is not None and x.size:
            x = np.ascontiguousarray(x, dtype=np.float32)
            keep.append(x)
            c_tabs[t] = _capi.NodeTable(x.ctypes.data, x.shape[1])
        else:
            c_tabs[t] = _capi.NodeTable(None, 0)
    out = C.c_void_p()
    nbytes = C.c_int64()
    offs = np.zeros(len(roots) + 1, dtype=np.int64)
    rc = L.gigl_encode_dag_samples_host(len(roots), roots.ctypes.data, int(root_node_type), len(ops), C.addressof(c_ops), len(node_tables),
                                        C.addressof(c_tabs), int(tfrecord_framing), C.byref(out), C.byref(nbytes), offs.ctypes.data)
    _check(rc, "gigl_encode_dag_samples_host")
    return _take(out, nbytes.value, zero_copy), offs


def _dag_tree(roots, root_node_type: int, ops: Sequence[dict], keep: list) -> "_capi.DagTree":
    roots = np.ascontiguousarray(roots, dtype=np.int32)
    c_ops = (_capi.DagOp * max(len(ops), 1))()
    for i, o in enumerate(ops):
        nbr = np.ascontiguousarray(o["nbr"], dtype=np.int32)
        keep.append(nbr)
        c_ops[i] = _capi.DagOp(int(o["parent"]), int(o["fanout"]), int(o["condensed_edge_type"]), int(o["result_node_type"]),
                               int(bool(o.get("outgoing", False))), nbr.ctypes.data)
    keep += [roots, c_ops]
    return _capi.DagTree(len(roots), roots.ctypes.data, int(root_node_type), len(ops), C.addressof(c_ops))


def encode_typed_samples(roots, root_node_type: int, ops: Sequence[dict], node_tables: Sequence[Optional[np.ndarray]],
                         edge_tables: Optional[Sequence[Optional["HostEdgeTable"]]] = None, kind: str = "rnn", pos=None, pos_tree=None,
                         pos_condensed_edge_type: int = -1, target_roots=None, target_node_type: int = 0,
                         target_ops: Optional[Sequence[dict]] = None, include_isolated: bool = False, hydrate_edges: bool = True,
                         hydrate_pos_edges: bool = True, tfrecord_framing: bool = True, zero_copy: bool = False) -> Tuple[bytes, np.ndarray]:
    """Typed RootedNodeNeighborhood (kind "rnn") / NodeAnchorBasedLinkPredictionSample (kind "nablp") messages from sampled
    SamplingOp DAGs, with edge hydration (gigl_encode_typed_samples_host).  ops as in :func:`encode_dag_samples`;
    edge_tables[t] = :class:`HostEdgeTable` of condensed edge type t (None entries / None = not hydrated).  nablp: pos
    [n_roots, num_pos] positive node ids (-1 = none), pos_tree = their index in target_roots (-1 = the node alone),
    target_* = the sampled DAG of the positives' node type."""
    L = _capi.lib()
    keep: list = []
    anchors = _dag_tree(roots, root_node_type, ops, keep)
    n = int(anchors.n_roots)
    c_tabs = (_capi.NodeTable * max(len(node_tables), 1))()
    for t, x in enumerate(node_tables):
        if x is not None and x.size:
            x = np.ascontiguousarray(x, dtype=np.float32)
            keep.append(x)
            c_tabs[t] = _capi.NodeTable(x.ctypes.data, x.shape[1])
        else:
            c_tabs[t] = _capi.NodeTable(None, 0)
    c_et, n_et = None, 0
    if edge_tables is not None:
        n_et = len(edge_tables)
        c_et = (_capi.EdgeTable * max(n_et, 1))()
        for t, tab in enumerate(edge_tables):
            c_et[t] = tab.struct() if tab is not None else _capi.EdgeTable(None, None, None, None, 0)
        keep.append(edge_tables)
    targets, num_pos, pos_p, tree_p = None, 0, None, None
    if kind == "nablp":
        pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(n, -1)
        pos_tree = np.ascontiguousarray(pos_tree, dtype=np.int64).reshape(n, -1)
        num_pos = pos.shape[1]
        targets = _dag_tree(target_roots if target_roots is not None else np.zeros(0, np.int32), target_node_type, target_ops or [], keep)
        pos_p, tree_p = pos.ctypes.data, pos_tree.ctypes.data
    elif kind != "rnn":
        raise ValueError(f"kind must be 'rnn' or 'nablp', got {kind!r}")
    out = C.c_void_p()
    nbytes = C.c_int64()
    offs = np.zeros(n + 1, dtype=np.int64)
    rc = L.gigl_encode_typed_samples_host(2 if kind == "nablp" else 0, C.addressof(anchors), C.addressof(targets) if targets is not None else None,
                                          num_pos, pos_p, tree_p, int(pos_condensed_edge_type), int(include_isolated),
                                          int(bool(hydrate_edges)) | (int(bool(hydrate_pos_edges)) << 1), len(node_tables), C.addressof(c_tabs),
                                          n_et, C.addressof(c_et) if c_et is not None else None, int(tfrecord_framing), C.byref(out),
                                          C.byref(nbytes), offs.ctypes.data)
    _check(rc, "gigl_encode_typed_samples_host")
    return _take(out, nbytes.value, zero_copy), offs


# ---- a minimal protobuf wire reader (tests, tooling): no generated code needed --------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    v = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def parse_fields(buf: bytes) -> List[Tuple[int, int, object]]:
    """[(field number, wire type, value)]: varints as int, length-delimited as bytes, fixed32/64 as bytes."""
    out, pos = [], 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        fn, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError(f"unsupported wire type {wt}")
        out.append((fn, wt, v))
    return out


def parse_example(buf: bytes) -> Dict[str, list]:
    """tf.Example -> {feature name: list of python values}."""
    feats = {}
    for fn, _, features in parse_fields(buf):
        if fn != 1:
            continue
        for _, _, entry in parse_fields(features):
            key, val = None, b""
            for f2, _, v in parse_fields(entry):
                if f2 == 1:
                    key = v.decode()
                elif f2 == 2:
                    val = v
            vals = []
            for kind, _, lst in parse_fields(val):
                for _, wt, v in parse_fields(lst):
                    if kind == 2:  # FloatList
                        vals += list(np.frombuffer(v, dtype="<f4")) if wt == 2 else [float(np.frombuffer(v, dtype="<f4")[0])]
                    elif kind == 3:  # Int64List
                        if wt == 2:
                            p = 0
                            while p < len(v):
                                iv, p = _varint(v, p)
                                vals.append(iv - (1 << 64) if iv >> 63 else iv)
                        else:
                            vals.append(v - (1 << 64) if v >> 63 else v)
                    else:
                        vals.append(v)
            feats[key] = vals
    return feats


def _parse_node(buf: bytes) -> dict:
    d = {"node_id": 0, "condensed_node_type": None, "feature_values": []}
    for fn, wt, v in parse_fields(buf):
        if fn == 1:
            d["node_id"] = v
        elif fn == 2:
            d["condensed_node_type"] = v
        elif fn == 3:
            d["feature_values"] += list(np.frombuffer(v, dtype="<f4")) if wt == 2 else [float(np.frombuffer(v, dtype="<f4")[0])]
    return d


def _parse_edge(buf: bytes) -> dict:
    d = {"src_node_id": 0, "dst_node_id": 0, "condensed_edge_type": None, "feature_values": []}
    for fn, wt, v in parse_fields(buf):
        if fn == 1:
            d["src_node_id"] = v
        elif fn == 2:
            d["dst_node_id"] = v
        elif fn == 3:
            d["condensed_edge_type"] = v
        elif fn == 4:
            d["feature_values"] += list(np.frombuffer(v, dtype="<f4")) if wt == 2 else [float(np.frombuffer(v, dtype="<f4")[0])]
    return d


def parse_sample(buf: bytes) -> dict:
    """RootedNodeNeighborhood / SupervisedNodeClassificationSample bytes -> plain dict."""
    d = {"root_node": None, "nodes": [], "edges": [], "root_node_labels": []}
    for fn, _, v in parse_fields(buf):
        if fn == 1:
            d["root_node"] = _parse_node(v)
        elif fn == 2:
            for f2, _, g in parse_fields(v):
                if f2 == 2:
                    d["nodes"].append(_parse_node(g))
                elif f2 == 3:
                    d["edges"].append(_parse_edge(g))
        elif fn == 3:
            lab = {"label_type": "", "label": 0}
            for f2, _, g in parse_fields(v):
                if f2 == 1:
                    lab["label_type"] = g.decode()
                elif f2 == 2:
                    lab["label"] = g - (1 << 64) if g >> 63 else g
            d["root_node_labels"].append(lab)
    return d


def parse_nablp_sample(buf: bytes) -> dict:
    """NodeAnchorBasedLinkPredictionSample bytes -> plain dict (training_samples_schema.proto:34-53)."""
    d = {"root_node": None, "nodes": [], "edges": [], "pos_edges": [], "hard_neg_edges": [], "neg_edges": []}
    for fn, _, v in parse_fields(buf):
        if fn == 1:
            d["root_node"] = _parse_node(v)
        elif fn == 3:
            for f2, _, g in parse_fields(v):
                if f2 == 2:
                    d["nodes"].append(_parse_node(g))
                elif f2 == 3:
                    d["edges"].append(_parse_edge(g))
        elif fn in (2, 4, 5):
            d[{2: "hard_neg_edges", 4: "pos_edges", 5: "neg_edges"}[fn]].append(_parse_edge(v))
    return d


def split_tfrecords(data: bytes, verify: bool = True) -> List[bytes]:
    t = ExampleTable(data, verify)
    return [t.record(i) for i in range(t.n)]
