"""Thin object layer over the C-ABI: one :class:`Context` per GPU, :class:`Graph` = CSR resident in HBM.

Host-buffer methods (``*_host``) take numpy arrays and go through the ``gigl_*_host`` entry points
(what a JNI binding would call; H2D/D2H inside).  Device methods take torch CUDA tensors and only
pass their ``data_ptr()`` to the ``gigl_*_dev`` entry points - torch is used for device memory
and streams, never for arithmetic.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import GiglError, check  # noqa: F401


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _hp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dp(t):
    """device pointer of a torch CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())


# SamplingOp.sampling_method (subgraph_sampling_strategy.proto:38-58) -> GIGL_SAMPLE_* of include/gigl_b200.h
SAMPLING_METHODS = {"random_uniform": 0, "top_k": 1, "random_weighted": 2}


class Context:
    """gigl_ctx: device + stream + scratch.  Not thread-safe; use one per thread/GPU
    (mirrors the per-partition setup()/teardown() of the reference's KHopSamplerService,
    scala_spark35/common/src/main/scala/graphdb/KHopSamplerService.scala:10-33)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._L = _capi.lib()
        h = C.c_void_p()
        if stream is None:
            rc = self._L.gigl_ctx_create(device, C.byref(h))
        else:
            rc = self._L.gigl_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h))
        check(rc, None)
        self.handle = h
        self.device = device

    @classmethod
    def on_torch_stream(cls, device: int = 0) -> "Context":
        """Enqueue on torch's current stream of ``device`` so torch tensors and events order naturally."""
        import torch

        return cls(device, stream=torch.cuda.current_stream(device).cuda_stream)

    def sync(self) -> None:
        check(self._L.gigl_ctx_sync(self.handle), self.handle)

    @property
    def launch_count(self) -> int:
        return int(self._L.gigl_ctx_launch_count(self.handle))

    @property
    def stream(self) -> int:
        return int(self._L.gigl_ctx_stream(self.handle) or 0)

    # ---- phase timing ----------------------------------------------------------------------
    def set_timing(self, enabled: bool) -> None:
        check(self._L.gigl_ctx_set_timing(self.handle, int(enabled)), self.handle)

    def reset_timing(self) -> None:
        check(self._L.gigl_ctx_reset_timing(self.handle), self.handle)

    def timings(self) -> dict:
        """{tag: (total device ms, launches of that phase)} accumulated since the last reset (synchronises)."""
        out = {}
        for t in range(self._L.gigl_timing_num_tags()):
            ms, n = C.c_double(), C.c_int64()
            check(self._L.gigl_ctx_get_timing(self.handle, t, C.byref(ms), C.byref(n)), self.handle)
            if n.value:
                out[self._L.gigl_timing_tag_name(t).decode()] = (ms.value, n.value)
        return out

    def close(self) -> None:
        if self.handle:
            self._L.gigl_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- aggregate, host buffers ---------------------------------------------------------
    def sage_conv_host(self, x, edge_index, Wl, bl, Wr, relu: bool = False) -> np.ndarray:
        x = _np(x, np.float32)
        ei = _np(edge_index, np.int64)
        Wl, Wr = _np(Wl, np.float32), _np(Wr, np.float32)
        bl = None if bl is None else _np(bl, np.float32)
        n, F = x.shape
        O = Wl.shape[0]
        assert ei.ndim == 2 and ei.shape[0] == 2 and Wl.shape == (O, F) and Wr.shape == (O, F)
        out = np.empty((n, O), dtype=np.float32)
        check(self._L.gigl_sage_conv_host(self.handle, n, ei.shape[1], F, O, _hp(ei), _hp(x), _hp(Wl), _hp(bl),
                                          _hp(Wr), _hp(out), int(relu)), self.handle)
        return out

    def gcn_conv_host(self, x, edge_index, W, b, relu: bool = False) -> np.ndarray:
        x = _np(x, np.float32)
        ei = _np(edge_index, np.int64)
        W = _np(W, np.float32)
        b = None if b is None else _np(b, np.float32)
        n, F = x.shape
        O = W.shape[0]
        out = np.empty((n, O), dtype=np.float32)
        check(self._L.gigl_gcn_conv_host(self.handle, n, ei.shape[1], F, O, _hp(ei), _hp(x), _hp(W), _hp(b), _hp(out),
                                         int(relu)), self.handle)
        return out

    # ---- aggregate, device tensors --------------------------------------------------------
    def edge_rows_host(self, n_nodes: int, src, dst, is_graph_directed: bool) -> np.ndarray:
        """Input edge record that hydrates every slot of the in-CSR `Graph.from_edges_host` builds from the same
        arguments (gigl_edge_rows_host)."""
        src, dst = _np(src, np.int32), _np(dst, np.int32)
        cap = len(src) * (1 if is_graph_directed else 2)
        rows = np.empty(max(cap, 1), dtype=np.int32)
        n = C.c_int64()
        check(self._L.gigl_edge_rows_host(self.handle, n_nodes, len(src), _hp(src), _hp(dst), int(is_graph_directed), _hp(rows), cap,
                                          C.byref(n)), self.handle)
        return rows[: n.value]

    def csr_from_coo(self, n: int, edge_index):
        """edge_index: int64 CUDA tensor [2, e] -> (rowptr int64 [n+1], col int32 [e]), rows stable."""
        import torch

        assert edge_index.dtype == torch.int64 and edge_index.dim() == 2 and edge_index.shape[0] == 2
        ei = edge_index.contiguous()
        e = ei.shape[1]
        rowptr = torch.empty(n + 1, dtype=torch.int64, device=ei.device)
        col = torch.empty(max(e, 1), dtype=torch.int32, device=ei.device)
        check(self._L.gigl_csr_from_coo_dev(self.handle, n, e, _dp(ei[0]), _dp(ei[1]), _dp(rowptr), _dp(col)), self.handle)
        return rowptr, col[:e]

    def sage_conv(self, x, rowptr, col, Wl, bl, Wr, relu: bool = False, n_rows_out: Optional[int] = None, out=None):
        import torch

        n, F = x.shape
        O = Wl.shape[0]
        m = n if n_rows_out is None else n_rows_out
        if out is None:
            out = torch.empty((m, O), dtype=torch.float32, device=x.device)
        check(self._L.gigl_sage_conv_dev(self.handle, n, m, F, O, _dp(rowptr), _dp(col), _dp(x), _dp(Wl), _dp(bl), _dp(Wr),
                                         _dp(out), int(relu)), self.handle)
        return out

    def linear(self, A, W, bias=None, relu: bool = False, out=None):
        """F.linear(A, W, bias) on the tensor cores (3xTF32): A [M, K], W [N, K] fp32 CUDA tensors."""
        import torch

        assert A.is_cuda and W.is_cuda and A.stride(1) == 1 and W.stride(1) == 1 and A.shape[1] == W.shape[1]
        M, K = A.shape
        N = W.shape[0]
        if out is None:
            out = torch.empty((M, N), dtype=torch.float32, device=A.device)
        check(self._L.gigl_linear_dev(self.handle, M, N, K, _dp_any(A), A.stride(0), _dp_any(W), W.stride(0),
                                      None if bias is None else _dp(bias), _dp_any(out), out.stride(0), int(relu)), self.handle)
        return out

    def linear_tn(self, G, A, out=None, accumulate: bool = False):
        """G[R, M]^T @ A[R, N] -> [M, N] (the weight gradient of F.linear) on the tensor cores, deterministic."""
        import torch

        assert G.is_cuda and A.is_cuda and G.shape[0] == A.shape[0]
        assert G.shape[0] == 0 or (G.stride(1) == 1 and A.stride(1) == 1)
        R, M = G.shape
        N = A.shape[1]
        if out is None:
            out = torch.empty((M, N), dtype=torch.float32, device=G.device)
        check(self._L.gigl_linear_tn_dev(self.handle, R, M, N, _dp_any(G), max(G.stride(0), M), _dp_any(A), max(A.stride(0), N), _dp_any(out),
                                         out.stride(0), int(accumulate)), self.handle)
        return out

    def frontier_distinct(self, cur, cur_slots: int, prev: Sequence = ()):
        """cur: int32 CUDA tensor [n_roots * cur_slots] (a parent level of a SamplingOp); prev: [(tensor, slots), ...] the
        levels of the op's earlier input instances.  Returns a copy with, per root, only the first slot of every distinct
        node kept (gigl_frontier_distinct_dev; GraphDBSampler.scala:66-82 expands the SET of the parents' results)."""
        import torch

        n_roots = cur.numel() // cur_slots
        out = torch.empty_like(cur)
        n_prev = len(prev)
        pp = (C.c_void_p * max(n_prev, 1))(*[t.data_ptr() for t, _ in prev])
        ps = _np([s for _, s in prev] or [1], np.int32)
        check(self._L.gigl_frontier_distinct_dev(self.handle, n_roots, n_prev, pp, _hp(ps), _dp(cur), cur_slots, _dp(out)), self.handle)
        return out

    def gather_mean(self, x, rowptr, col, n_rows_out: Optional[int] = None, out=None):
        import torch

        n, F = x.shape
        m = n if n_rows_out is None else n_rows_out
        if out is None:
            out = torch.empty((m, F), dtype=torch.float32, device=x.device)
        check(self._L.gigl_gather_mean_dev(self.handle, m, F, _dp(rowptr), _dp(col), _dp(x), _dp(out)), self.handle)
        return out

    def gcn_conv(self, x, rowptr, col, W, b, relu: bool = False, out=None):
        import torch

        n, F = x.shape
        O = W.shape[0]
        if out is None:
            out = torch.empty((n, O), dtype=torch.float32, device=x.device)
        check(self._L.gigl_gcn_conv_dev(self.handle, n, F, O, _dp(rowptr), _dp(col), _dp(x), _dp(W), _dp(b), _dp(out),
                                        int(relu)), self.handle)
        return out


class Graph:
    """gigl_graph: sorted CSR (rows = in-neighbours, or out-neighbours with ``by_source``) in HBM."""

    def __init__(self, ctx: Context, handle, keepalive=None):
        self.ctx = ctx
        self.handle = handle
        self._keepalive = keepalive
        n, e = C.c_int64(), C.c_int64()
        check(ctx._L.gigl_graph_num_nodes(handle, C.byref(n), C.byref(e)), ctx.handle)
        self.n_nodes, self.n_edges = n.value, e.value

    # -- constructors
    @classmethod
    def from_csr_host(cls, ctx: Context, rowptr, col) -> "Graph":
        rowptr, col = _np(rowptr, np.int64), _np(col, np.int32)
        h = C.c_void_p()
        check(ctx._L.gigl_graph_create_host(ctx.handle, len(rowptr) - 1, len(col), _hp(rowptr), _hp(col), C.byref(h)), ctx.handle)
        return cls(ctx, h)

    @classmethod
    def from_edges_host(cls, ctx: Context, n_nodes: int, src, dst, is_graph_directed: bool, by_source: bool = False) -> "Graph":
        src, dst = _np(src, np.int32), _np(dst, np.int32)
        assert src.shape == dst.shape
        h = C.c_void_p()
        check(ctx._L.gigl_graph_from_edges_host(ctx.handle, n_nodes, len(src), _hp(src), _hp(dst), int(is_graph_directed),
                                                int(by_source), C.byref(h)), ctx.handle)
        return cls(ctx, h)

    @classmethod
    def from_edges_dev(cls, ctx: Context, n_nodes: int, src, dst, is_graph_directed: bool, by_source: bool = False) -> "Graph":
        import torch

        assert src.dtype == torch.int32 and dst.dtype == torch.int32 and src.shape == dst.shape
        h = C.c_void_p()
        check(ctx._L.gigl_graph_from_edges_dev(ctx.handle, n_nodes, src.numel(), _dp(src), _dp(dst), int(is_graph_directed),
                                               int(by_source), C.byref(h)), ctx.handle)
        return cls(ctx, h)

    @classmethod
    def wrap_dev(cls, ctx: Context, rowptr, col) -> "Graph":
        h = C.c_void_p()
        check(ctx._L.gigl_graph_wrap_dev(ctx.handle, rowptr.numel() - 1, col.numel(), _dp(rowptr), _dp(col), C.byref(h)), ctx.handle)
        return cls(ctx, h, keepalive=(rowptr, col))

    def csr_tensors(self):
        """(rowptr int64 [n+1], col int32 [E]) as torch views of the resident arrays (no copy)."""
        import torch

        if self._keepalive is not None:
            return self._keepalive
        pr, pc = C.c_void_p(), C.c_void_p()
        check(self.ctx._L.gigl_graph_device_ptrs(self.handle, C.byref(pr), C.byref(pc)), self.ctx.handle)
        dev = torch.device("cuda", self.ctx.device)
        rowptr = _tensor_from_ptr(pr.value, (self.n_nodes + 1,), torch.int64, dev, owner=self)
        col = _tensor_from_ptr(pc.value, (max(self.n_edges, 1),), torch.int32, dev, owner=self)[: self.n_edges]
        return rowptr, col

    def csr_host(self) -> Tuple[np.ndarray, np.ndarray]:
        rowptr, col = self.csr_tensors()
        return rowptr.cpu().numpy(), col.cpu().numpy()

    # -- sampling
    def set_hash_index(self, enabled: bool) -> None:
        """Switch the sampler's hash-window index on/off (results are identical either way)."""
        check(self.ctx._L.gigl_graph_set_hash_index(self.handle, int(enabled)), self.ctx.handle)

    def sample_khop_host(self, roots, fanouts: Sequence[int], base_seed: int = 42, first_call_no: int = 1
                         ) -> Tuple[List[np.ndarray], List[np.ndarray]]:
        """Padded-tree k-hop sample (layout: include/gigl_b200.h) through host buffers."""
        roots = _np(roots, np.int32)
        fan = _np(fanouts, np.int32)
        n_roots, n_hops = len(roots), len(fan)
        nbr, cnt, width = [], [], 1
        for f in fan:
            cnt.append(np.zeros(n_roots * width, dtype=np.int32))
            width *= int(f)
            nbr.append(np.full(n_roots * width, -1, dtype=np.int32))
        pn = (C.c_void_p * max(n_hops, 1))(*[a.ctypes.data for a in nbr])
        pc = (C.c_void_p * max(n_hops, 1))(*[a.ctypes.data for a in cnt])
        check(self.ctx._L.gigl_sample_khop_host(self.handle, _hp(roots), n_roots, _hp(fan), n_hops, base_seed, first_call_no,
                                                pn, pc), self.ctx.handle)
        return nbr, cnt

    def sample_khop(self, roots, fanouts: Sequence[int], base_seed: int = 42, first_call_no: int = 1, out=None, stage_into=None):
        """Device variant: ``roots`` int32 CUDA tensor -> (nbr, cnt) lists of int32 CUDA tensors.
        Asynchronous; device-side errors surface at ``ctx.sync()``.  ``stage_into``: the :class:`Batch` this sample is
        collated into next - with a halo table registered there (``Batch.set_halo_staging(True, x)``) every sampled level's
        feature rows start their way over NVLink while the next hop is sampled (gigl_sample_khop_staged_dev)."""
        import torch

        assert roots.dtype == torch.int32
        fan = _np(fanouts, np.int32)
        n_roots, n_hops = roots.numel(), len(fan)
        if out is None:
            nbr, cnt, width = [], [], 1
            for f in fan:
                cnt.append(torch.empty(n_roots * width, dtype=torch.int32, device=roots.device))
                width *= int(f)
                nbr.append(torch.empty(n_roots * width, dtype=torch.int32, device=roots.device))
        else:
            nbr, cnt = out
        pn = (C.c_void_p * max(n_hops, 1))(*[t.data_ptr() for t in nbr])
        pc = (C.c_void_p * max(n_hops, 1))(*[t.data_ptr() for t in cnt])
        if stage_into is not None:
            check(self.ctx._L.gigl_sample_khop_staged_dev(self.handle, stage_into.handle, _dp(roots), n_roots, _hp(fan), n_hops, base_seed,
                                                          first_call_no, pn, pc), self.ctx.handle)
        else:
            check(self.ctx._L.gigl_sample_khop_dev(self.handle, _dp(roots), n_roots, _hp(fan), n_hops, base_seed, first_call_no,
                                                   pn, pc), self.ctx.handle)
        return nbr, cnt

    def sample_op(self, roots, chain_fanouts: Sequence[int], chain_nbr: Sequence, call_no: int, base_seed: int = 42, weights=None,
                  method: str = "random_uniform"):
        """One SamplingOp over this graph's edge type (gigl_sample_op_dev): expands the frontier `chain_nbr[-1]` (or the
        roots when the chain is empty).  chain_fanouts = the ancestors' fanouts followed by this op's.  Device tensors in,
        (nbr, cnt) device tensors out.  method "top_k" / "random_weighted" (gigl_sample_op_weighted_dev) takes `weights`:
        a float32 CUDA tensor with the op's edge feature per CSR position of this graph."""
        import torch

        if method not in SAMPLING_METHODS:
            raise ValueError(f"sampling method {method!r}: expected one of {sorted(SAMPLING_METHODS)}")

        fan = _np(chain_fanouts, np.int32)
        depth = len(fan)
        assert len(chain_nbr) == depth - 1, "one ancestor level per ancestor fanout"
        n_roots = int(roots.numel())
        parents = n_roots * int(np.prod(fan[:-1], dtype=np.int64))
        nbr = torch.empty(parents * int(fan[-1]), dtype=torch.int32, device=roots.device)
        cnt = torch.empty(parents, dtype=torch.int32, device=roots.device)
        pn = (C.c_void_p * max(depth - 1, 1))(*[t.data_ptr() for t in chain_nbr])
        if SAMPLING_METHODS[method] != 0:
            if weights is None or weights.dtype != torch.float32 or not weights.is_cuda or weights.numel() != self.n_edges:
                raise ValueError(f"{method} sampling needs a float32 CUDA tensor of {self.n_edges} edge weights (one per CSR position)")
            weights = weights.contiguous()
            check(self.ctx._L.gigl_sample_op_weighted_dev(self.handle, _dp(roots), n_roots, depth, _hp(fan), pn, _dp(weights),
                                                          SAMPLING_METHODS[method], base_seed, call_no, _dp(nbr), _dp(cnt)), self.ctx.handle)
            return nbr, cnt
        check(self.ctx._L.gigl_sample_op_dev(self.handle, _dp(roots), n_roots, depth, _hp(fan), pn, base_seed, call_no, _dp(nbr), _dp(cnt)),
              self.ctx.handle)
        return nbr, cnt

    def sample_op_host(self, roots, chain_fanouts: Sequence[int], chain_nbr: Sequence[np.ndarray], call_no: int, base_seed: int = 42):
        """Host-buffer form of :meth:`sample_op` (gigl_sample_op_host)."""
        roots = _np(roots, np.int32)
        fan = _np(chain_fanouts, np.int32)
        depth = len(fan)
        chain = [_np(a, np.int32) for a in chain_nbr]
        parents = len(roots) * int(np.prod(fan[:-1], dtype=np.int64))
        nbr = np.full(parents * int(fan[-1]), -1, dtype=np.int32)
        cnt = np.zeros(parents, dtype=np.int32)
        pn = (C.c_void_p * max(depth - 1, 1))(*[a.ctypes.data for a in chain])
        check(self.ctx._L.gigl_sample_op_host(self.handle, _hp(roots), len(roots), depth, _hp(fan), pn, base_seed, call_no, _hp(nbr),
                                              _hp(cnt)), self.ctx.handle)
        return nbr, cnt

    def sample_positives_host(self, srcs, num_pos: int, base_seed: int = 42, call_no: int = 3):
        srcs = _np(srcs, np.int32)
        pos = np.full(len(srcs) * num_pos, -1, dtype=np.int32)
        cnt = np.zeros(len(srcs), dtype=np.int32)
        check(self.ctx._L.gigl_sample_positives_host(self.handle, _hp(srcs), len(srcs), num_pos, base_seed, call_no, _hp(pos),
                                                     _hp(cnt)), self.ctx.handle)
        return pos, cnt

    # -- resident features
    def set_features_host(self, x) -> None:
        x = _np(x, np.float32)
        assert x.ndim == 2 and x.shape[0] == self.n_nodes
        check(self.ctx._L.gigl_graph_set_features_host(self.handle, _hp(x), x.shape[1]), self.ctx.handle)
        self._x = None
        self.F = x.shape[1]

    def set_features(self, x) -> None:
        """Wrap a [n_nodes, F] fp32 CUDA tensor (kept alive by this object); rows may be pitched (x.stride(0) >= F)."""
        assert x.is_cuda and x.dim() == 2 and x.shape[0] == self.n_nodes and x.stride(1) == 1 and x.stride(0) >= x.shape[1]
        check(self.ctx._L.gigl_graph_set_features_pitched_dev(self.handle, C.c_void_p(x.data_ptr()), x.shape[1], x.stride(0)), self.ctx.handle)
        self._x = x
        self.F = x.shape[1]

    def infer_khop_sage_host(self, batch: "Batch", model: "SageModel", roots, fanouts: Sequence[int], base_seed: int = 42,
                             first_call_no: int = 1, return_samples: bool = False, out=None, samples_out=None):
        """roots (host) -> sample -> collate -> GraphSAGE -> root embeddings (host numpy [n_roots, O])."""
        roots = _np(roots, np.int32)
        fan = _np(fanouts, np.int32)
        n_roots, n_hops = len(roots), len(fan)
        if out is None:
            out = np.empty((n_roots, model.dims[-1]), dtype=np.float32)
        pn = pc = None
        nbr = cnt = None
        if return_samples:
            if samples_out is None:
                nbr, cnt, width = [], [], 1
                for f in fan:
                    cnt.append(np.empty(n_roots * width, dtype=np.int32))
                    width *= int(f)
                    nbr.append(np.empty(n_roots * width, dtype=np.int32))
            else:
                nbr, cnt = samples_out
            pn = (C.c_void_p * n_hops)(*[a.ctypes.data for a in nbr])
            pc = (C.c_void_p * n_hops)(*[a.ctypes.data for a in cnt])
        check(self.ctx._L.gigl_infer_khop_sage_host(self.handle, batch.handle, model.handle, _hp(roots), n_roots, _hp(fan),
                                                    n_hops, base_seed, first_call_no, _hp(out), pn, pc), self.ctx.handle)
        return (out, nbr, cnt) if return_samples else out

    def infer_khop_sage_packed_host(self, batch: "Batch", model: "SageModel", roots, fanouts: Sequence[int], base_seed: int = 42,
                                    first_call_no: int = 1, out=None, packed_out=None):
        """As :meth:`infer_khop_sage_host` with the index sets PACKED (gigl_infer_khop_sage_packed_host): returns
        (embeddings, packed int32 [n_filled], [cnt_u8 per hop]); :func:`unpack_tree` restores the padded layout.
        packed_out = (packed buffer of capacity >= the tree's slot count, [cnt_u8 buffers]) to reuse (pinned) host memory."""
        roots = _np(roots, np.int32)
        fan = _np(fanouts, np.int32)
        n_roots, n_hops = len(roots), len(fan)
        if out is None:
            out = np.empty((n_roots, model.dims[-1]), dtype=np.float32)
        if packed_out is None:
            cnt_u8, width, slots = [], 1, 0
            for f in fan:
                cnt_u8.append(np.empty(n_roots * width, dtype=np.uint8))
                width *= int(f)
                slots += n_roots * width
            packed = np.empty(max(slots, 1), dtype=np.int32)
        else:
            packed, cnt_u8 = packed_out
        pc = (C.c_void_p * n_hops)(*[a.ctypes.data for a in cnt_u8])
        n_packed = C.c_int64()
        check(self.ctx._L.gigl_infer_khop_sage_packed_host(self.handle, batch.handle, model.handle, _hp(roots), n_roots, _hp(fan), n_hops,
                                                           base_seed, first_call_no, _hp(out), pc, _hp(packed), packed.size,
                                                           C.byref(n_packed)), self.ctx.handle)
        return out, packed[: n_packed.value], cnt_u8

    def infer_khop_sage_bitpacked_host(self, batch: "Batch", model: "SageModel", roots, fanouts: Sequence[int], base_seed: int = 42,
                                       first_call_no: int = 1, out=None, packed_out=None):
        """As :meth:`infer_khop_sage_packed_host` with the ids as a bit stream (gigl_infer_khop_sage_bitpacked_host): returns
        (embeddings, words uint32 [ceil(n * bits / 32)], [cnt_u8 per hop], n, bits); `unpack_bits(words, n, bits)` gives the
        int32 `packed` array of the packed form.  packed_out = (uint32 word buffer, [cnt_u8 buffers]) to reuse host memory."""
        roots = _np(roots, np.int32)
        fan = _np(fanouts, np.int32)
        n_roots, n_hops = len(roots), len(fan)
        if out is None:
            out = np.empty((n_roots, model.dims[-1]), dtype=np.float32)
        if packed_out is None:
            cnt_u8, width, slots = [], 1, 0
            for f in fan:
                cnt_u8.append(np.empty(n_roots * width, dtype=np.uint8))
                width *= int(f)
                slots += n_roots * width
            words = np.empty(max(slots, 1) + 1, dtype=np.uint32)
        else:
            words, cnt_u8 = packed_out
        pc = (C.c_void_p * n_hops)(*[a.ctypes.data for a in cnt_u8])
        n_packed, bits = C.c_int64(), C.c_int32()
        check(self.ctx._L.gigl_infer_khop_sage_bitpacked_host(self.handle, batch.handle, model.handle, _hp(roots), n_roots, _hp(fan), n_hops,
                                                              base_seed, first_call_no, _hp(out), pc, _hp(words), words.size,
                                                              C.byref(n_packed), C.byref(bits)), self.ctx.handle)
        return out, words[: (n_packed.value * bits.value + 31) // 32], cnt_u8, n_packed.value, bits.value

    def close(self) -> None:
        if self.handle and self.ctx.handle:
            self.ctx._L.gigl_graph_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unpack_bits(words, n: int, bits: int) -> np.ndarray:
    """The bit stream of gigl_infer_khop_sage_bitpacked_host -> int32 ids (what gigl_unpack_bits_host does, vectorised)."""
    w = np.concatenate([np.asarray(words, dtype=np.uint32), np.zeros(1, np.uint32)]).astype(np.uint64)
    bit = np.arange(n, dtype=np.int64) * bits
    wi, s = bit >> 5, (bit & 31).astype(np.uint64)
    v = (w[wi] >> s) | (w[wi + 1] << (np.uint64(32) - s))
    return (v & np.uint64((1 << bits) - 1)).astype(np.int32)


def unpack_tree(packed, cnt_u8: Sequence[np.ndarray], fanouts: Sequence[int]):
    """Packed index sets (gigl_infer_khop_sage_packed_host) -> the padded tree of `Graph.sample_khop_host`:
    (nbr per hop with -1 padding, cnt per hop as int32)."""
    packed = np.asarray(packed, dtype=np.int32)
    nbr, cnt, pos = [], [], 0
    for c8, f in zip(cnt_u8, fanouts):
        c = np.asarray(c8).astype(np.int32)
        total = int(c.sum(dtype=np.int64))
        if pos + total > len(packed):
            raise ValueError("packed index sets are shorter than their counts say")
        lvl = np.full((len(c), int(f)), -1, dtype=np.int32)
        lvl[np.arange(int(f), dtype=np.int32)[None, :] < c[:, None]] = packed[pos:pos + total]
        pos += total
        nbr.append(lvl.reshape(-1))
        cnt.append(c)
    if pos != len(packed):
        raise ValueError("packed index sets are longer than their counts say")
    return nbr, cnt


class SageModel:
    """gigl_sage_model: torch_geometric.nn.GraphSAGE weights resident on the device.

    ``layers`` = [(lin_l.weight [O,F], lin_l.bias [O] | None, lin_r.weight [O,F]), ...] as numpy arrays
    (PyG state_dict order, graphsage_template_modeling_spec.py:143-148) or CUDA tensors."""

    def __init__(self, ctx: Context, layers):
        self.ctx = ctx
        n = len(layers)
        on_dev = hasattr(layers[0][0], "is_cuda")
        if on_dev:
            keep = [(Wl.contiguous(), None if bl is None else bl.contiguous(), Wr.contiguous()) for Wl, bl, Wr in layers]
            ptr = lambda t: 0 if t is None else t.data_ptr()
        else:
            keep = [(_np(Wl, np.float32), None if bl is None else _np(bl, np.float32), _np(Wr, np.float32)) for Wl, bl, Wr in layers]
            ptr = lambda a: 0 if a is None else a.ctypes.data
        self.dims = [int(keep[0][0].shape[1])] + [int(k[0].shape[0]) for k in keep]
        for l, (Wl, bl, Wr) in enumerate(keep):
            assert tuple(Wl.shape) == (self.dims[l + 1], self.dims[l]) and tuple(Wr.shape) == tuple(Wl.shape)
        dims = _np(self.dims, np.int32)
        pWl = (C.c_void_p * n)(*[ptr(k[0]) for k in keep])
        pbl = (C.c_void_p * n)(*[ptr(k[1]) for k in keep])
        pWr = (C.c_void_p * n)(*[ptr(k[2]) for k in keep])
        h = C.c_void_p()
        fn = ctx._L.gigl_sage_model_create_dev if on_dev else ctx._L.gigl_sage_model_create_host
        check(fn(ctx.handle, n, _hp(dims), pWl, pbl, pWr, C.byref(h)), ctx.handle)
        self.handle = h
        self.n_layers = n

    def close(self) -> None:
        if self.handle and self.ctx.handle:
            self.ctx._L.gigl_sage_model_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """gigl_batch: reusable collation workspace (B sampled neighbourhoods -> one coalesced graph)."""

    def __init__(self, ctx: Context, n_graph_nodes: int):
        self.ctx = ctx
        h = C.c_void_p()
        check(ctx._L.gigl_batch_create(ctx.handle, n_graph_nodes, C.byref(h)), ctx.handle)
        self.handle = h
        self.level_sizes: List[int] = []
        self.n_edges = 0

    def collate(self, roots, fanouts: Sequence[int], nbr, n_layers: int):
        """roots: int32 CUDA tensor; nbr: the per-hop tensors of Graph.sample_khop.  Returns level sizes."""
        fan = _np(fanouts, np.int32)
        n_hops = len(fan)
        pn = (C.c_void_p * n_hops)(*[t.data_ptr() for t in nbr])
        ls = (C.c_int64 * n_layers)()
        ne = C.c_int64()
        check(self.ctx._L.gigl_batch_collate_dev(self.handle, _dp(roots), roots.numel(), _hp(fan), n_hops, pn, n_layers, ls,
                                                 C.byref(ne)), self.ctx.handle)
        self.level_sizes = [int(v) for v in ls]
        self.n_edges = int(ne.value)
        self._n_roots = roots.numel()
        self._device = roots.device
        return self.level_sizes

    def set_halo_staging(self, enabled: bool, x=None) -> None:
        """Sharded feature table: copy every unique batch node's row into local HBM once (one row per node over NVLink)
        and let layer 1 gather from that copy, instead of one peer load per unique edge (gigl_batch_set_halo_staging).
        With the feature table `x` (the tensor later passed to sage_forward) the copy is forked onto a side stream inside
        collate() and runs under the collation (gigl_batch_set_halo_table_dev)."""
        check(self.ctx._L.gigl_batch_set_halo_staging(self.handle, int(bool(enabled))), self.ctx.handle)
        if enabled and x is not None:
            assert x.is_cuda and x.dim() == 2 and x.stride(1) == 1
            check(self.ctx._L.gigl_batch_set_halo_table_dev(self.handle, _dp_any(x), x.shape[1], x.stride(0)), self.ctx.handle)
            self._halo_x = x
        else:
            check(self.ctx._L.gigl_batch_set_halo_table_dev(self.handle, None, 0, 0), self.ctx.handle)
            self._halo_x = None

    def set_hot_rows(self, graph: "Graph", x, fraction: float) -> int:
        """Staged halo of a sharded feature table: replicate the rows of the `fraction` of the vertices with the highest
        degree on this GPU (gigl_batch_set_hot_rows_dev), so only a batch's cold tail crosses NVLink.  x = the flat
        [n_nodes, F] table (peer rows are read once, here).  Returns the number of replicated rows."""
        from .sharding import hot_rows

        slot, table = hot_rows(graph, x, fraction)
        n_hot = int(table.shape[0])
        if n_hot == 0:
            check(self.ctx._L.gigl_batch_set_hot_rows_dev(self.handle, None, None, 0, 0), self.ctx.handle)
            self._hot = None
            return 0
        check(self.ctx._L.gigl_batch_set_hot_rows_dev(self.handle, _dp(slot), _dp_any(table), x.shape[1], table.stride(0)), self.ctx.handle)
        self._hot = (slot, table)  # owned here: the library keeps the pointers
        return n_hot

    def share_hot_rows(self, other: "Batch", F: int) -> None:
        """Use the hot-row replica another batch workspace of the same GPU already built (one copy per GPU, not per stream)."""
        slot, table = other._hot
        check(self.ctx._L.gigl_batch_set_hot_rows_dev(self.handle, _dp(slot), _dp_any(table), F, table.stride(0)), self.ctx.handle)
        self._hot = other._hot

    def sage_forward(self, model: SageModel, x, out=None):
        import torch

        assert x.is_cuda and x.dtype == torch.float32 and x.stride(1) == 1
        if out is None:
            out = torch.empty((self._n_roots, model.dims[-1]), dtype=torch.float32, device=x.device)
        check(self.ctx._L.gigl_batch_sage_forward_dev(self.handle, model.handle, _dp_any(x), x.stride(0), _dp(out)), self.ctx.handle)
        return out

    def export(self):
        """(node_ids int32 [n], edge_index int64 [2, e] in local ids) as CUDA tensors."""
        import torch

        n, e = C.c_int64(), C.c_int64()
        check(self.ctx._L.gigl_batch_finalize_nodes(self.handle, C.byref(n), C.byref(e)), self.ctx.handle)
        node_ids = torch.empty(n.value, dtype=torch.int32, device=self._device)
        ei = torch.empty((2, e.value), dtype=torch.int64, device=self._device)
        check(self.ctx._L.gigl_batch_export_dev(self.handle, _dp(node_ids) if n.value else None,
                                                _dp(ei) if e.value else None), self.ctx.handle)
        return node_ids, ei

    def close(self) -> None:
        if self.handle and self.ctx.handle:
            self.ctx._L.gigl_batch_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _dp_any(t):
    assert t.is_cuda
    return C.c_void_p(t.data_ptr())


def _tensor_from_ptr(ptr: int, shape, dtype, device, owner=None):
    """Zero-copy torch view of library-owned device memory via __cuda_array_interface__."""
    import torch

    typestr = {torch.int64: "<i8", torch.int32: "<i4", torch.float32: "<f4"}[dtype]

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    h._owner = owner
    with torch.cuda.device(device):
        t = torch.as_tensor(h, device=device)
    t._gigl_owner = owner
    return t
