"""Drop-in ``torch.nn`` modules for the message-passing layers the reference trains and serves, backed by the
library's CUDA kernels through the C-ABI (forward AND backward; autograd-visible ``nn.Parameter``s).

Reference interfaces mirrored (state_dict keys are PyG 2.5.3's, so checkpoints interchange):

* ``SAGEConv`` / ``GraphSAGE``  <- ``torch_geometric.nn.SAGEConv`` / ``torch_geometric.nn.GraphSAGE`` as built by
  ``python/gigl/src/common/models/pyg/homogeneous.py:171-202`` and
  ``python/gigl/src/common/modeling_task_specs/graphsage_template_modeling_spec.py:143-148``
  (keys ``convs.{l}.lin_l.weight``, ``convs.{l}.lin_l.bias``, ``convs.{l}.lin_r.weight``);
* ``GCNConv`` / ``TwoLayerGCN`` <- ``python/gigl/src/common/models/pyg/homogeneous.py:488-546``
  (keys ``conv{1,2}.lin.weight``, ``conv{1,2}.bias``).

torch is used for parameters, autograd bookkeeping, memory and streams; every FLOP of the layers runs in
``libgigl_b200.so`` (gather kernels + tcgen05 GEMMs).  The layers are dispatcher ops - ``torch.ops.gigl_b200.sage_conv`` /
``gcn_conv`` / ``csr_from_coo``, registered by the C++ extension ``csrc/torch_ops.cpp`` (built in-tree as
``lib/libgigl_b200_torch.so`` over the same C-ABI) with CUDA, Meta and Autograd implementations - so DDP, which the
reference wraps the model in (``training_process.py:298-303``), and graph capture see ordinary ops.  There is no CPU
path: CPU tensors raise in the dispatcher, and a missing extension fails the import.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
from torch import nn

from .engine import Context

_CTX = {}


def _load_torch_ops():
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libgigl_b200_torch.so")
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python __graft_entry__.py` (make -C gigl_b200/csrc torch); "
                          "gigl_b200.nn has no fallback implementation")
    torch.ops.load_library(path)
    return torch.ops.gigl_b200


_OPS = _load_torch_ops()


def context_for(device: torch.device) -> Context:
    """One library context per (device, current torch stream): kernels enqueue where torch's own work is ordered."""
    if device.type != "cuda":
        raise RuntimeError("gigl_b200.nn layers run on CUDA tensors only (there is no CPU fallback)")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    ctx = _CTX.get(key)
    if ctx is None:
        ctx = _CTX[key] = Context(idx, stream=key[1])
    return ctx


class GraphIndex:
    """``edge_index`` [2, e] (row 0 = src j, row 1 = dst i; PyG convention, ``pyg_graph_builder.py:20-69``) held as the
    CSR by destination the forward gathers over, plus - lazily - the CSR by source the backward gathers over."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int):
        assert edge_index.dtype == torch.int64 and edge_index.dim() == 2 and edge_index.shape[0] == 2
        self.ctx = context_for(edge_index.device)
        self.n = int(num_nodes)
        self.e = int(edge_index.shape[1])
        self._ei = edge_index.contiguous()
        self.rowptr, self.col = self._build(self._ei[0], self._ei[1])
        self._t = None

    def _build(self, src, dst):
        if src.device.type != "cuda":
            raise RuntimeError("gigl_b200.nn layers run on CUDA tensors only (there is no CPU fallback)")
        return _OPS.csr_from_coo(src, dst, self.n)

    @property
    def transposed(self):
        if self._t is None:
            self._t = self._build(self._ei[1], self._ei[0])
        return self._t


def _as_index(edge_index, n: int) -> GraphIndex:
    if isinstance(edge_index, GraphIndex):
        assert edge_index.n == n, "GraphIndex was built for a different node count"
        return edge_index
    return GraphIndex(edge_index, n)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError("gigl_b200.nn computes in fp32")
    return t.contiguous()


class Linear(nn.Module):
    """Parameter holder with ``torch_geometric.nn.dense.linear.Linear``'s names and default initialisation."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, weight_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer = weight_initializer
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight_initializer == "glorot":
            a = math.sqrt(6.0 / (self.in_channels + self.out_channels))
            nn.init.uniform_(self.weight, -a, a)
        else:
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)


class SAGEConv(nn.Module):
    """``SAGEConv(aggr='mean', root_weight=True, bias=True)``: out_i = lin_l(mean_{j->i} x_j) + lin_r(x_i)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = Linear(in_channels, out_channels, bias=bias)
        self.lin_r = Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index, num_rows_out: Optional[int] = None, relu: bool = False):
        """``edge_index``: int64 [2, e] or a :class:`GraphIndex`.  ``num_rows_out`` computes only the first rows
        (what the next layer / the root read-out needs); ``relu`` fuses the inter-layer activation."""
        gi = _as_index(edge_index, x.shape[0])
        m = x.shape[0] if num_rows_out is None else int(num_rows_out)
        # the CSR by source is what the gradient w.r.t. x gathers over: built (once per GraphIndex) only when needed
        t_rowptr, t_col = gi.transposed if (x.requires_grad and torch.is_grad_enabled()) else (None, None)
        return _OPS.sage_conv(_f32c(x), gi.rowptr, gi.col, t_rowptr, t_col, self.lin_l.weight, self.lin_l.bias, self.lin_r.weight, relu, m)


class GraphSAGE(nn.Module):
    """``torch_geometric.nn.GraphSAGE(in, hidden, num_layers, out)`` (BasicGNN: ReLU + dropout between layers, nothing
    after the last one)."""

    def __init__(self, in_channels: int, hidden_channels: int, num_layers: int, out_channels: Optional[int] = None,
                 dropout: float = 0.0):
        super().__init__()
        out_channels = hidden_channels if out_channels is None else out_channels
        dims = [in_channels] + [hidden_channels] * (num_layers - 1) + [out_channels]
        self.in_channels, self.hidden_channels, self.out_channels, self.num_layers = in_channels, hidden_channels, out_channels, num_layers
        self.dropout = dropout
        self.convs = nn.ModuleList([SAGEConv(dims[l], dims[l + 1]) for l in range(num_layers)])

    def forward(self, x, edge_index, level_sizes: Optional[Sequence[int]] = None):
        """``level_sizes`` (optional, from :meth:`gigl_b200.Batch.collate`): ``level_sizes[j]`` = how many leading rows
        the layer ``num_layers - j`` has to produce; with it only the rows the roots depend on are computed and the
        output has ``level_sizes[0]`` (= root) rows.  Without it every layer runs on all nodes, as the reference does."""
        gi = _as_index(edge_index, x.shape[0])
        L = self.num_layers
        for l, conv in enumerate(self.convs):
            m = None if level_sizes is None else int(level_sizes[L - 1 - l])
            x = conv(x, _RowView(gi, x.shape[0]), num_rows_out=m, relu=(l < L - 1))
            if l < L - 1 and self.dropout > 0:
                x = nn.functional.dropout(x, p=self.dropout, training=self.training)
        return x

    @property
    def graph_backend(self) -> str:
        return "PYG"


class _RowView(GraphIndex):
    """The same CSR arrays seen by a layer whose input has only the first ``n`` rows (pruned upper layers)."""

    def __init__(self, base: GraphIndex, n: int):  # noqa: super().__init__ deliberately not called
        self.ctx, self.n, self.e, self._base = base.ctx, n, base.e, base
        self.rowptr, self.col = base.rowptr, base.col

    @property
    def transposed(self):
        return self._base.transposed


class GCNConv(nn.Module):
    """``GCNConv(add_self_loops=True, normalize=True, bias=True)``."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x, edge_index, relu: bool = False):
        gi = _as_index(edge_index, x.shape[0])
        t_rowptr, t_col = gi.transposed
        return _OPS.gcn_conv(_f32c(x), gi.rowptr, gi.col, t_rowptr, t_col, self.lin.weight, self.bias, relu)


class TwoLayerGCN(nn.Module):
    """``homogeneous.py:488-546``: GCNConv -> relu -> dropout(training=is_training) -> GCNConv [-> L2 normalise]."""

    def __init__(self, in_dim: int, out_dim: int, hid_dim: int = 16, is_training: bool = True,
                 should_l2_normalize_output: bool = False, **kwargs):
        super().__init__()
        self.is_training = is_training
        self.should_normalize = should_l2_normalize_output
        self.conv1 = GCNConv(in_dim, hid_dim, bias=kwargs.get("bias", True))
        self.conv2 = GCNConv(hid_dim, out_dim, bias=kwargs.get("bias", True))

    def forward(self, data, edge_index=None):
        x, ei = (data.x, data.edge_index) if edge_index is None else (data, edge_index)
        gi = _as_index(ei, x.shape[0])
        x = self.conv1(x, gi, relu=True)
        x = nn.functional.dropout(x, training=self.is_training)
        x = self.conv2(x, gi)
        if self.should_normalize:
            x = nn.functional.normalize(x, p=2, dim=1)
        return x

    @property
    def graph_backend(self) -> str:
        return "PYG"


def load_reference_state_dict(module: nn.Module, state_dict) -> List[str]:
    """Loads a PyG checkpoint (``convs.0.lin_l.weight`` ...) - the key names are identical, so this is
    ``load_state_dict``; returns the keys for the caller's log."""
    module.load_state_dict(state_dict)
    return list(state_dict.keys())
