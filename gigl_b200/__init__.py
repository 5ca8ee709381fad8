"""gigl_b200 - B200 (sm_100a) k-hop sampling and GNN aggregate behind GiGL's component API.

The product path is the CUDA library ``gigl_b200/lib/libgigl_b200.so`` (C-ABI: include/gigl_b200.h);
importing this package does not load it, the first use does, and raises if it is missing.
"""
__version__ = "0.2.0"

from ._capi import GiglError  # noqa: F401
from .engine import Batch, Context, Graph, SageModel, unpack_bits, unpack_tree  # noqa: F401
