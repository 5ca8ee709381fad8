"""ctypes binding of ``libgigl_b200.so`` (C-ABI declared in ``include/gigl_b200.h``).

This is the only place the shared library is loaded.  There is no CPU fallback: if the library
is missing, or no sm_100 device is usable, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GIGL_B200_LIB") or os.path.join(_HERE, "lib", "libgigl_b200.so")  # env override: kernel-variant experiments

OK, E_INVALID, E_CUDA, E_RANGE, E_OVERFLOW, E_NOMEM = 0, -1, -2, -3, -4, -5
MAX_HOPS, MAX_FANOUT = 8, 128
_ERR_NAMES = {E_INVALID: "GIGL_E_INVALID", E_CUDA: "GIGL_E_CUDA", E_RANGE: "GIGL_E_RANGE",
              E_OVERFLOW: "GIGL_E_OVERFLOW", E_NOMEM: "GIGL_E_NOMEM"}


class GiglError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


i32, i64, vp, cp = C.c_int32, C.c_int64, C.c_void_p, C.c_char_p
pvp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); kept in the order of include/gigl_b200.h
SIGNATURES = {
    "gigl_version": (cp, []),
    "gigl_ctx_create": (C.c_int, [C.c_int, pvp]),
    "gigl_ctx_create_on_stream": (C.c_int, [C.c_int, vp, pvp]),
    "gigl_ctx_destroy": (None, [vp]),
    "gigl_ctx_sync": (C.c_int, [vp]),
    "gigl_last_error": (cp, [vp]),
    "gigl_ctx_launch_count": (i64, [vp]),
    "gigl_ctx_stream": (vp, [vp]),
    "gigl_ctx_set_timing": (C.c_int, [vp, i32]),
    "gigl_ctx_reset_timing": (C.c_int, [vp]),
    "gigl_ctx_get_timing": (C.c_int, [vp, i32, C.POINTER(C.c_double), C.POINTER(i64)]),
    "gigl_timing_num_tags": (i32, []),
    "gigl_timing_tag_name": (cp, [i32]),
    "gigl_graph_create_host": (C.c_int, [vp, i64, i64, vp, vp, pvp]),
    "gigl_graph_wrap_dev": (C.c_int, [vp, i64, i64, vp, vp, pvp]),
    "gigl_graph_from_edges_host": (C.c_int, [vp, i64, i64, vp, vp, i32, i32, pvp]),
    "gigl_graph_from_edges_dev": (C.c_int, [vp, i64, i64, vp, vp, i32, i32, pvp]),
    "gigl_edge_rows_host": (C.c_int, [vp, i64, i64, vp, vp, i32, vp, i64, C.POINTER(i64)]),
    "gigl_graph_num_nodes": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]),
    "gigl_graph_device_ptrs": (C.c_int, [vp, pvp, pvp]),
    "gigl_graph_destroy": (None, [vp]),
    "gigl_graph_set_hash_index": (C.c_int, [vp, i32]),
    "gigl_sample_khop_host": (C.c_int, [vp, vp, i64, vp, i32, i32, i32, pvp, pvp]),
    "gigl_sample_khop_dev": (C.c_int, [vp, vp, i64, vp, i32, i32, i32, pvp, pvp]),
    "gigl_sample_khop_staged_dev": (C.c_int, [vp, vp, vp, i64, vp, i32, i32, i32, pvp, pvp]),
    "gigl_sample_op_dev": (C.c_int, [vp, vp, i64, i32, vp, pvp, i32, i32, vp, vp]),
    "gigl_sample_op_weighted_dev": (C.c_int, [vp, vp, i64, i32, vp, pvp, vp, i32, i32, i32, vp, vp]),
    "gigl_sample_op_host": (C.c_int, [vp, vp, i64, i32, vp, pvp, i32, i32, vp, vp]),
    "gigl_sample_positives_host": (C.c_int, [vp, vp, i64, i32, i32, i32, vp, vp]),
    "gigl_validate_samples_host": (C.c_int, [vp, i64, i32, i32, vp, vp, vp, vp, vp]),
    "gigl_frontier_distinct_dev": (C.c_int, [vp, i64, i32, pvp, vp, vp, i32, vp]),
    "gigl_csr_from_coo_dev": (C.c_int, [vp, i64, i64, vp, vp, vp, vp]),
    "gigl_sage_conv_dev": (C.c_int, [vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32]),
    "gigl_sage_conv_host": (C.c_int, [vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp, i32]),
    "gigl_gather_mean_dev": (C.c_int, [vp, i64, i32, vp, vp, vp, vp]),
    "gigl_gcn_conv_dev": (C.c_int, [vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, i32]),
    "gigl_gcn_conv_host": (C.c_int, [vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, i32]),
    "gigl_linear_dev": (C.c_int, [vp, i64, i32, i32, vp, i64, vp, i64, vp, vp, i64, i32]),
    "gigl_sage_conv_train_fwd_dev": (C.c_int, [vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32]),
    "gigl_sage_conv_bwd_dev": (C.c_int, [vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32]),
    "gigl_gcn_conv_bwd_dev": (C.c_int, [vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32]),
    "gigl_linear_tn_dev": (C.c_int, [vp, i64, i32, i32, vp, i64, vp, i64, vp, i64, i32]),
    "gigl_graph_set_features_host": (C.c_int, [vp, vp, i32]),
    "gigl_graph_set_features_dev": (C.c_int, [vp, vp, i32]),
    "gigl_graph_set_features_pitched_dev": (C.c_int, [vp, vp, i32, i64]),
    "gigl_graph_features_dev": (C.c_int, [vp, pvp, C.POINTER(i32)]),
    "gigl_shared_table_row_granule": (C.c_int, [vp, i32, C.POINTER(i64)]),
    "gigl_shared_table_create": (C.c_int, [vp, i32, i32, i64, i32, pvp, C.POINTER(i32)]),
    "gigl_shared_table_attach": (C.c_int, [vp, i32, i32]),
    "gigl_shared_table_ptrs": (C.c_int, [vp, pvp, pvp, C.POINTER(i64)]),
    "gigl_shared_table_destroy": (None, [vp]),
    "gigl_sage_model_create_host": (C.c_int, [vp, i32, vp, pvp, pvp, pvp, pvp]),
    "gigl_sage_model_create_dev": (C.c_int, [vp, i32, vp, pvp, pvp, pvp, pvp]),
    "gigl_sage_model_destroy": (None, [vp]),
    "gigl_batch_create": (C.c_int, [vp, i64, pvp]),
    "gigl_batch_destroy": (None, [vp]),
    "gigl_batch_collate_dev": (C.c_int, [vp, vp, i64, vp, i32, pvp, i32, C.POINTER(i64), C.POINTER(i64)]),
    "gigl_batch_finalize_nodes": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]),
    "gigl_batch_export_dev": (C.c_int, [vp, vp, vp]),
    "gigl_batch_sage_forward_dev": (C.c_int, [vp, vp, vp, i64, vp]),
    "gigl_batch_set_halo_staging": (C.c_int, [vp, i32]),
    "gigl_batch_set_hot_rows_dev": (C.c_int, [vp, vp, vp, i32, i64]),
    "gigl_batch_set_halo_table_dev": (C.c_int, [vp, vp, i32, i64]),
    "gigl_crc32c_masked": (C.c_uint32, [vp, i64]),
    "gigl_free_host": (None, [vp]),
    "gigl_encode_samples_host": (C.c_int, [i32, i64, vp, vp, i32, pvp, vp, i32, i32, i32, vp, cp, i32, pvp, C.POINTER(i64), vp]),
    "gigl_encode_samples_ex_host": (C.c_int, [i32, i64, i64, vp, vp, i32, pvp, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp, cp, i32, vp, vp,
                                              i32, pvp, C.POINTER(i64), vp]),
    "gigl_encode_link_samples_host": (C.c_int, [i64, i64, vp, vp, i32, pvp, vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, i32, vp, vp, i32, pvp,
                                                C.POINTER(i64), vp]),
    "gigl_encode_dag_samples_host": (C.c_int, [i64, vp, i32, i32, vp, i32, vp, i32, pvp, C.POINTER(i64), vp]),
    "gigl_encode_typed_samples_host": (C.c_int, [i32, vp, vp, i32, vp, vp, i32, i32, i32, i32, vp, i32, vp, i32, pvp, C.POINTER(i64), vp]),
    "gigl_tfrecord_index_host": (i64, [vp, i64, i32, vp, vp, i64]),
    "gigl_examples_column_host": (C.c_int, [vp, i64, vp, vp, cp, i32, i32, vp, vp]),
    "gigl_infer_khop_sage_host": (C.c_int, [vp, vp, vp, vp, i64, vp, i32, i32, i32, vp, pvp, pvp]),
    "gigl_infer_khop_sage_packed_host": (C.c_int, [vp, vp, vp, vp, i64, vp, i32, i32, i32, vp, pvp, vp, i64, C.POINTER(i64)]),
    "gigl_infer_khop_sage_bitpacked_host": (C.c_int, [vp, vp, vp, vp, i64, vp, i32, i32, i32, vp, pvp, vp, i64, C.POINTER(i64), C.POINTER(i32)]),
    "gigl_unpack_bits_host": (C.c_int, [vp, i64, i32, vp]),
}



class EdgeTable(C.Structure):
    """gigl_edge_table (include/gigl_b200.h)."""
    _fields_ = [("rowptr", vp), ("col", vp), ("edge_rows", vp), ("feat", vp), ("n_feat", i32)]



class DagOp(C.Structure):
    """gigl_dag_op (include/gigl_b200.h)."""
    _fields_ = [("parent", i32), ("fanout", i32), ("condensed_edge_type", i32), ("result_node_type", i32), ("outgoing", i32), ("nbr", vp)]


class NodeTable(C.Structure):
    """gigl_node_table (include/gigl_b200.h)."""
    _fields_ = [("x", vp), ("n_feat", i32)]


class DagTree(C.Structure):
    """gigl_dag_tree (include/gigl_b200.h)."""
    _fields_ = [("n_roots", i64), ("roots", vp), ("root_node_type", i32), ("n_ops", i32), ("ops", vp)]


_lib = None


def lib() -> C.CDLL:
    """Loads the library once; raises ImportError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C gigl_b200/csrc` (or "
                "`python -c 'import __graft_entry__ as g; g.build()'`). gigl_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, ctx_handle=None) -> None:
    if rc != OK:
        msg = lib().gigl_last_error(ctx_handle)
        raise GiglError(rc, msg.decode() if msg else "")
