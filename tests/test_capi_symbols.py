"""CPU: the C-ABI library loads and exports every symbol include/gigl_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gigl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gigl_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = _declared()
    for must in ("gigl_sample_khop_host", "gigl_sample_khop_dev", "gigl_sage_conv_dev", "gigl_sage_conv_host",
                 "gigl_gcn_conv_dev", "gigl_graph_from_edges_host", "gigl_csr_from_coo_dev"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from gigl_b200 import _capi

    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes table covers the header exactly
    assert sorted(_capi.SIGNATURES) == _declared()
    assert b"sm_100a" in _capi.lib().gigl_version()


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly (GIGL_E_CUDA), not compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gigl_b200 import Context, GiglError

    with pytest.raises(GiglError) as ei:
        Context(0)
    assert ei.value.code == -2


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gigl_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "libgigl_oracle" not in src and "gigl_oracle.c" not in src, fn
