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


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C11 with no C++ / torch types, and a C program linked against the
    library calls it (host-only entry points here: the TFRecord framing checksum and the record splitter)."""
    import shutil
    import subprocess

    from gigl_b200 import _capi

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "gigl_b200.h")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    _capi.lib()
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "gigl_b200.h"
int main(void) {
    /* one TFRecord: u64 length, masked crc32c(length), payload, masked crc32c(payload) */
    unsigned char rec[8 + 4 + 5 + 4];
    const unsigned long long n = 5;
    memcpy(rec, &n, 8);
    unsigned int c = gigl_crc32c_masked(rec, 8);
    memcpy(rec + 8, &c, 4);
    memcpy(rec + 12, "hello", 5);
    c = gigl_crc32c_masked(rec + 12, 5);
    memcpy(rec + 17, &c, 4);
    int64_t off = -1, len = -1;
    const int64_t k = gigl_tfrecord_index_host(rec, (int64_t)sizeof rec, 1, &off, &len, 1);
    rec[13] ^= 1; /* corrupt the payload: the checksum must catch it */
    const int64_t bad = gigl_tfrecord_index_host(rec, (int64_t)sizeof rec, 1, &off, &len, 1);
    printf("%s|%lld|%lld|%lld|%lld|%u\n", gigl_version(), (long long)k, (long long)off, (long long)len, (long long)bad,
           gigl_crc32c_masked("123456789", 9));
    return 0;
}
''')
    exe = tmp_path / "t"
    libdir = os.path.dirname(_capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                           "-lgigl_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.check_output([str(exe)]).decode().strip().split("|")
    assert "sm_100a" in out[0] and out[1:4] == ["1", "12", "5"] and int(out[4]) < 0
    # crc32c("123456789") = 0xE3069283 (the Castagnoli check value), masked as TFRecord does
    crc = 0xE3069283
    assert int(out[5]) == ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
