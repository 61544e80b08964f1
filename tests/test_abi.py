"""CPU-only: the C-ABI library loads and exports every symbol include/dgll_b200.h declares; no compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dgll_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    names = re.findall(r"\b(?:int|void|int64_t|const\s+char\s*\*)\s+\*?\s*((?:dgllb_|launch_)\w+)\s*\(", src)
    return sorted(set(names))


@pytest.fixture(scope="module")
def built_lib():
    from dgll_b200 import build as b
    return b.build()


def test_header_declares_expected_families():
    names = declared_symbols()
    for must in ("dgllb_spmm_csr", "dgllb_gather_rows", "dgllb_gat_forward", "dgllb_bin_spmm_csr",
                 "dgllb_gemm_f32", "launch_gcn_fused_kernel", "launch_gcn_fused_kernel_backward_optimized"):
        assert must in names
    assert len(names) >= 20


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, "declared in include/dgll_b200.h but not exported: %s" % missing


def test_ctypes_signatures_cover_header(built_lib):
    from dgll_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.lib()  # resolves all of them with argtypes
    assert _lib.lib().dgllb_version() >= 1000


def test_no_cpu_fallback_in_product():
    """Nothing under dgll_b200/ may import the oracle (the product has no CPU path)."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "dgll_b200")):
        for f in fs:
            if f.endswith(".py"):
                s = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_cpu_tensor_raises():
    import torch
    from dgll_b200 import kernels as K
    rp = torch.tensor([0, 1], dtype=torch.int64)
    col = torch.tensor([0], dtype=torch.int32)
    x = torch.ones(1, 4)
    with pytest.raises(RuntimeError):
        K.spmm_csr(rp, col, x)


def test_option_table_round_trip(built_lib):
    """dgllb_set_option / dgllb_get_option need no device: every documented option is known, word values map to their
    codes, numbers round-trip, ``None`` restores the default, unknown names and bad values are rejected with a message."""
    from dgll_b200 import _lib
    words = {"spmm_kernel": {"auto": 0, "rowsplit": 1, "stream": 2, "wholerow": 3},
             "gat_kernel": {"auto": 0, "group": 1, "row": 2},
             "gat_bwd_kernel": {"auto": 0, "twopass": 1, "fused": 2}}
    numeric = ["spmm_tb", "rows_tb", "rows_ns", "rows_d", "rows_stream", "rows_sharded_bps", "gat_row_warps", "gat_bwd_tb",
               "gat_bwd_depth", "bin_tb", "gemm_kernel", "nvtx"]
    try:
        for name, table in words.items():
            for w, code in table.items():
                _lib.set_option(name, w)
                assert _lib.get_option(name) == code
            _lib.set_option(name, None)
            assert _lib.get_option(name) == 0
        for name in numeric:
            _lib.set_option(name, 7)
            assert _lib.get_option(name) == 7
            _lib.set_option(name, None)
            assert _lib.get_option(name) == 0
        with pytest.raises(RuntimeError):
            _lib.set_option("no_such_option", "1")
        with pytest.raises(RuntimeError):
            _lib.set_option("spmm_kernel", "bogus")
    finally:
        for name in list(words) + numeric:
            _lib.set_option(name, None)
