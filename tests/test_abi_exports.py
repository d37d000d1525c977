"""CPU-side checks: the C-ABI library loads and exports every symbol include/faqcs_b200.h declares;
no compute calls are made (there is no GPU in the CPU container)."""
import ctypes
import os
import re

import pytest

from faqcs_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "faqcs_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fq_[a-z_0-9]+)\s*\(", txt)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("fq_create", "fq_destroy", "fq_autodetect", "fq_process_host", "fq_process_device", "fq_stats",
                 "fq_last_error", "fq_stats_device_buffer"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    path = api.default_library_path()
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/faqcs_b200.h but not exported: {missing}"
    lib.fq_abi_version.restype = ctypes.c_int
    assert lib.fq_abi_version() == 4
    lib.fq_build_info.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.fq_build_info()


def test_struct_sizes_match_the_header():
    # sizes the C compiler gives the PODs (x86-64 SysV): guards the ctypes mirror against drift
    assert ctypes.sizeof(api.CReadResult) == 16
    assert ctypes.sizeof(api.COptions) == 80
    assert ctypes.sizeof(api.CBatchOut) == 216
    assert ctypes.sizeof(api.CStatsView) == 336
    assert api.READ_RESULT_DTYPE.itemsize == 16


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    path = api.default_library_path()
    if not os.path.exists(path):
        pytest.skip("library not built")
    with pytest.raises(api.FaqcsError) as e:
        api.Engine(api.Options())
    assert "no CUDA device" in str(e.value) or "FQ_ERR" in str(e.value)


def test_product_never_touches_the_oracle():
    """The shipped package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "faqcs_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "faqcs_oracle" not in txt and "fqo_" not in txt and "oracle/" not in txt, f
