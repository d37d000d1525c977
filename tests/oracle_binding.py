"""Test-only binding of the CPU oracle (oracle/_ref/libfaqcs_oracle.so).

The oracle exports the same C shapes as the product under the ``fqo_`` prefix,
so ``faqcs_b200.api.Engine`` drives it unchanged.  Product code never imports
this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from faqcs_b200.api import Engine, Options

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_ref", "libfaqcs_oracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "FaQCs")

_lib = None


def build_oracle():
    src = os.path.join(ORACLE_DIR, "faqcs_oracle.cpp")
    if (not os.path.exists(ORACLE_LIB)) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
    return ORACLE_LIB


def oracle_lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.fqo_quality_trim.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_uint32,
                                          C.POINTER(C.c_uint32)]
        _lib.fqo_quality_trim.restype = C.c_uint32
        _lib.fqo_align.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        _lib.fqo_find_mask_range.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        _lib.fqo_match_threshold.argtypes = [C.c_float, C.c_uint64]
        _lib.fqo_match_threshold.restype = C.c_int32
        _lib.fqo_composition_bin.argtypes = [C.c_uint32, C.c_uint32]
        _lib.fqo_composition_bin.restype = C.c_uint32
        _lib.fqo_average_quality.argtypes = [C.c_char_p, C.c_uint32, C.c_int]
        _lib.fqo_average_quality.restype = C.c_float
    return _lib


class OracleEngine(Engine):
    def __init__(self, options: Options):
        super().__init__(options, lib=oracle_lib(), prefix="fqo_")


def quality_trim(mode, quality, in_offset, protect_5, qual: bytes):
    f5 = C.c_uint32()
    n = oracle_lib().fqo_quality_trim(mode, quality, in_offset, int(protect_5), qual, len(qual), C.byref(f5))
    return f5.value, n


def align(read: bytes, target: bytes, stale=(0, 0)):
    sc, st, sp = C.c_int32(), C.c_int32(stale[0]), C.c_int32(stale[1])
    rc = oracle_lib().fqo_align(read, len(read), target, len(target), C.byref(sc), C.byref(st), C.byref(sp))
    return rc, sc.value, st.value, sp.value


def find_mask_range(mask) -> tuple:
    m = bytes(bytearray(int(bool(x)) for x in mask))
    s, l = C.c_uint32(), C.c_uint32()
    oracle_lib().fqo_find_mask_range(m, len(m), C.byref(s), C.byref(l))
    return s.value, l.value
