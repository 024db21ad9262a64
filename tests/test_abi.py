"""CPU: the C-ABI shared library loads and exports every symbol include/crnerf_b200.h
declares, the ctypes binding covers exactly that set, and the host-only entry points work.
No compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "crnerf_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crnerf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from crnerf_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_ctypes_binding_matches_header():
    from crnerf_b200 import _lib
    assert sorted(_lib.SIGNATURES.keys()) == declared_symbols()
    lib = _lib.load()
    assert lib.crnerf_abi_version() == 1


def test_host_only_entry_points():
    from crnerf_b200 import _lib
    lib = _lib.load()
    # weights: 77 chunks of 16 KB + 2 of 8 KB; fp32 side blob: sigma head (264) + the biases
    # of the 11 tensor-core layers (9*256 + 128 + 64)
    # ... + the kernel's program tables (104 chunks x 28 B, 20 units x 16 B, 4 ints)
    image, blob, tables = 77 * 16384 + 2 * 8192, (264 + 2496) * 4, 104 * 28 + 20 * 16 + 16
    assert lib.crnerf_mlp_packed_bytes(93, 27) == image + blob + tables
    assert lib.crnerf_mlp_packed_bytes_op(93, 27, 1) == image + blob + tables
    assert lib.crnerf_mlp_packed_bytes_op(93, 27, 2) == 2 * image + blob + tables     # fp16x3: W_hi and W_lo
    assert lib.crnerf_mlp_packed_bytes_op(93, 27, 3) == 0                             # unknown operand
    assert lib.crnerf_style_scratch_floats(1024) > 296 * 1088
    buf = (ctypes.c_int32 * 4096)()
    n = lib.crnerf_debug_program(93, 27, buf, 4096)
    assert n == 3 + 79 * 11 + 20 * 7 and buf[0] == 79 and buf[1] == 20
    assert lib.crnerf_debug_program(200, 27, buf, 4096) < 0
    assert b"bad argument" in lib.crnerf_last_error()


def test_no_header_symbol_takes_torch_types():
    text = open(HEADER).read()
    assert "torch" not in text.lower().replace("pytorch", "") and "at::" not in text
    assert 'extern "C"' in text
