"""CPU: the C-ABI shared library loads and exports every symbol include/robot_gnn.h declares, and the ctypes
binding lists exactly those symbols.  No compute calls (no GPU here) except argument-validation paths."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "robot_gnn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rgnn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(R):
    syms = header_symbols()
    assert len(syms) >= 22
    l = ctypes.CDLL(R._lib.LIB_PATH)
    for s in syms:
        assert hasattr(l, s), f"{s} declared in include/robot_gnn.h but not exported"


def test_binding_covers_header(R):
    bound = set(R._lib.SIGNATURES) | {"rgnn_last_error", "rgnn_launch_count"}
    assert bound == set(header_symbols())


def test_version_and_arch(R):
    l = R._lib.lib()
    assert l.rgnn_version() == 100 and l.rgnn_arch() == 100


def test_argument_validation_without_gpu(R):
    l = R._lib.lib()
    rc = l.rgnn_aggregate(7, None, 0, None, 0, None, 0, None, None, 0, None, 0, None, None, None, None, None, 0, 0, 0, 0, 1, 0, 1, 1, None)
    assert rc == -1 and b"dtype" in l.rgnn_last_error()
    rc = l.rgnn_gemm(0, None, 0, 0, 0, 0, 0, None, 0, 0, 0, 0, 0, None, 0, 0, 0, 4, 4, 4, 1, 1, 1, None, 0, None, None, 1.0, 0.0, None)
    assert rc == -1
    rc = l.rgnn_build_batched_csr(None, None, 2 ** 31, 0, None, None, 1, *([None] * 13))
    assert rc == -3
    with pytest.raises(RuntimeError):
        R._lib.check(rc, "rgnn_build_batched_csr")
