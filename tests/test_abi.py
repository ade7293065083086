"""The C-ABI library builds, loads and exports every symbol include/pof.h declares.
No compute call is made here (no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from planar_optical_flow_b200 import build

    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pof.h")).read()
    return sorted(set(re.findall(r"POF_API\s+[\w\s\*]+?\b(pof_\w+)\s*\(", text)))


def test_header_declares_the_three_stages():
    syms = declared_symbols()
    for s in ("pof_cutout_fwd", "pof_spaam_gate_fwd", "pof_spaam_gate_bwd", "pof_nms_centers",
              "pof_cutout_ws_bytes", "pof_nms_ws_bytes", "pof_last_error", "pof_abi_version"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    h = ctypes.CDLL(built_lib)
    for s in declared_symbols():
        assert hasattr(h, s), s


def test_binding_table_matches_header(built_lib):
    from planar_optical_flow_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    L = _lib.lib()
    assert L.pof_abi_version() == _lib.ABI_VERSION
    # pure host arithmetic entry points (no device needed)
    assert L.pof_cutout_ws_bytes(256) == 256 * 8
    assert L.pof_cutout_ws_bytes(0) == 0
    assert L.pof_nms_ws_bytes(1, 1091) >= 2 * 1091 * 8 + 1091 * 35 * 4
    assert L.pof_spaam_gate_bwd_ws_bytes(2, 450, 11) == 2 * 450 * 11 * 4


def test_argument_validation_happens_before_any_cuda_call(built_lib):
    """Bad arguments are rejected with a negative status and a message, GPU or not."""
    from planar_optical_flow_b200 import _lib

    L = _lib.lib()
    st = L.pof_cutout_fwd(None, None, 0, 1, 1, 450, 1, 56, 1.0, 0.5, 29.99, 1, 1, 1, 0, None, None, None, None, None, 0, None)
    assert st == -1 and "null" in _lib.last_error()
    one = ctypes.c_void_p(256)
    st = L.pof_cutout_fwd(one, one, 0, 1, 1, 450, 1, 55, 1.0, 0.5, 29.99, 1, 1, 1, 0, one, None, None, None, one, 8, None)
    assert st == -2 and "multiple of 4" in _lib.last_error()
    st = L.pof_cutout_fwd(one, one, 0, 1, 1, 450, 1, 56, 1.0, 0.5, 29.99, 1, 1, 1, 3, one, None, None, None, one, 8, None)
    assert st < 0 and "numerics" in _lib.last_error()          # 0 EXACT, 1 FAST, 2 EXACT_PIECES
    st = L.pof_spaam_gate_fwd(one, one, one, one, 1, 10, 3584, 128, 10, 0.5, ctypes.c_void_p(512), one, None, None, 0, None, None)
    assert st == -5 and "odd" in _lib.last_error()
    st = L.pof_spaam_gate_fwd(one, one, one, one, 1, 10, 3584, 128, 11, 0.5, one, one, None, None, 0, None, None)
    assert st == -3 and "alias" in _lib.last_error()
    st = L.pof_spaam_gate_fwd(one, one, one, one, 1, 10, 3584, 128, 11, 0.5, ctypes.c_void_p(512), one, None, one, 100, None, None)
    assert st == -2 and "split_channels" in _lib.last_error()
    st = L.pof_nms_centers(one, 0, one, 1, one, one, 1, 5000, 0.5, one, one, one, one, one, one, one, 1 << 30, None)
    assert st == -2


def test_product_refuses_cpu_tensors():
    import numpy as np
    import torch

    from planar_optical_flow_b200 import ops, utils

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.cutout(torch.zeros(1, 1, 450), torch.zeros(450))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            utils.scans_to_cutout(np.ones((1, 450), np.float32), utils.get_laser_phi())


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "planar_optical_flow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), os.path.join(dirpath, f)
