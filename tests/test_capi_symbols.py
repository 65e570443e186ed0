"""The C-ABI library loads without a GPU and exports every symbol include/kdme_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kdme_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:jbf|buf2d|kdme)_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = declared_symbols()
    for must in ("jbf_create", "jbf_process", "jbf_process_batch", "jbf_filtered_device", "jbf_filtered_host",
                 "jbf_smooth_device", "jbf_destroy", "kdme_last_error", "jbf_upsample", "kdme_guided_fill",
                 "buf2d_create", "buf2d_insert_f32", "buf2d_insert_dw", "buf2d_insert_f32x2", "buf2d_update_f32",
                 "buf2d_get_depth", "buf2d_get_weight", "buf2d_raw"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from kinectdepthmapenhancement_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(L, name), f"{name} declared in include/kdme_b200.h but not exported"
    assert set(declared_symbols()) == set(_lib.SIGNATURES), "python binding table out of sync with the header"
    assert b"sm_100a" in _lib.lib().kdme_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import kinectdepthmapenhancement_b200 as k
    with pytest.raises(RuntimeError):
        k.JointBilateralFilter(640, 480)
    with pytest.raises(RuntimeError):
        k.Buffer2D(640, 480)
    # straight through the C ABI: a clean error, not a crash and not a CPU result
    from kinectdepthmapenhancement_b200 import _lib
    h = ctypes.c_void_p()
    rc = _lib.lib().jbf_create(ctypes.byref(h), 640, 480, 70.0, 50.0, 20.0, 2, 1, 0, None)
    assert rc != 0 and not h.value and _lib.lib().kdme_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "kinectdepthmapenhancement_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert not re.search(r"#\s*include[^\n]*oracle", src), f
                assert "libkdme_oracle" not in src and "libkdme_ref" not in src, f
