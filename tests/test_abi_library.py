"""The C-ABI shared library loads and exports exactly what include/sse_b200.h declares; without a
GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from sse_b200 import _abi, _lib, cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sse_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sse_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert _lib.load().sse_abi_version() == _abi.SSE_ABI_VERSION


def test_struct_layout_matches_header():
    # field order/size of the ctypes mirror: compile-time check through a tiny C program
    import subprocess, tempfile, textwrap
    src = textwrap.dedent(f"""
        #include <stdio.h>
        #include "{ROOT}/include/sse_b200.h"
        int main(void) {{ printf("%zu %zu %zu %zu\\n", sizeof(sse_config), sizeof(sse_arrays),
                          offsetof(sse_config, half_lambda), offsetof(sse_arrays, mapP)); return 0; }}
    """)
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-o", os.path.join(td, "t"), os.path.join(td, "t.c")])
        out = subprocess.check_output([os.path.join(td, "t")]).split()
    assert [int(x) for x in out] == [C.sizeof(_abi.sse_config), C.sizeof(_abi.sse_arrays),
                                     _abi.sse_config.half_lambda.offset, _abi.sse_arrays.mapP.offset]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    img = cases.advection_2d(M=2).image()
    h = C.c_void_p()
    arr = img.c_arrays()
    rc = _lib.load().sse_create(C.byref(img.cfg), C.byref(arr), 0, C.byref(h))
    assert rc == _abi.SSE_ERR_CUDA and not h
    assert b"no CPU fallback" in _lib.load().sse_last_error_string()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cloud.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|libsse_oracle|oracle[/\\]|sse_oracle", txt), os.path.join(dp, f)
