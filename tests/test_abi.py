"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/asw_b200.h declares, fills the reference's literals as defaults, and fails loudly
(no fallback) when no CUDA device is present.  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, have_gpu


@pytest.fixture(scope="module")
def lib():
    from stereo_matchin_b200 import api, build
    build.build_lib()            # nvcc cross-compiles without a GPU
    return api.load_library()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "asw_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"ASW_API\s+[\w\s\*]+?\b(asw_\w+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    from stereo_matchin_b200 import api
    syms = header_symbols()
    assert len(syms) >= 25
    assert sorted(api.EXPORTS) == syms, "api.EXPORTS must list exactly the header's entry points"


def test_library_exports_every_header_symbol(lib):
    for s in header_symbols():
        assert hasattr(lib, s), f"libasw_b200.so does not export {s}"


def test_signatures_have_no_torch_types():
    src = open(os.path.join(ROOT, "include", "asw_b200.h")).read()
    assert "torch" not in src.lower() and "at::" not in src and "std::" not in src
    assert 'extern "C"' in src


def test_defaults_are_the_reference_literals(lib):
    from stereo_matchin_b200 import api
    p = api.default_params()
    # asw_vsupport.cl:19,22,24; asw_aggr.cl:16; main.cpp:177
    assert (p.radius, p.ndisp, p.iterations) == (16, 61, 7)
    assert abs(p.gamma_c - 30.91) < 1e-6 and abs(p.gamma_p - 28.21) < 1e-6
    assert p.trunc == float("inf")


def test_strerror_and_version(lib):
    assert lib.asw_strerror(0) == b"ok"
    assert b"invalid" in lib.asw_strerror(1)
    assert b"sm_100a" in lib.asw_version()


def test_null_arguments_are_rejected(lib):
    assert lib.asw_create(None, 0) == 1
    assert lib.asw_destroy(None) == 1
    assert lib.asw_sync(None) == 1


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly(lib):
    """Without a CUDA device the product refuses to run; it never routes to a CPU path."""
    from stereo_matchin_b200 import api
    h = C.c_void_p()
    assert lib.asw_create(C.byref(h), 0) == api.ASW_ERR_CUDA and not h.value
    with pytest.raises(api.AswError):
        api.AswContext(0)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing in the product package may reference it."""
    pkg = os.path.join(ROOT, "stereo_matchin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libasw_oracle" not in txt, f
    for f in os.listdir(os.path.join(ROOT, "src", "host")):
        if f.endswith((".cpp", ".h")):
            assert "oracle" not in open(os.path.join(ROOT, "src", "host", f)).read().lower(), f
