"""The C-ABI boundary without a GPU: include/nsm_b200.h <-> libnsm_b200.so exports <-> ctypes mirror; loud failure
when no device exists (no CPU fallback anywhere in the product path)."""
import os
import re
import subprocess

import pytest

from tests.conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "nsm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nsm_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_library_and_mirror_agree():
    from nimblesm_b200 import capi

    hdr = _header_symbols()
    assert sorted(capi.SYMBOLS) == hdr
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (nsm_b200_[a-z0-9_]+)", out))
    assert set(hdr) <= exported, sorted(set(hdr) - exported)
    L = capi.lib()
    for s in hdr:
        assert hasattr(L, s)
    assert "fmad=off" in capi.version() and "sm_100a" in capi.version()


def test_library_contains_only_sm100a_code():
    from nimblesm_b200 import capi

    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_fallback_without_a_device():
    """On a box without a GPU every entry fails loudly; on a GPU box this test only checks argument errors."""
    import ctypes as C

    from nimblesm_b200 import capi

    L = capi.lib()
    h = C.c_void_p()
    rc = L.nsm_b200_create(9999, C.byref(h))
    assert rc in (capi.ERR_CUDA, capi.ERR_ARG) and not h.value
    assert L.nsm_b200_last_error(None)
    try:
        ctx = capi.Context(0)
    except capi.NsmError as e:
        assert e.code == capi.ERR_CUDA and "no CPU path" in str(e)
    else:
        ctx.close()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nimblesm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "hex8_oracle" not in text and "libnimble_ref" not in text, f
