"""The C-ABI shared objects load without a GPU and export every symbol that include/*.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stringsearchlib_b200", "lib")

HEADERS = {
    "StringSearchLib.h": "libStringSearchLib.so",
    "StringSearchLib_handle.h": "libStringSearchLib_h.so",
    "ssb_engine.h": "libssb_core.so",
}


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = []
    for m in re.finditer(r"^[A-Za-z_][\w\s\*]*?\b(\w+)\s*\(([^;{]*);", text, flags=re.M):
        alias = re.search(r'__asm__\("(\w+)"\)', m.group(2))      # declared under another C name, exported as `alias`
        names.append(alias.group(1) if alias else m.group(1))
    return names


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(os.path.join(LIB, "libssb_core.so")):
        import sys
        sys.path.insert(0, ROOT)
        import build
        build.build()


@pytest.mark.parametrize("header,lib", sorted(HEADERS.items()))
def test_every_declared_symbol_is_exported(header, lib):
    names = declared_functions(header)
    assert len(names) >= 8, names
    so = ctypes.CDLL(os.path.join(LIB, lib), mode=ctypes.RTLD_LOCAL)
    missing = [n for n in names if not hasattr(so, n)]
    assert not missing, f"{lib} lacks {missing}"


def test_expected_surface_names():
    assert set(declared_functions("StringSearchLib_handle.h")) == {
        "indexN", "search", "score", "release", "dispose", "getSize", "getLibSize", "setValidChar"}      # dllmain.cpp:37-151
    assert set(declared_functions("StringSearchLib.h")) == {
        "index2D", "index2DW", "index", "indexW", "search", "searchW", "release", "releaseW", "dispose", "disposeW",
        "getSize", "getSizeW", "getLibSize", "getLibSizeW"}                                              # Readme.md:31-226


def test_no_cpu_fallback_without_gpu():
    """the product fails loudly when CUDA is unavailable: no silent CPU path (skipped where a GPU exists)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import stringsearchlib_b200 as ssb
    with pytest.raises(RuntimeError, match="CUDA"):
        ssb.Index([b"alpha beta", b"gamma delta"])
    H = ssb.HandleLibrary()
    assert H.indexN([b"alpha beta", b"gamma delta"], 1) == 0        # indexN returns 0 on failure (dllmain.cpp:45-46)


def test_product_does_not_reference_the_oracle():
    """nothing under stringsearchlib_b200/ or include/ may import, include or link oracle/ (test infrastructure only)."""
    bad = []
    for base in ("stringsearchlib_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".txt")):
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(import|from)\s+oracle|oracle/_ref|libspec|libref|ngram_spec", t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_host_only_entry_points_behave_without_gpu(tmp_path):
    """entry points that need no device work validate their arguments and fail loudly (never silently) without a GPU"""
    import sys
    sys.path.insert(0, ROOT)
    import stringsearchlib_b200 as ssb
    for path in (-1, 0, 1, -1):
        ssb.set_search_path(path)
    with pytest.raises(ValueError):
        ssb.set_search_path(2)
    junk = tmp_path / "junk.ssbidx"
    junk.write_bytes(b"not an index file" * 8)
    with pytest.raises(RuntimeError):
        ssb.Index.load(str(junk))                      # bad magic on a GPU box, CUDA initialisation error without one
    with pytest.raises(RuntimeError):
        ssb.Index.load(str(tmp_path / "missing.ssbidx"))
