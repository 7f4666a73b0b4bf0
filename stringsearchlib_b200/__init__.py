"""stringsearchlib_b200 — B200-native n-gram fuzzy search behind the StringSearchLib C ABI.

The product is the three shared objects under stringsearchlib_b200/lib (built by build.py / CMake):
libssb_core.so (engine), libStringSearchLib.so (README guid surface), libStringSearchLib_h.so (as-shipped handle
surface). This package is only a thin ctypes mirror of those C ABIs for tests and benchmarks; it has NO CPU
fallback: importing works without a GPU (so that symbols can be checked), every compute call needs a B200.
"""
from .api import (GuidLibrary, HandleLibrary, Index, SearchResult, core_library, kernel_launches, lib_path,  # noqa: F401
                  set_device, set_search_path)
