"""oracle/loader.py — ctypes loaders for the CPU oracle. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
  SpecIndex  oracle/_build/libspec.so     the CPU restatement (oracle/spec), deterministic ids
  RefLib     oracle/_ref/libref.so        the UNMODIFIED reference compiled by oracle/Makefile (dllmain.cpp:37-151)
  ref_tables oracle/_ref/libref_dump.so   the reference's private tables, ids mapped back to strings
Nothing here reads /root/reference at run time; `make -C oracle ref` (run by __graft_entry__.build() where the
reference tree exists) produces oracle/_ref ahead of time.
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_uint16, c_uint32, c_uint64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SPEC_SO = os.path.join(HERE, "_build", "libspec.so")
REF_SO = os.path.join(HERE, "_ref", "libref.so")
REF_DUMP_SO = os.path.join(HERE, "_ref", "libref_dump.so")


def build(ref=True):
    subprocess.run(["make", "-s", "-C", HERE, "spec"] + (["ref"] if ref else []), check=True)


def have_ref():
    return os.path.exists(REF_SO) and os.path.exists(REF_DUMP_SO)


_spec = None


def spec_lib():
    global _spec
    if _spec is None:
        if not os.path.exists(SPEC_SO):
            build(ref=False)
        L = ctypes.CDLL(SPEC_SO, mode=ctypes.RTLD_LOCAL)
        L.spec_create.restype = c_void_p
        L.spec_create.argtypes = [POINTER(c_void_p), c_uint64, c_uint16, POINTER(c_float), c_uint16, c_int, c_uint64]
        L.spec_free.argtypes = [c_void_p]
        for f in ("spec_size", "spec_lib_size", "spec_num_keys", "spec_num_edges", "spec_num_postings", "spec_num_short",
                  "spec_longest", "spec_key_chars_total", "spec_term_chars_total"):
            getattr(L, f).restype = c_uint64
            getattr(L, f).argtypes = [c_void_p]
        L.spec_indexed.argtypes = [c_void_p]
        L.spec_keys.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        L.spec_terms.argtypes = [c_void_p, c_void_p, c_void_p]
        L.spec_edges.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
        L.spec_grams.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
        L.spec_long_counts.restype = c_uint32
        L.spec_long_counts.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_uint32]
        L.spec_search.restype = c_uint32
        L.spec_search.argtypes = [c_void_p, c_void_p, c_float, c_uint32, c_void_p, c_void_p, c_uint32]
        L.spec_search_batch.argtypes = [c_void_p, POINTER(c_void_p), c_uint32, c_float, c_uint32, c_void_p, c_void_p, c_void_p,
                                        c_uint32, c_int]
        L.spec_set_valid_char.argtypes = [c_void_p, c_char_p, c_int]
        _spec = L
    return _spec


def _ptr_array(strings, wide):
    keep, arr = [], (c_void_p * max(len(strings), 1))()
    for i, s in enumerate(strings):
        if s is None:
            arr[i] = None
            continue
        if wide:
            b = ctypes.create_unicode_buffer(s if isinstance(s, str) else s.decode("latin-1"))
        else:
            b = ctypes.create_string_buffer(s if isinstance(s, bytes) else s.encode("latin-1"))
        keep.append(b)
        arr[i] = ctypes.addressof(b)
    return arr, keep


class SpecIndex:
    """CPU restatement (oracle/spec/ngram_spec.hpp). words: list of str/bytes/None or a synth.Packed."""

    def __init__(self, words, row_size=1, weights=None, g_size=3, wide=False, row_base=0):
        self.L = spec_lib()
        self.wide = bool(wide)
        self.g = g_size
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        wp = None if w is None else w.ctypes.data_as(POINTER(c_float))
        if hasattr(words, "units"):
            self.wide = words.wide
            self._keep = words
            ptrs, n = words.pointers(), len(words)
        else:
            ptrs, self._keep = _ptr_array(words, self.wide)
            n = len(words)
        self.h = self.L.spec_create(ptrs, n, row_size, wp, g_size, int(self.wide), row_base)

    def close(self):
        if getattr(self, "h", None):
            self.L.spec_free(self.h)
            self.h = None

    __del__ = close

    def size(self):
        return int(self.L.spec_size(self.h))

    def lib_size(self):
        return int(self.L.spec_lib_size(self.h))

    def set_valid_char(self, chars):
        b = chars if isinstance(chars, bytes) else chars.encode("latin-1")
        self.L.spec_set_valid_char(self.h, b, len(b))

    def tables(self):
        L, h = self.L, self.h
        udt = np.uint32 if self.wide else np.uint8
        K, T, E = int(L.spec_num_keys(h)), int(L.spec_size(h)), int(L.spec_num_edges(h))
        G, P = int(L.spec_lib_size(h)), int(L.spec_num_postings(h))
        t = dict(K=K, T=T, E=E, G=G, P=P, nShort=int(L.spec_num_short(h)), indexed=bool(L.spec_indexed(h)))
        keyOff = np.zeros(K + 1, np.uint64)
        keyChars = np.zeros(max(int(L.spec_key_chars_total(h)), 1), udt)
        kidOfExt = np.zeros(max(K, 1), np.uint32)
        firstRow = np.zeros(max(K, 1), np.uint64)
        L.spec_keys(h, keyOff.ctypes.data, keyChars.ctypes.data, kidOfExt.ctypes.data, firstRow.ctypes.data)
        t.update(keyOffExt=keyOff, keyCharsExt=keyChars, kidOfExt=kidOfExt[:K], firstRowExt=firstRow[:K])
        termOff = np.zeros(T + 1, np.uint64)
        termChars = np.zeros(max(int(L.spec_term_chars_total(h)), 1), udt)
        L.spec_terms(h, termOff.ctypes.data, termChars.ctypes.data)
        t.update(termOff=termOff, termChars=termChars[:int(termOff[-1])])
        eoff = np.zeros(T + 1, np.uint32)
        ekid = np.zeros(max(E, 1), np.uint32)
        ew = np.zeros(max(E, 1), np.float32)
        L.spec_edges(h, eoff.ctypes.data, ekid.ctypes.data, ew.ctypes.data)
        t.update(termEdgeOff=eoff, edgeKid=ekid[:E], edgeW=ew[:E])
        units = np.zeros(max(G * self.g, 1), np.uint32)
        goff = np.zeros(G + 1, np.uint64)
        post = np.zeros(max(P, 1), np.uint32)
        L.spec_grams(h, units.ctypes.data, goff.ctypes.data, post.ctypes.data)
        t.update(gramUnits=units[:G * self.g].reshape(G, self.g), gramOff=goff, postings=post[:P])
        return t

    def _q(self, query):
        if self.wide:
            return ctypes.create_unicode_buffer(query if isinstance(query, str) else query.decode("latin-1"))
        return ctypes.create_string_buffer(query if isinstance(query, bytes) else query.encode("latin-1"))

    def search(self, query, threshold=0.0, limit=100, cap=None):
        """[(ext key id, score)]"""
        q = self._q(query)
        cap = cap or (limit if 0 < limit < (1 << 20) else max(int(self.L.spec_num_keys(self.h)), 1))
        ext = np.zeros(cap, np.uint32)
        sc = np.zeros(cap, np.float32)
        n = self.L.spec_search(self.h, ctypes.addressof(q), threshold, limit, ext.ctypes.data, sc.ctypes.data, cap)
        n = min(n, cap)
        return ext[:n], sc[:n]

    def long_counts(self, query, cap=1 << 22):
        q = self._q(query)
        t = np.zeros(cap, np.uint32)
        c = np.zeros(cap, np.uint32)
        n = self.L.spec_long_counts(self.h, ctypes.addressof(q), t.ctypes.data, c.ctypes.data, cap)
        return t[:n], c[:n]

    def search_batch(self, queries, threshold=0.0, limit=100, threads=8):
        """queries: list or synth.Packed -> (counts [nq], ext [nq, limit], scores [nq, limit])"""
        if hasattr(queries, "units"):
            ptrs, nq, keep = queries.pointers(), len(queries), queries
        else:
            ptrs, keep = _ptr_array(queries, self.wide)
            nq = len(queries)
        stride = limit if 0 < limit < (1 << 20) else max(int(self.L.spec_num_keys(self.h)), 1)
        counts = np.zeros(nq, np.uint32)
        ext = np.zeros((nq, stride), np.uint32)
        sc = np.zeros((nq, stride), np.float32)
        self.L.spec_search_batch(self.h, ptrs, nq, threshold, limit, counts.ctypes.data, ext.ctypes.data, sc.ctypes.data, stride, threads)
        return np.minimum(counts, stride), ext, sc


class RefLib:
    """The compiled, unmodified reference (oracle/_ref/libref.so): indexN/search/score/release/dispose/getSize/getLibSize."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " missing: run `make -C oracle ref` where /root/reference exists")
        L = ctypes.CDLL(REF_SO, mode=ctypes.RTLD_LOCAL)
        L.indexN.restype = c_uint32
        L.indexN.argtypes = [c_void_p, c_uint64, c_uint16, POINTER(c_float)]
        L.score.restype = c_uint32
        L.score.argtypes = [c_uint32, c_char_p, POINTER(POINTER(c_char_p)), POINTER(POINTER(c_float)), c_float, c_uint32]
        L.search.restype = c_uint32
        L.search.argtypes = [c_uint32, c_char_p, POINTER(POINTER(c_char_p)), c_float, c_uint32]
        L.release.argtypes = [c_uint32, POINTER(c_char_p), POINTER(c_float)]
        L.dispose.argtypes = [c_uint32]
        L.getSize.restype = c_uint64
        L.getSize.argtypes = [c_uint32]
        L.getLibSize.restype = c_uint64
        L.getLibSize.argtypes = [c_uint32]
        L.setValidChar.argtypes = [c_uint32, c_char_p, c_int]
        self.L = L
        self._keep = {}

    def indexN(self, words, row_size=1, weights=None):
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        wp = None if w is None else w.ctypes.data_as(POINTER(c_float))
        if hasattr(words, "units"):
            ptrs, n, keep = words.pointers(), len(words), words
        else:
            ptrs, keep = _ptr_array(words, False)
            n = len(words)
        h = int(self.L.indexN(ptrs, n, row_size, wp))
        self._keep[h] = (keep, w)
        return h

    def score(self, h, query, threshold=0.0, limit=0):
        res, sc = POINTER(c_char_p)(), POINTER(c_float)()
        q = query if isinstance(query, bytes) else query.encode("latin-1")
        n = self.L.score(h, q, ctypes.byref(res), ctypes.byref(sc), threshold, limit)
        out = [(res[i], sc[i]) for i in range(n)]
        self.L.release(h, res, sc)
        return out

    def dispose(self, h):
        self.L.dispose(h)
        self._keep.pop(h, None)


def ref_tables(words, row_size=1, weights=None):
    """The reference's own tables, ids mapped to strings: (terms {term: {key: weight}}, grams {int32: set(term)}, counts)."""
    L = ctypes.CDLL(REF_DUMP_SO, mode=ctypes.RTLD_LOCAL)
    L.refdump_create.restype = c_void_p
    L.refdump_create.argtypes = [c_void_p, c_uint64, c_uint16, POINTER(c_float)]
    L.refdump_tables.restype = c_void_p
    L.refdump_tables.argtypes = [c_void_p]
    L.refdump_release.argtypes = [c_void_p]
    L.refdump_free.argtypes = [c_void_p]
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
    wp = None if w is None else w.ctypes.data_as(POINTER(c_float))
    if hasattr(words, "units"):
        ptrs, n = words.pointers(), len(words)
    else:
        ptrs, keep = _ptr_array(words, False)
        n = len(words)
    h = L.refdump_create(ptrs, n, row_size, wp)
    p = L.refdump_tables(h)
    text = ctypes.string_at(p).decode()
    L.refdump_release(p)
    L.refdump_free(h)

    def unhex(s):
        return b"" if s == "-" else bytes.fromhex(s)

    terms, grams, counts = {}, {}, None
    cur = None
    for line in text.splitlines():
        f = line.split(" ")
        if f[0] == "N":
            counts = dict(terms=int(f[1]), grams=int(f[2]), long=int(f[3]), short=int(f[4]), longest=int(f[5]))
        elif f[0] == "T":
            cur = terms.setdefault(unhex(f[1]), {})
        elif f[0] == "E":
            cur[unhex(f[1])] = float.fromhex(f[2])
        elif f[0] == "G":
            cur = grams.setdefault(int(f[1]), set())
        elif f[0] == "P":
            cur.add(unhex(f[1]))
    return terms, grams, counts
