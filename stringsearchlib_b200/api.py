"""api.py — ctypes mirror of include/ssb_engine.h, include/StringSearchLib.h and include/StringSearchLib_handle.h.

Names and argument meaning follow the reference's exports (dllmain.cpp:37-151, Readme.md:25-231). Fails loudly when
the CUDA extension has not been built; nothing here computes on the CPU.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_uint8, c_uint16, c_uint32, c_uint64, c_void_p, c_wchar_p

import numpy as np

_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
_cache = {}


def lib_path(name):
    p = os.path.join(_LIB_DIR, name)
    if not os.path.exists(p):
        raise ImportError(f"{p} is missing: build the CUDA engine first (python build.py, or __graft_entry__.build()). "
                          "There is no CPU fallback.")
    return p


def _load(name):
    if name not in _cache:
        _cache[name] = ctypes.CDLL(lib_path(name), mode=ctypes.RTLD_LOCAL)
    return _cache[name]


class _Batch(ctypes.Structure):
    _fields_ = [("nq", c_uint32), ("offsets", POINTER(c_uint32)), ("kids", POINTER(c_uint32)), ("scores", POINTER(c_float)),
                ("rows", POINTER(c_uint64)), ("strings", POINTER(c_void_p))]


def core_library():
    lib = _load("libssb_core.so")
    if getattr(lib, "_ssb_typed", False):
        return lib
    lib.ssb_last_error.restype = c_char_p
    lib.ssb_kernel_launches.restype = c_uint64
    lib.ssb_set_device.argtypes = [c_int]
    lib.ssb_set_search_path.argtypes = [c_int]
    lib.ssb_index_create.restype = c_void_p
    lib.ssb_index_create.argtypes = [POINTER(c_void_p), c_uint64, c_uint16, POINTER(c_float), c_uint16, c_int, c_uint64]
    lib.ssb_index_create_packed.restype = c_void_p
    lib.ssb_index_create_packed.argtypes = [c_void_p, POINTER(c_uint64), POINTER(c_uint8), c_uint64, c_uint16, POINTER(c_float),
                                            c_uint16, c_int, c_uint64]
    lib.ssb_index_destroy.argtypes = [c_void_p]
    lib.ssb_index_size.restype = c_uint64
    lib.ssb_index_size.argtypes = [c_void_p]
    lib.ssb_index_lib_size.restype = c_uint64
    lib.ssb_index_lib_size.argtypes = [c_void_p]
    lib.ssb_index_set_valid_char.argtypes = [c_void_p, c_char_p, c_int]
    lib.ssb_index_info.argtypes = [c_void_p, POINTER(c_uint64)]
    lib.ssb_index_build_ms.argtypes = [c_void_p, POINTER(c_double)]
    lib.ssb_index_copy.argtypes = [c_void_p, c_int, c_void_p, c_uint64]
    lib.ssb_search_batch.restype = POINTER(_Batch)
    lib.ssb_search_batch.argtypes = [c_void_p, POINTER(c_void_p), c_uint32, c_float, c_uint32]
    lib.ssb_search_batch_packed.restype = POINTER(_Batch)
    lib.ssb_search_batch_packed.argtypes = [c_void_p, c_void_p, POINTER(c_uint32), c_uint32, c_float, c_uint32]
    lib.ssb_batch_free.argtypes = [POINTER(_Batch)]
    lib.ssb_search_batch_device.argtypes = [c_void_p, POINTER(c_void_p), c_uint32, c_float, c_uint32]
    lib.ssb_search_resident.argtypes = [c_void_p, c_uint32, c_float, c_uint32, POINTER(c_float), POINTER(c_uint64)]
    lib.ssb_export_records.argtypes = [c_void_p, c_uint32, c_uint32, c_void_p]
    lib.ssb_merge_records.argtypes = [c_void_p, c_uint32, c_uint32, c_uint32, c_void_p, c_void_p]
    lib.ssb_device_alloc.restype = c_void_p
    lib.ssb_device_alloc.argtypes = [c_uint64]
    lib.ssb_device_free.argtypes = [c_void_p]
    lib.ssb_memcpy_d2h.argtypes = [c_void_p, c_void_p, c_uint64]
    lib.ssb_ipc_get_handle.argtypes = [c_void_p, c_char_p]
    lib.ssb_ipc_open_handle.restype = c_void_p
    lib.ssb_ipc_open_handle.argtypes = [c_char_p]
    lib.ssb_ipc_close_handle.argtypes = [c_void_p]
    lib.ssb_export_records_p2p.argtypes = [c_void_p, c_uint32, c_uint32, POINTER(c_void_p), c_uint32, c_uint32]
    lib.ssb_merge_records_p2p.argtypes = [c_void_p, c_uint32, c_uint32, c_uint32, POINTER(c_void_p), POINTER(c_void_p), c_uint32, c_uint32]
    lib.ssb_export_records_p2p_signal.argtypes = [c_void_p, c_uint32, c_uint32, POINTER(c_void_p), c_uint32, c_uint32, POINTER(c_void_p), c_uint32]
    lib.ssb_merge_records_p2p_wait.argtypes = [c_void_p, c_uint32, c_uint32, c_uint32, POINTER(c_void_p), POINTER(c_void_p), c_uint32, c_uint32,
                                               c_void_p, c_uint32]
    lib.ssb_device_memset.argtypes = [c_void_p, c_int, c_uint64]
    lib.ssb_index_digest_count.restype = c_uint64
    lib.ssb_index_digest_count.argtypes = [c_void_p, c_int]
    lib.ssb_index_export_digests.argtypes = [c_void_p, c_int, c_void_p]
    lib.ssb_count_distinct_digests.argtypes = [c_void_p, c_uint64, c_uint32, c_uint32, POINTER(c_uint64)]
    lib.ssb_search_one.argtypes = [c_void_p, c_void_p, c_float, c_uint32]
    lib.ssb_search_one.restype = POINTER(_Batch)
    lib.ssb_search_one_stats.argtypes = [c_void_p, POINTER(c_uint64)]
    lib.ssb_index_save.argtypes = [c_void_p, c_char_p]
    lib.ssb_index_load.argtypes = [c_char_p]
    lib.ssb_index_load.restype = c_void_p
    lib._ssb_typed = True
    return lib


def kernel_launches():
    return int(core_library().ssb_kernel_launches())


def set_search_path(path):
    """1 = k_sieve_search for every query it can take, 0 = k_fast_search alone, -1 = default (per query by posting volume; 0 when
    SSB_MERGE is set); identical results (include/ssb_engine.h)."""
    if core_library().ssb_set_search_path(int(path)) != 0:
        raise ValueError("search path must be -1, 0 or 1")


def set_device(device):
    if core_library().ssb_set_device(int(device)) != 0:
        raise RuntimeError(core_library().ssb_last_error().decode())


INFO = ("K", "T", "E", "G", "P", "nShort", "symBits", "A", "identity", "uniformW", "gSize", "unit", "indexed", "termUnits",
        "keyUnits", "deviceBytes", "windows", "packed", "multiEdgeKeys", "hashRetries")
RECORD_DTYPE = np.dtype([("scoreBits", "<u4"), ("keyLen", "<u4"), ("row", "<u8"), ("hash", "<u8")])


class SearchResult:
    """CSR result of one batch: offsets [nq+1], kids, scores, rows (+ the key strings on demand)."""

    def __init__(self, offsets, kids, scores, rows, strings):
        self.offsets, self.kids, self.scores, self.rows, self.strings = offsets, kids, scores, rows, strings

    def __len__(self):
        return len(self.offsets) - 1

    def query(self, q):
        a, b = int(self.offsets[q]), int(self.offsets[q + 1])
        return self.kids[a:b], self.scores[a:b], self.rows[a:b], self.strings[a:b]


def _as_pointer_array(strings, wide):
    """list of str/bytes/None -> (ctypes void* array, keep-alive list)."""
    keep, arr = [], (c_void_p * len(strings))()
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


class Index:
    """One indexed library (one key-range shard) on the current CUDA device."""

    def __init__(self, words, row_size=1, weights=None, g_size=3, wide=False, row_base=0):
        """words: list of str/bytes/None, or a synth.Packed (then the packed entry point is used)."""
        self._lib = core_library()
        self.wide = bool(wide)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        wp = None if w is None else w.ctypes.data_as(POINTER(c_float))
        if hasattr(words, "units"):
            self.wide = words.wide
            units = words.units if len(words.units) else np.zeros(1, words.units.dtype)
            self._h = self._lib.ssb_index_create_packed(units.ctypes.data, words.offsets.ctypes.data_as(POINTER(c_uint64)), None,
                                                        len(words), row_size, wp, g_size, int(self.wide), row_base)
        else:
            arr, keep = _as_pointer_array(words, self.wide)
            self._h = self._lib.ssb_index_create(arr, len(words), row_size, wp, g_size, int(self.wide), row_base)
        if not self._h:
            raise RuntimeError("ssb_index_create failed: " + self._lib.ssb_last_error().decode())

    @classmethod
    def from_pointers(cls, ptrs, n, row_size=1, weights=None, g_size=3, wide=False, row_base=0):
        """ptrs: ctypes POINTER(c_void_p) to n char*/wchar_t* (e.g. synth.Packed.pointers()) — the char** entry point."""
        self = cls.__new__(cls)
        self._lib = core_library()
        self.wide = bool(wide)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        wp = None if w is None else w.ctypes.data_as(POINTER(c_float))
        self._h = self._lib.ssb_index_create(ptrs, n, row_size, wp, g_size, int(wide), row_base)
        if not self._h:
            raise RuntimeError("ssb_index_create failed: " + self._lib.ssb_last_error().decode())
        return self

    def save(self, path):
        """writes every table of the index to `path` (ssb_index_save)."""
        if self._lib.ssb_index_save(self._h, os.fsencode(path)) != 0:
            raise RuntimeError(self._lib.ssb_last_error().decode())

    @classmethod
    def load(cls, path):
        """an index that answers exactly like the one saved to `path` (ssb_index_load)."""
        self = cls.__new__(cls)
        self._lib = core_library()
        self._h = self._lib.ssb_index_load(os.fsencode(path))
        if not self._h:
            raise RuntimeError("ssb_index_load failed: " + self._lib.ssb_last_error().decode())
        self.wide = self.info()["unit"] == 4
        return self

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ssb_index_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- getSize / getLibSize (dllmain.cpp:120,133)
    def size(self):
        return int(self._lib.ssb_index_size(self._h))

    def lib_size(self):
        return int(self._lib.ssb_index_lib_size(self._h))

    def set_valid_char(self, chars):
        b = chars if isinstance(chars, bytes) else chars.encode("latin-1")
        self._lib.ssb_index_set_valid_char(self._h, b, len(b))

    def info(self):
        a = (c_uint64 * 24)()
        self._lib.ssb_index_info(self._h, a)
        return {k: int(a[i]) for i, k in enumerate(INFO)}

    def build_ms(self):
        a = (c_double * 8)()
        self._lib.ssb_index_build_ms(self._h, a)
        return dict(total=a[0], pack=a[1], h2d=a[2], intern=a[3], edges=a[4], grams=a[5], d2h=a[6])

    def _copy(self, which, dtype, count):
        out = np.empty(count, dtype=dtype)
        if self._lib.ssb_index_copy(self._h, which, out.ctypes.data, out.nbytes) != 0:
            raise RuntimeError(self._lib.ssb_last_error().decode())
        return out

    def tables(self):
        """Every index table as numpy arrays (for table-level parity against the oracle)."""
        i = self.info()
        udt = np.uint32 if i["unit"] == 4 else np.uint8
        if not i["indexed"] or i["T"] == 0:
            return dict(info=i)
        t = dict(info=i)
        t["gramCodes"] = self._copy(0, np.uint64, i["G"])
        t["gramOff"] = self._copy(1, np.uint32, i["G"] + 1)
        t["postings"] = self._copy(2, np.uint32, i["P"])
        t["termOff"] = self._copy(3, np.uint32, i["T"] + 1)
        t["termChars"] = self._copy(4, udt, i["termUnits"])
        t["termEdgeOff"] = self._copy(5, np.uint32, i["T"] + 1)
        t["edgeKid"] = self._copy(6, np.uint32, i["E"])
        t["edgeW"] = self._copy(7, np.float32, i["E"])
        t["keyOff"] = self._copy(8, np.uint32, i["K"] + 1)
        t["keyChars"] = self._copy(9, udt, i["keyUnits"])
        t["keyRow"] = self._copy(10, np.uint64, i["K"])
        t["symTable"] = self._copy(11, np.uint32, 256 if i["unit"] == 1 else i["A"])
        t["shortTerms"] = self._copy(12, np.uint32, i["nShort"])
        return t

    def gram_units(self, tables):
        """decodes symbol-packed gram codes back to code units: [G, gSize] uint32."""
        i = tables["info"]
        g, bits = i["gSize"], i["symBits"]
        if i["unit"] == 1:
            lut = tables["symTable"]
            inv = np.zeros(257, dtype=np.uint32)
            present = np.nonzero(lut)[0]
            inv[lut[present]] = present
        else:
            inv = np.concatenate([[0], tables["symTable"]]).astype(np.uint32)
        codes = tables["gramCodes"]
        out = np.zeros((len(codes), g), dtype=np.uint32)
        for j in range(g):
            sym = (codes >> np.uint64(bits * (g - 1 - j))) & np.uint64((1 << bits) - 1)
            out[:, j] = inv[sym.astype(np.int64)]
        return out

    # ---- batched search (the extension the metric is measured on)
    def _wrap(self, bp, want_strings):
        if not bp:
            raise RuntimeError("ssb_search_batch failed: " + self._lib.ssb_last_error().decode())
        b = bp.contents
        nq = b.nq
        off = np.ctypeslib.as_array(b.offsets, shape=(nq + 1,)).copy()
        tot = int(off[-1])
        kids = np.ctypeslib.as_array(b.kids, shape=(max(tot, 1),))[:tot].copy()
        scores = np.ctypeslib.as_array(b.scores, shape=(max(tot, 1),))[:tot].copy()
        rows = np.ctypeslib.as_array(b.rows, shape=(max(tot, 1),))[:tot].copy()
        strings = []
        if want_strings:
            for i in range(tot):
                p = b.strings[i]
                strings.append(ctypes.wstring_at(p) if self.wide else ctypes.string_at(p))
        self._lib.ssb_batch_free(bp)
        return SearchResult(off, kids, scores, rows, strings)

    def search_batch(self, queries, threshold=0.0, limit=100, strings=False):
        """queries: list of str/bytes, or a synth.Packed."""
        if hasattr(queries, "units"):
            units = queries.units if len(queries.units) else np.zeros(1, queries.units.dtype)
            off = queries.offsets.astype(np.uint32)
            bp = self._lib.ssb_search_batch_packed(self._h, units.ctypes.data, off.ctypes.data_as(POINTER(c_uint32)), len(queries),
                                                   threshold, limit)
        else:
            arr, keep = _as_pointer_array(queries, self.wide)
            bp = self._lib.ssb_search_batch(self._h, arr, len(queries), threshold, limit)
        return self._wrap(bp, strings)

    def search_batch_pointers(self, ptrs, nq, threshold=0.0, limit=100, strings=False):
        """the char** entry point with caller-owned host pointers (what an FFI host passes)."""
        return self._wrap(self._lib.ssb_search_batch(self._h, ptrs, nq, threshold, limit), strings)

    def search_batch_raw(self, ptrs, nq, threshold=0.0, limit=100):
        """the char** entry point exactly as an FFI host sees it: returns the ssb_batch* (arrays owned by the library,
        valid until free_raw); no numpy copies — this is what bench.py times as e2e."""
        bp = self._lib.ssb_search_batch(self._h, ptrs, nq, threshold, limit)
        if not bp:
            raise RuntimeError("ssb_search_batch failed: " + self._lib.ssb_last_error().decode())
        return bp

    def search_batch_device(self, ptrs, nq, threshold=0.0, limit=100):
        """H2D + kernels; the per-shard results stay in HBM (sharded path: export_records* + cross-shard merge follow)."""
        if self._lib.ssb_search_batch_device(self._h, ptrs, nq, threshold, limit) != 0:
            raise RuntimeError(self._lib.ssb_last_error().decode())

    def free_raw(self, bp):
        self._lib.ssb_batch_free(bp)

    def search(self, query, threshold=0.0, limit=100):
        """single query: [(key string, score)], like score() of the reference (dllmain.cpp:82)."""
        r = self.search_batch([query], threshold, limit, strings=True)
        return list(zip(r.strings, r.scores.tolist()))

    def search_one(self, query, threshold=0.0, limit=100):
        """ssb_search_one: one query through the coalescing front end (what the shims' search()/score() use);
        returns (kids, scores) as numpy arrays."""
        arr, keep = _as_pointer_array([query], self.wide)
        bp = self._lib.ssb_search_one(self._h, arr[0], threshold, limit)
        if not bp:
            raise RuntimeError("ssb_search_one failed: " + self._lib.ssb_last_error().decode())
        b = bp.contents
        n = int(b.offsets[1])
        kids = np.ctypeslib.as_array(b.kids, (max(n, 1),))[:n].copy()
        scores = np.ctypeslib.as_array(b.scores, (max(n, 1),))[:n].copy()
        self._lib.ssb_batch_free(bp)
        return kids, scores

    def search_one_stats(self):
        a = (c_uint64 * 2)()
        self._lib.ssb_search_one_stats(self._h, a)
        return dict(calls=int(a[0]), batches=int(a[1]))

    def search_resident(self, nq, threshold=0.0, limit=100):
        """re-runs the device part of the last uploaded batch; returns (ms_device, ms_fast_kernel, stats dict)."""
        ms = (c_float * 2)()
        st = (c_uint64 * 8)()
        if self._lib.ssb_search_resident(self._h, nq, threshold, limit, ms, st) != 0:
            raise RuntimeError(self._lib.ssb_last_error().decode())
        return ms[0], ms[1], dict(nFast=int(st[0]), nDense=int(st[1]), nOverflow=int(st[2]), postings=int(st[3]),
                                  nSieve=int(st[4]), nSieveHandover=int(st[5]), nSieveGiveUp=int(st[6]), sieveProbes=int(st[7]))

    def export_records_p2p(self, nq, limit, peer_recv, rank):
        """fused export + exchange: query q's records go straight into peer_recv[q // (nq/N)] (peer-mapped pointers)."""
        _check(self._lib.ssb_export_records_p2p(self._h, nq, limit, _ptr_list(peer_recv), len(peer_recv), rank))

    def export_records_p2p_signal(self, nq, limit, peer_recv, rank, peer_flags, epoch):
        """export_records_p2p + device-side signal: peer_flags[j][rank] = epoch once the records have landed on every rank."""
        _check(self._lib.ssb_export_records_p2p_signal(self._h, nq, limit, _ptr_list(peer_recv), len(peer_recv), rank,
                                                       _ptr_list(peer_flags), epoch))

    def digest_count(self, which):
        """16-byte digests this shard exports: which 0 = terms (getSize), 1 = grams (getLibSize)."""
        return int(self._lib.ssb_index_digest_count(self._h, which))

    def export_digests(self, which, device_ptr):
        _check(self._lib.ssb_index_export_digests(self._h, which, device_ptr))

    def export_records(self, nq, limit, device_ptr):
        if self._lib.ssb_export_records(self._h, nq, limit, device_ptr) != 0:
            raise RuntimeError(self._lib.ssb_last_error().decode())


def _check(rc):
    if rc != 0:
        raise RuntimeError(core_library().ssb_last_error().decode())


def device_alloc(nbytes):
    """cudaMalloc'ed (IPC-shareable) device buffer, filled with 0xFF; returns the device address."""
    p = core_library().ssb_device_alloc(int(nbytes))
    if not p:
        raise RuntimeError(core_library().ssb_last_error().decode())
    return int(p)


def device_free(ptr):
    core_library().ssb_device_free(ptr)


def memcpy_d2h(dst_array, ptr, nbytes=None):
    _check(core_library().ssb_memcpy_d2h(dst_array.ctypes.data, ptr, int(nbytes if nbytes is not None else dst_array.nbytes)))


def ipc_get_handle(ptr):
    buf = ctypes.create_string_buffer(64)
    _check(core_library().ssb_ipc_get_handle(ptr, buf))
    return buf.raw


def ipc_open_handle(handle):
    p = core_library().ssb_ipc_open_handle(handle)
    if not p:
        raise RuntimeError(core_library().ssb_last_error().decode())
    return int(p)


def _ptr_list(ptrs):
    return (c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def merge_records_p2p(recv_ptr, n_shards, nq_local, limit, peer_merged, peer_counts, rank):
    """merges this rank's slice of queries and stores the result into every rank's buffers (peer-mapped pointers)."""
    _check(core_library().ssb_merge_records_p2p(recv_ptr, n_shards, nq_local, limit, _ptr_list(peer_merged), _ptr_list(peer_counts),
                                                len(peer_merged), rank))


def merge_records_p2p_wait(recv_ptr, n_shards, nq_local, limit, out_merged, out_counts, rank, my_flags, epoch):
    """like merge_records_p2p, but the kernel itself waits for the peers' export kernels (flags >= epoch); out_merged /
    out_counts: the result buffers to store into (usually just this rank's own)."""
    _check(core_library().ssb_merge_records_p2p_wait(recv_ptr, n_shards, nq_local, limit, _ptr_list(out_merged), _ptr_list(out_counts),
                                                     len(out_merged), rank, my_flags, epoch))


def device_memset(ptr, value, nbytes):
    _check(core_library().ssb_device_memset(ptr, int(value), int(nbytes)))


def count_distinct_digests(digests_ptr, n, part=0, n_parts=1):
    """distinct 16-byte digests among n at digests_ptr (device) whose residue class is `part` of `n_parts`."""
    out = c_uint64(0)
    _check(core_library().ssb_count_distinct_digests(digests_ptr, int(n), part, n_parts, ctypes.byref(out)))
    return int(out.value)


def global_sizes(ix, dist=None, device=None):
    """(getSize, getLibSize) of the whole index when `ix` is this rank's key-range shard: distinct terms
    and grams over all shards. dist = torch.distributed with an initialised process group (None: one shard = the shard's own
    numbers). One all-gather of the digest counts, one of the padded digests and one 8-byte all-reduce per counter."""
    import torch
    if dist is None or dist.get_world_size() == 1:
        return ix.size(), ix.lib_size()
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
    out = []
    for which in (0, 1):
        n = ix.digest_count(which)
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        mine_n = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, mine_n)
        counts = counts.cpu().tolist()
        cap = max(max(counts), 1)
        mine = torch.zeros((cap, 16), dtype=torch.uint8, device=dev)
        if n:
            ix.export_digests(which, mine.data_ptr())
        gathered = torch.empty((world, cap, 16), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, mine)
        # drop the padding: the shards' digests back to back
        flat = torch.cat([gathered[r, :counts[r]] for r in range(world)]).contiguous()
        torch.cuda.synchronize()
        part = count_distinct_digests(flat.data_ptr(), flat.shape[0], rank, world) if flat.shape[0] else 0
        t = torch.tensor([part], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        out.append(int(t.item()))
    return out[0], out[1]


def merge_records(gathered_ptr, n_shards, nq, limit, merged_ptr, counts_ptr):
    lib = core_library()
    if lib.ssb_merge_records(gathered_ptr, n_shards, nq, limit, merged_ptr, counts_ptr) != 0:
        raise RuntimeError(lib.ssb_last_error().decode())


class HandleLibrary:
    """libStringSearchLib_h.so — the as-shipped exports (dllmain.cpp:37-151), same names and signatures."""

    def __init__(self, path=None):
        lib = ctypes.CDLL(path or lib_path("libStringSearchLib_h.so"), mode=ctypes.RTLD_LOCAL)
        lib.indexN.restype = c_uint32
        lib.indexN.argtypes = [POINTER(c_char_p), c_uint64, c_uint16, POINTER(c_float)]
        lib.search.restype = c_uint32
        lib.search.argtypes = [c_uint32, c_char_p, POINTER(POINTER(c_char_p)), c_float, c_uint32]
        lib.score.restype = c_uint32
        lib.score.argtypes = [c_uint32, c_char_p, POINTER(POINTER(c_char_p)), POINTER(POINTER(c_float)), c_float, c_uint32]
        lib.release.argtypes = [c_uint32, POINTER(c_char_p), POINTER(c_float)]
        lib.dispose.argtypes = [c_uint32]
        lib.getSize.restype = c_uint64
        lib.getSize.argtypes = [c_uint32]
        lib.getLibSize.restype = c_uint64
        lib.getLibSize.argtypes = [c_uint32]
        lib.setValidChar.argtypes = [c_uint32, c_char_p, c_int]
        self.lib = lib

    def indexN(self, words, row_size=1, weights=None):
        arr = (c_char_p * len(words))(*[None if w is None else (w if isinstance(w, bytes) else w.encode("latin-1")) for w in words])
        w = None if weights is None else (c_float * len(weights))(*weights)
        return int(self.lib.indexN(arr, len(words), row_size, w))

    def score(self, handle, query, threshold=0.0, limit=0):
        res, sc = POINTER(c_char_p)(), POINTER(c_float)()
        q = query if isinstance(query, bytes) else query.encode("latin-1")
        n = self.lib.score(handle, q, ctypes.byref(res), ctypes.byref(sc), threshold, limit)
        out = [(res[i], sc[i]) for i in range(n)]
        if n or res:
            self.lib.release(handle, res, sc)
        return out

    def search(self, handle, query, threshold=0.0, limit=0):
        res = POINTER(c_char_p)()
        q = query if isinstance(query, bytes) else query.encode("latin-1")
        n = self.lib.search(handle, q, ctypes.byref(res), threshold, limit)
        out = [res[i] for i in range(n)]
        if n or res:
            self.lib.release(handle, res, None)
        return out


class GuidLibrary:
    """libStringSearchLib.so — the README exports (Readme.md:31-226), same names and signatures."""

    def __init__(self, path=None):
        lib = ctypes.CDLL(path or lib_path("libStringSearchLib.so"), mode=ctypes.RTLD_LOCAL)
        lib.index.argtypes = [c_char_p, POINTER(c_char_p), c_uint64, c_uint16, POINTER(c_float), c_uint16]
        lib.indexW.argtypes = [c_char_p, POINTER(c_wchar_p), c_uint64, c_uint16, POINTER(c_float), c_uint16]
        lib.index2D.argtypes = [c_char_p, POINTER(POINTER(c_char_p)), c_uint64, c_uint16, POINTER(POINTER(c_float)), c_uint16]
        lib.index2DW.argtypes = [c_char_p, POINTER(POINTER(c_wchar_p)), c_uint64, c_uint16, POINTER(POINTER(c_float)), c_uint16]
        lib.search.argtypes = [c_char_p, c_char_p, POINTER(POINTER(c_char_p)), POINTER(c_uint32), c_float, c_uint32]
        lib.searchW.argtypes = [c_char_p, c_wchar_p, POINTER(POINTER(c_wchar_p)), POINTER(c_uint32), c_float, c_uint32]
        lib.release.argtypes = [c_char_p, POINTER(POINTER(c_char_p)), c_uint64]
        lib.releaseW.argtypes = [c_char_p, POINTER(POINTER(c_wchar_p)), c_uint64]
        for f in ("dispose", "disposeW"):
            getattr(lib, f).argtypes = [c_char_p]
        for f in ("getSize", "getSizeW", "getLibSize", "getLibSizeW"):
            getattr(lib, f).argtypes = [c_char_p]
            getattr(lib, f).restype = c_uint64
        self.lib = lib

    def index(self, guid, words, row_size=1, weights=None, g_size=3):
        arr = (c_char_p * len(words))(*[None if w is None else (w if isinstance(w, bytes) else w.encode("latin-1")) for w in words])
        w = None if weights is None else (c_float * len(weights))(*weights)
        self.lib.index(guid, arr, len(words), row_size, w, g_size)

    def indexW(self, guid, words, row_size=1, weights=None, g_size=3):
        arr = (c_wchar_p * len(words))(*words)
        w = None if weights is None else (c_float * len(weights))(*weights)
        self.lib.indexW(guid, arr, len(words), row_size, w, g_size)

    def index2D(self, guid, rows, weights=None, g_size=3, wide=False):
        ct = c_wchar_p if wide else c_char_p
        row_size = len(rows[0])
        rarr = [(ct * row_size)(*[(c if (wide or c is None or isinstance(c, bytes)) else c.encode("latin-1")) for c in r]) for r in rows]
        karr = (POINTER(ct) * len(rows))(*[ctypes.cast(r, POINTER(ct)) for r in rarr])
        warr = None
        if weights is not None:
            wr = [(c_float * row_size)(*w) for w in weights]
            warr = (POINTER(c_float) * len(rows))(*[ctypes.cast(w, POINTER(c_float)) for w in wr])
        (self.lib.index2DW if wide else self.lib.index2D)(guid, karr, len(rows), row_size, warr, g_size)

    def search(self, guid, query, threshold=0.0, limit=100):
        res, n = POINTER(c_char_p)(), c_uint32(0)
        q = query if isinstance(query, bytes) else query.encode("latin-1")
        self.lib.search(guid, q, ctypes.byref(res), ctypes.byref(n), threshold, limit)
        out = [res[i] for i in range(n.value)]
        if res:
            self.lib.release(guid, ctypes.byref(res), n.value)
        return out

    def searchW(self, guid, query, threshold=0.0, limit=100):
        res, n = POINTER(c_wchar_p)(), c_uint32(0)
        self.lib.searchW(guid, query, ctypes.byref(res), ctypes.byref(n), threshold, limit)
        out = [res[i] for i in range(n.value)]
        if res:
            self.lib.releaseW(guid, ctypes.byref(res), n.value)
        return out
