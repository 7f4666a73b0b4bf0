"""synth.py — synthetic vocabularies and query batches, shared by tests/ and bench.py.

PRNG: splitmix64, counter-based so that any row range can be generated independently (key-range shards):
the j-th draw of item i is mix(seed + (i * STRIDE + j + 1) * GAMMA).
  ASCII key   : len = 6 + r0 % 11 (6..16), chars 'a' + r % 26                              (seed 0x5EED)
  description : one word of 4..10 letters + three times (' ' + word of 3..9 letters)       (C2, rowSize 2)
  weights     : 0.5 + (r % 1000) / 1000 per WORD, float32                                   (C2)
  wide key    : len = 4 + r0 % 9 (4..12) code points uniform in U+4E00..U+9FFF             (C3)
  query       : key r0 % nRows with ONE position overwritten by a random alphabet symbol   (seed 0xC0FFEE)
"""
import ctypes

import numpy as np

GAMMA = np.uint64(0x9E3779B97F4A7C15)
KEY_SEED = 0x5EED
QUERY_SEED = 0xC0FFEE
STRIDE = 64


def _mix(z):
    z = z.astype(np.uint64, copy=True)
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def draws_at(seed, items, ndraw):
    """[len(items), ndraw] uint64 draws for arbitrary item indices."""
    with np.errstate(over="ignore"):
        idx = (np.asarray(items, dtype=np.uint64)[:, None] * np.uint64(STRIDE)
               + np.arange(1, ndraw + 1, dtype=np.uint64)[None, :])
        return _mix(np.uint64(seed) + idx * GAMMA)


def draws(seed, first, count, ndraw):
    """[count, ndraw] uint64 draws for items first..first+count."""
    return draws_at(seed, np.arange(first, first + count, dtype=np.uint64), ndraw)


class Packed:
    """CSR-packed strings: units (uint8 or uint32), offsets uint64 [n+1]; NUL-terminated copy on demand for char**."""

    def __init__(self, units, offsets):
        self.units = np.ascontiguousarray(units)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.n = len(self.offsets) - 1
        self.wide = self.units.dtype == np.uint32
        self._z = None
        self._ptrs = None

    def __len__(self):
        return self.n

    def get(self, i):
        u = self.units[int(self.offsets[i]):int(self.offsets[i + 1])]
        return "".join(map(chr, u.tolist())) if self.wide else u.tobytes()

    def zero_terminated(self):
        """(buffer with a 0 unit after every string, start offsets in units)."""
        if self._z is None:
            lens = np.diff(self.offsets).astype(np.int64)
            starts = self.offsets[:-1].astype(np.int64) + np.arange(self.n, dtype=np.int64)
            buf = np.zeros(int(self.offsets[-1]) + self.n, dtype=self.units.dtype)
            if len(self.units):
                dst = np.arange(len(self.units), dtype=np.int64) + np.repeat(np.arange(self.n, dtype=np.int64), lens)
                buf[dst] = self.units
            self._z = (buf, starts)
        return self._z

    def pointers(self):
        """ctypes array of pointers (char** / wchar_t**) into the zero-terminated buffer (kept alive by self)."""
        if self._ptrs is None:
            buf, starts = self.zero_terminated()
            addr = (buf.ctypes.data + starts * buf.itemsize).astype(np.uint64)
            self._ptrs = addr
        return self._ptrs.ctypes.data_as(ctypes.POINTER(ctypes.c_void_p))

    def slice(self, lo, hi):
        o = self.offsets[lo:hi + 1]
        return Packed(self.units[int(o[0]):int(o[-1])], o - o[0])


def _letters(r, lens, maxlen):
    """rows of letters: r [n, maxlen] draws, lens [n] -> (units uint8 flat, offsets)."""
    n = len(lens)
    chars = (ord("a") + (r % np.uint64(26))).astype(np.uint8)
    mask = np.arange(maxlen)[None, :] < lens[:, None]
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return chars[mask], off


def ascii_keys(n, first=0, seed=KEY_SEED, chunk=1 << 20):
    """n keys (rows first..first+n) of 6..16 lower-case letters."""
    units, lens_all = [], []
    for b in range(0, n, chunk):
        c = min(chunk, n - b)
        r = draws(seed, first + b, c, 17)
        lens = (6 + r[:, 0] % np.uint64(11)).astype(np.int64)
        u, _ = _letters(r[:, 1:], lens, 16)
        units.append(u)
        lens_all.append(lens)
    lens = np.concatenate(lens_all) if lens_all else np.zeros(0, np.int64)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return Packed(np.concatenate(units) if units else np.zeros(0, np.uint8), off)


def key_desc_rows(nrows, first=0, seed=KEY_SEED, chunk=1 << 19):
    """C2: rowSize=2 rows (key, description) and per-word weights. Returns (Packed of 2*nrows words, weights float32)."""
    units, lens_all, weights = [], [], []
    for b in range(0, nrows, chunk):
        c = min(chunk, nrows - b)
        r = draws(seed, first + b, c, 63)
        klen = (6 + r[:, 0] % np.uint64(11)).astype(np.int64)
        kchars = (ord("a") + (r[:, 1:17] % np.uint64(26))).astype(np.uint8)
        # description: word(4..10) + 3 x (' ' + word(3..9))
        w0 = (4 + r[:, 17] % np.uint64(7)).astype(np.int64)
        wl = (3 + r[:, 18:21] % np.uint64(7)).astype(np.int64)
        dchars = np.full((c, 10 + 3 * 10), ord(" "), dtype=np.uint8)
        letters = (ord("a") + (r[:, 21:21 + 40] % np.uint64(26))).astype(np.uint8)
        dchars[:, :] = letters[:, :40]
        pos = w0.copy()
        sep = np.zeros((c, 40), dtype=bool)
        for k in range(3):
            sep[np.arange(c), pos] = True
            pos = pos + 1 + wl[:, k]
        dchars[sep] = ord(" ")
        dlen = pos
        wts = (0.5 + (r[:, 61:63] % np.uint64(1000)).astype(np.float64) / 1000.0).astype(np.float32)
        # interleave key, desc
        lens = np.empty(2 * c, dtype=np.int64)
        lens[0::2] = klen
        lens[1::2] = dlen
        both = np.full((2 * c, 40), 0, dtype=np.uint8)
        both[0::2, :16] = kchars
        both[1::2, :] = dchars
        mask = np.arange(40)[None, :] < lens[:, None]
        units.append(both[mask])
        lens_all.append(lens)
        weights.append(wts.reshape(-1))
    lens = np.concatenate(lens_all)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return Packed(np.concatenate(units), off), np.concatenate(weights)


def wide_keys(n, first=0, seed=KEY_SEED, chunk=1 << 20):
    """C3: n rows of 4..12 code points uniform in U+4E00..U+9FFF (UTF-32 units)."""
    units, lens_all = [], []
    for b in range(0, n, chunk):
        c = min(chunk, n - b)
        r = draws(seed ^ 0x3, first + b, c, 13)
        lens = (4 + r[:, 0] % np.uint64(9)).astype(np.int64)
        cp = (0x4E00 + r[:, 1:] % np.uint64(20992)).astype(np.uint32)
        mask = np.arange(12)[None, :] < lens[:, None]
        units.append(cp[mask])
        lens_all.append(lens)
    lens = np.concatenate(lens_all)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return Packed(np.concatenate(units), off)


def queries_from(keys, nq, seed=QUERY_SEED, stride=1, alphabet=None):
    """nq queries: key (r0 % nKeys) * stride with one random position overwritten by a random alphabet symbol."""
    nkeys = len(keys) // stride
    r = draws(seed, 0, nq, 3)
    pick = (r[:, 0] % np.uint64(nkeys)).astype(np.int64) * stride
    starts = keys.offsets[pick].astype(np.int64)
    lens = (keys.offsets[pick + 1] - keys.offsets[pick]).astype(np.int64)
    off = np.zeros(nq + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    src = np.repeat(starts, lens) + (np.arange(int(off[-1]), dtype=np.int64) - np.repeat(off[:-1].astype(np.int64), lens))
    units = keys.units[src].copy()
    posn = (r[:, 1] % lens.astype(np.uint64)).astype(np.int64)
    if alphabet is None:
        alphabet = (np.arange(0x4E00, 0x4E00 + 20992, dtype=np.uint32) if keys.wide
                    else np.frombuffer(b"abcdefghijklmnopqrstuvwxyz", dtype=np.uint8))
    sym = alphabet[(r[:, 2] % np.uint64(len(alphabet))).astype(np.int64)]
    units[off[:-1].astype(np.int64) + posn] = sym
    return Packed(units, off)


def ascii_keys_at(rows, seed=KEY_SEED):
    """the ASCII keys of arbitrary global row indices (same stream as ascii_keys)."""
    r = draws_at(seed, rows, 17)
    lens = (6 + r[:, 0] % np.uint64(11)).astype(np.int64)
    u, off = _letters(r[:, 1:], lens, 16)
    return Packed(u, off)


def ascii_queries(nq, nkeys, first_query=0, seed=QUERY_SEED):
    """query batch over a (possibly sharded) vocabulary of nkeys global rows, generated without materialising the keys:
    query i = key (r0 % nkeys) with one position overwritten by a random lower-case letter."""
    r = draws(seed, first_query, nq, 3)
    pick = r[:, 0] % np.uint64(nkeys)
    keys = ascii_keys_at(pick)
    lens = np.diff(keys.offsets).astype(np.uint64)
    posn = (r[:, 1] % lens).astype(np.int64)
    units = keys.units.copy()
    units[keys.offsets[:-1].astype(np.int64) + posn] = (ord("a") + (r[:, 2] % np.uint64(26))).astype(np.uint8)
    return Packed(units, keys.offsets)
