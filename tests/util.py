"""Shared helpers for the parity tests."""
import struct

import numpy as np


def f32_bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def tie_groups(results):
    """[(key bytes, score)] -> list of (scoreBits, keyLen, sorted keys). The reference orders by score desc, then key
    length asc (nGramSearch.h:262-269); the order inside such a group is unspecified."""
    groups = []
    for key, score in results:
        k = (f32_bits(score), len(key))
        if groups and groups[-1][0] == k:
            groups[-1][1].append(key)
        else:
            groups.append((k, [key]))
    return [(k, sorted(v)) for k, v in groups]


def assert_same_modulo_ties(ref, ours, limit_cut=False, what=""):
    """ref/ours: [(key bytes, score)]. Equal up to permutation inside (score, length) groups; when `limit` cut the list,
    the last group is compared by cardinality only (which members survive the cut is unspecified in the reference)."""
    assert len(ref) == len(ours), f"{what}: {len(ref)} vs {len(ours)} results\nref={ref[:8]}\nours={ours[:8]}"
    gr, go = tie_groups(ref), tie_groups(ours)
    assert [g[0] for g in gr] == [g[0] for g in go], f"{what}: group structure differs\nref={gr[:6]}\nours={go[:6]}"
    last = len(gr) - 1
    for i, (a, b) in enumerate(zip(gr, go)):
        if limit_cut and i == last:
            assert len(a[1]) == len(b[1]), what
        else:
            assert a[1] == b[1], f"{what}: group {a[0]} differs: {a[1][:5]} vs {b[1][:5]}"


def spec_keys(tables):
    """ext id -> key string (bytes for narrow, str for wide) from SpecIndex.tables()."""
    off, ch = tables["keyOffExt"], tables["keyCharsExt"]
    wide = ch.dtype == np.uint32
    out = []
    for i in range(tables["K"]):
        u = ch[int(off[i]):int(off[i + 1])]
        out.append("".join(map(chr, u.tolist())) if wide else u.tobytes())
    return out


def random_corpus(rng, n_rows, row_size, weighted=True, alphabet=b"abcABC12 .#-", min_len=0, max_len=12, null_rate=0.03):
    """small-alphabet words: many duplicate keys/terms, blanks, invalid chars ('#', '-'), empty cells, NULLs."""
    words, weights = [], []
    for _ in range(n_rows * row_size):
        if rng.random() < null_rate:
            words.append(None)
        else:
            ln = int(rng.integers(min_len, max_len + 1))
            words.append(bytes(rng.choice(list(alphabet), size=ln).astype(np.uint8)))
        weights.append(float(rng.choice([0.0, 0.5, 1.0, 1.0, 1.25, 2.0])) if weighted else 1.0)
    return words, (np.array(weights, np.float32) if weighted else None)
