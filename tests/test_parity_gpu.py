"""Parity tests proper: the CUDA engine, called through its C ABI, against the CPU oracle (oracle/spec) on the same
seeded inputs — bit-exact tables (keys, terms, edges, gram table, postings) and bit-exact ranked lists + scores — and
against the reference-generated golden fixture. All tests need a B200 (-m gpu)."""
import json
import os

import numpy as np
import pytest

from util import assert_same_modulo_ties, random_corpus, spec_keys

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")


@pytest.fixture(scope="module")
def ssb():
    import stringsearchlib_b200 as m
    return m


def check_tables(ix, S):
    t, o = ix.tables(), S.tables()
    i = t["info"]
    assert bool(i["indexed"]) == o["indexed"]
    for k in ("K", "T", "E", "G", "P", "nShort"):
        assert i[k] == o[k], (k, i[k], o[k])
    if i["T"] == 0:
        return t, o
    # keys: by kid on the device, by ext id in the oracle
    ext_of_kid = np.argsort(o["kidOfExt"], kind="stable")
    lens_o = np.diff(o["keyOffExt"]).astype(np.int64)[ext_of_kid]
    lens_d = np.diff(t["keyOff"]).astype(np.int64) - 1
    assert np.array_equal(lens_o, lens_d)
    chars_o = np.concatenate([o["keyCharsExt"][int(o["keyOffExt"][e]):int(o["keyOffExt"][e + 1])] for e in ext_of_kid]) \
        if i["K"] < 200000 else None
    if chars_o is not None:
        mask = np.ones(len(t["keyChars"]), bool)
        mask[t["keyOff"][1:].astype(np.int64) - 1] = False          # terminators
        assert np.all(t["keyChars"][~mask] == 0)
        assert np.array_equal(t["keyChars"][mask], chars_o)
    assert np.array_equal(t["keyRow"], o["firstRowExt"][ext_of_kid])
    # terms, edges
    assert np.array_equal(t["termOff"].astype(np.uint64), o["termOff"])
    assert np.array_equal(t["termChars"], o["termChars"])
    assert np.array_equal(t["termEdgeOff"], o["termEdgeOff"])
    assert np.array_equal(t["edgeKid"], o["edgeKid"])
    assert np.array_equal(t["edgeW"].view(np.uint32), o["edgeW"].view(np.uint32))
    # gram table + postings
    assert np.array_equal(ix.gram_units(t), o["gramUnits"])
    assert np.array_equal(t["gramOff"].astype(np.uint64), o["gramOff"])
    assert np.array_equal(t["postings"], o["postings"])
    return t, o


def check_search(ix, S, queries, threshold, limit, o=None):
    """every batch is answered twice — k_sieve_search first (path 1, the default) and k_fast_search alone (path 0) — and
    both must equal the oracle bit for bit"""
    import stringsearchlib_b200 as m
    o = o or S.tables()
    counts, ext, sc = S.search_batch(queries, threshold, limit)
    nq = len(queries)
    kid_of_ext = o["kidOfExt"]
    r = None
    try:
        for path in (1, 0):
            m.set_search_path(path)
            r = ix.search_batch(queries, threshold, limit)
            assert len(r) == nq
            got_counts = np.diff(r.offsets)
            assert np.array_equal(got_counts, counts), (path, np.nonzero(got_counts != counts)[0][:10])
            for q in range(nq):
                n = int(counts[q])
                kids, scores, _, _ = r.query(q)
                assert np.array_equal(scores.view(np.uint32), sc[q, :n].view(np.uint32)), (path, q, scores[:8], sc[q, :8])
                assert np.array_equal(kids, kid_of_ext[ext[q, :n]]), (path, q, kids[:8], kid_of_ext[ext[q, :8]])
    finally:
        m.set_search_path(-1)
    return r


def test_searchtest_fixture_handle_abi(ssb):
    """SearchTest/test.cpp:7-18 through the as-shipped exports (rowSize=1)."""
    from conftest import SEVEN
    H = ssb.HandleLibrary()
    h = H.indexN(SEVEN, 1)
    assert h >= 1
    assert H.lib.getSize(h) == 7
    assert H.lib.getLibSize(h) == 16
    res = H.score(h, "LWMS", 0.5, 2 ** 31 - 1)
    assert [(k, s) for k, s in res] == [(b"LWMS", 100.0), (b"LWM", 0.75), (b"LWMA", 0.75), (b"LWYY", 0.5)]
    assert H.search(h, "LWMS", 0.5, 0) == [b"LWMS", b"LWM", b"LWMA", b"LWYY"]
    assert H.score(h, "   ", 0.0, 0) == []
    assert len(H.score(h, "*", 0.0, 0)) == 7
    assert H.score(9999, "LWMS", 0.0, 0) == []           # unknown handle => 0 (dllmain.cpp:69)
    h2 = H.indexN(SEVEN, 7)
    assert h2 == h + 1 and H.lib.getSize(h2) == 7 and H.lib.getLibSize(h2) == 16
    assert {k for k, _ in H.score(h2, "LWMA", 0.0, 0)} == {b"LWMS"}
    H.lib.dispose(h)
    assert H.lib.getSize(h) == 0
    assert H.indexN(SEVEN, 1) == h                        # smallest free handle is reused (dllmain.cpp:41-46)
    h1 = H.indexN([b"LWMS"], 1)                           # size < 2 => handle, nothing indexed (nGramSearch.hpp:122)
    assert h1 >= 1 and H.lib.getSize(h1) == 0 and H.score(h1, "LWMS", 0.0, 0) == []


def test_golden_fixture(ssb, oracle_built):
    with open(GOLDEN) as f:
        cases = json.load(f)["cases"]
    for case in cases:
        words = [None if w is None else bytes.fromhex(w) for w in case["words"]]
        weights = None if case["weights"] is None else np.array(case["weights"], np.float32)
        S = oracle_built.SpecIndex(words, case["rowSize"], weights)
        with ssb.Index(words, case["rowSize"], weights) as ix:
            assert ix.size() == case["getSize"] and ix.lib_size() == case["getLibSize"]
            _, o = check_tables(ix, S)
            keys = spec_keys(o) if o["K"] else []
            for q in case["queries"]:
                query = bytes.fromhex(q["q"])
                r = ix.search_batch([query], q["threshold"], q["limit"], strings=True)
                ours = list(zip(r.strings, r.scores.tolist()))
                ref = [(bytes.fromhex(k), np.uint32(b).view(np.float32).item()) for k, b in q["results"]]
                cut = q["limit"] > 0 and len(ref) == q["limit"]
                assert_same_modulo_ties(ref, ours, cut, what=f"{case['name']} q={query!r}")
                ext, sc = S.search(query, q["threshold"], q["limit"])
                assert [k for k, _ in ours] == [keys[e] for e in ext]
                assert np.array_equal(r.scores.view(np.uint32), sc.view(np.uint32))


@pytest.mark.parametrize("n,nq,th,limit", [(10000, 512, 0.0, 100), (10000, 128, 0.3, 100), (200000, 256, 0.0, 100)])
def test_c1_ascii_rowsize1(ssb, oracle_built, n, nq, th, limit):
    """C1 / C4 shape: ASCII keys of 6-16 letters, rowSize=1, gSize=3 (fast path: identity + uniform weights)."""
    from stringsearchlib_b200 import synth
    keys = synth.ascii_keys(n)
    S = oracle_built.SpecIndex(keys, 1)
    with ssb.Index(keys, 1) as ix:
        info = ix.info()
        assert info["identity"] and info["uniformW"]
        _, o = check_tables(ix, S)
        q = synth.queries_from(keys, nq)
        check_search(ix, S, q, th, limit, o)
        # unmodified keys are exact matches: promoted to 100 (nGramSearch.hpp:333-335) only if the key is upper-case;
        # these keys are lower-case, so the top score is 1.0
        r = ix.search_batch(keys.slice(0, 8), 0.0, 3, strings=True)
        for i in range(8):
            k, s, _, strs = r.query(i)
            assert strs[0] == keys.get(i) and s[0] == 1.0
        _, _, st = ix.search_resident(8, 0.0, 3)
        assert st["nFast"] == 8 and st["nDense"] == 0


def test_sieve_16_byte_layout_without_threshold(ssb, oracle_built, monkeypatch):
    """indexes of more than 32M terms have only the 16-byte sieve groups (index.h): their path through k_sieve_search at threshold 0
    (dynamic pass, postings read for every candidate block) is forced here on a small index with SSB_SIEVE_INLINE=0."""
    from stringsearchlib_b200 import synth
    monkeypatch.setenv("SSB_SIEVE_INLINE", "0")
    keys = synth.ascii_keys(200000)
    S = oracle_built.SpecIndex(keys, 1)
    with ssb.Index(keys, 1) as ix:
        _, o = check_tables(ix, S)
        q = synth.queries_from(keys, 256)
        check_search(ix, S, q, 0.0, 100, o)
        check_search(ix, S, q, 0.0, 10, o)
        check_search(ix, S, q, 0.3, 500, o)
    monkeypatch.delenv("SSB_SIEVE_INLINE")
    with ssb.Index(keys, 1) as ix:                      # and the default (both layouts) on the same inputs
        _, o = check_tables(ix, S)
        check_search(ix, S, q, 0.0, 10, o)
        check_search(ix, S, q, 0.3, 500, o)


def test_c2_rowsize2_weights(ssb, oracle_built):
    """C2 shape (scaled): rows (key, description), per-word weights, threshold 0.3 and 0."""
    from stringsearchlib_b200 import synth
    words, w = synth.key_desc_rows(40000)
    S = oracle_built.SpecIndex(words, 2, w)
    with ssb.Index(words, 2, w) as ix:
        _, o = check_tables(ix, S)
        q = synth.queries_from(words, 300, stride=2)
        check_search(ix, S, q, 0.3, 100, o)
        check_search(ix, S, q, 0.0, 10, o)
        # description words as queries (several terms point to one key)
        dq = [words.get(2 * i + 1).split(b" ")[0] for i in range(100)]
        check_search(ix, S, dq, 0.0, 20, o)


def test_c3_wide_gsize2(ssb, oracle_built):
    """C3 shape (scaled): indexW-style UTF-32 rows in U+4E00..U+9FFF, gSize=2, limit 50."""
    from stringsearchlib_b200 import synth
    keys = synth.wide_keys(30000)
    S = oracle_built.SpecIndex(keys, 1, None, 2)
    with ssb.Index(keys, 1, None, 2) as ix:
        assert ix.info()["unit"] == 4
        _, o = check_tables(ix, S)
        q = synth.queries_from(keys, 200)
        check_search(ix, S, q, 0.0, 50, o)
        check_search(ix, S, q, 0.5, 50, o)


def test_large_limits_threshold_zero(ssb, oracle_built):
    """limits whose first pass overflows the candidate buffer (3 * limit > capacity: 342-512, 683-1024, 1366-2048) at threshold
    0 on an identity index — the pass is rolled back and redone, and every counter that feeds tau has to be rolled back
    with it (k_fast_search) / the query has to survive the hand-over (k_sieve_search)."""
    from stringsearchlib_b200 import synth
    keys = synth.ascii_keys(300000)
    S = oracle_built.SpecIndex(keys, 1)
    with ssb.Index(keys, 1) as ix:
        o = S.tables()
        q = synth.queries_from(keys, 48)
        for limit in (400, 1000, 2000):
            check_search(ix, S, q, 0.0, limit, o)


@pytest.mark.parametrize("g", [2, 3, 4, 5])
def test_gsize_sweep(ssb, oracle_built, g):
    """C5 shape (scaled): gSize 2-5, threshold 0.5, limit 1000."""
    from stringsearchlib_b200 import synth
    keys = synth.ascii_keys(20000)
    S = oracle_built.SpecIndex(keys, 1, None, g)
    with ssb.Index(keys, 1, None, g) as ix:
        _, o = check_tables(ix, S)
        q = synth.queries_from(keys, 128)
        check_search(ix, S, q, 0.5, 1000, o)
        check_search(ix, S, q, 0.0, 1000, o)


@pytest.mark.parametrize("seed,rows,rs", [(1, 3000, 1), (2, 2000, 3), (3, 1500, 2)])
def test_messy_corpus_all_paths(ssb, oracle_built, seed, rows, rs):
    """duplicates, blank/NULL cells, invalid characters, zero weights, short terms (edit-distance path,
    nGramSearch.hpp:182-270), wildcard, limit 0 (unlimited) -> exercises the dense path and many-to-many edges."""
    rng = np.random.default_rng(seed)
    words, w = random_corpus(rng, rows, rs, weighted=(seed != 1), max_len=14)
    S = oracle_built.SpecIndex(words, rs, w)
    with ssb.Index(words, rs, w) as ix:
        _, o = check_tables(ix, S)
        pool = [x for x in words if x]
        qs = []
        for _ in range(150):
            b = bytearray(pool[int(rng.integers(len(pool)))])
            if b and rng.random() < 0.6:
                b[int(rng.integers(len(b)))] = int(rng.choice(list(b"abcABC12 .#")))
            qs.append(bytes(b))
        qs += [b"", b"*", b"   ", b"a", b"ab", b"abc", b"#", b"abcabcabcabcabcabcabcabcabcabcabcabcabcabc"]
        for th, lim in ((0.0, 0), (0.3, 10), (0.0, 100), (0.75, 0)):
            if seed != 1 and lim == 0:
                qq = [x for x in qs if x not in (b"", b"*")]    # wildcard over multi-edge keys is order-dependent in the reference
            else:
                qq = qs
            check_search(ix, S, qq, th, lim, o)


def test_long_terms_only_fast_generic(ssb, oracle_built):
    """rowSize=3 rows of long words with weights: fast path in generic mode (edges, weights, per-key max-merge)."""
    from stringsearchlib_b200 import synth
    keys = synth.ascii_keys(30000)
    rng = np.random.default_rng(7)
    w = rng.choice([0.5, 1.0, 1.5, 2.0], size=30000).astype(np.float32)
    S = oracle_built.SpecIndex(keys, 3, w)
    with ssb.Index(keys, 3, w) as ix:
        info = ix.info()
        assert not info["identity"] and info["nShort"] == 0
        _, o = check_tables(ix, S)
        q = synth.queries_from(keys, 300)
        r = check_search(ix, S, q, 0.0, 50, o)
        _, _, st = ix.search_resident(300, 0.0, 50)
        assert st["nFast"] == 300
        check_search(ix, S, q, 0.4, 50, o)


def test_upper_case_exact_match_promotion(ssb, oracle_built):
    """exact-match promotion to 100 compares the escaped, NOT upper-cased key with the upper-cased query (hpp:330-335)."""
    words = [b"ALPHABETIC", b"alphabetic", b"ALPHABETIC!", b" ALPHABETIC ", b"ALPHABETICAL", b"BETA GAMMA", b"beta gamma"]
    S = oracle_built.SpecIndex(words, 1)
    with ssb.Index(words, 1) as ix:
        _, o = check_tables(ix, S)
        r = check_search(ix, S, [b"alphabetic", b"ALPHABETIC", b"beta gamma", b"alphabetic#"], 0.0, 0, o)
        res = ix.search("alphabetic")
        assert res[0][1] == 100.0 and res[0][0] == b"ALPHABETIC"


def test_set_valid_char(ssb, oracle_built):
    """setValidChar (dllmain.cpp:142-151): affects later queries and the exact-match test; the index is not rebuilt."""
    from stringsearchlib_b200 import synth
    keys = synth.ascii_keys(5000)
    S = oracle_built.SpecIndex(keys, 1)
    with ssb.Index(keys, 1) as ix:
        o = S.tables()
        q = [keys.get(i) for i in range(50)]
        valid = b"abcdefghijklm ABCDEFGHIJKLM"
        ix.set_valid_char(valid)
        S.set_valid_char(valid)
        check_search(ix, S, q, 0.0, 20, o)


def test_guid_abi(ssb):
    """README surface (Readme.md:31-226): index/indexW/index2D/index2DW, search/searchW, release[W], dispose[W], sizes."""
    from conftest import SEVEN
    G = ssb.GuidLibrary()
    G.index(b"lib-a", SEVEN, 1, None, 3)
    assert G.lib.getSize(b"lib-a") == 7 and G.lib.getLibSize(b"lib-a") == 16
    assert G.search(b"lib-a", "LWMS", 0.5, 100) == [b"LWMS", b"LWM", b"LWMA", b"LWYY"]
    assert G.search(b"nope", "LWMS", 0.5, 100) == []
    G.index2D(b"lib-2d", [[b"apple pie", b"dessert"], [b"beef stew", b"dinner"]], [[1.0, 0.5], [1.0, 0.5]], 3)
    assert G.search(b"lib-2d", "dessert", 0.0, 10) == [b"apple pie"]
    G.indexW(b"lib-w", ["中文字符串", "中文检索库", "latin text"], 1, None, 2)
    assert G.lib.getSizeW(b"lib-w") == 3
    assert G.searchW(b"lib-w", "中文字符", 0.0, 10)[0] == "中文字符串"
    G.index2D(b"lib-2dw", [["中文字符串", "描述文本"]] * 2, None, 2, wide=True)
    assert G.searchW(b"lib-2dw", "描述文本", 0.0, 10) == ["中文字符串"]
    G.index(b"lib-a", [b"replaced key one", b"replaced key two"], 1, None, 3)      # re-index replaces
    assert G.lib.getSize(b"lib-a") == 2
    G.lib.dispose(b"lib-a")
    G.lib.disposeW(b"lib-w")
    assert G.lib.getSize(b"lib-a") == 0 and G.lib.getSizeW(b"lib-w") == 0
    G.lib.dispose(b"never-existed")


def test_key_range_shards_merge(ssb, oracle_built):
    """multi-GPU path on one GPU: N key-range shards -> export records -> merge == unsharded oracle."""
    import torch
    from stringsearchlib_b200 import synth
    from stringsearchlib_b200.api import RECORD_DTYPE, merge_records
    n, nq, limit = 60000, 200, 20
    keys = synth.ascii_keys(n)
    # duplicate some keys across shards: the same key string must collapse in the merge
    q = synth.queries_from(keys, nq)
    S = oracle_built.SpecIndex(keys, 1)
    o = S.tables()
    counts, ext, sc = S.search_batch(q, 0.0, limit)
    for shards in (2, 3):
        bounds = np.linspace(0, n, shards + 1).astype(int)
        gathered = torch.empty((shards, nq, limit, RECORD_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
        idx = []
        for s in range(shards):
            ix = ssb.Index(keys.slice(bounds[s], bounds[s + 1]), 1, row_base=int(bounds[s]))
            ix.search_batch(q, 0.0, limit)
            ix.export_records(nq, limit, gathered[s].data_ptr())
            idx.append(ix)
        torch.cuda.synchronize()
        merged = torch.empty((nq, limit, RECORD_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
        mcount = torch.empty(nq, dtype=torch.int32, device="cuda")
        merge_records(gathered.data_ptr(), shards, nq, limit, merged.data_ptr(), mcount.data_ptr())
        torch.cuda.synchronize()
        rec = merged.cpu().numpy().view(RECORD_DTYPE).reshape(nq, limit)
        mc = mcount.cpu().numpy()
        assert np.array_equal(mc, counts)
        for i in range(nq):
            c = int(counts[i])
            assert np.array_equal(rec["scoreBits"][i, :c], sc[i, :c].view(np.uint32))
            assert np.array_equal(rec["row"][i, :c], o["firstRowExt"][ext[i, :c]])
        for ix in idx:
            ix.close()


def test_key_range_shards_p2p_exchange(ssb, oracle_built):
    """the fused exchange (export kernel stores into the merging rank's buffer, merge kernel stores into every rank's result
    buffer) with N emulated ranks on one GPU: all buffers are local, the indexing is the multi-GPU one."""
    from stringsearchlib_b200 import synth
    from stringsearchlib_b200 import api
    n, nq, limit, N = 40000, 128, 20, 4
    keys = synth.ascii_keys(n)
    q = synth.queries_from(keys, nq)
    S = oracle_built.SpecIndex(keys, 1)
    o = S.tables()
    counts, ext, sc = S.search_batch(q, 0.0, limit)
    rb = api.RECORD_DTYPE.itemsize
    part = nq // N
    recv = [api.device_alloc(N * part * limit * rb) for _ in range(N)]
    merged = [api.device_alloc(nq * limit * rb) for _ in range(N)]
    cnts = [api.device_alloc(nq * 4) for _ in range(N)]
    bounds = np.linspace(0, n, N + 1).astype(int)
    shards = []
    for r in range(N):
        ix = ssb.Index(keys.slice(bounds[r], bounds[r + 1]), 1, row_base=int(bounds[r]))
        ix.search_batch(q, 0.0, limit)
        ix.export_records_p2p(nq, limit, recv, r)
        shards.append(ix)
    for r in range(N):
        api.merge_records_p2p(recv[r], N, part, limit, merged, cnts, r)
    import torch
    torch.cuda.synchronize()
    for r in range(N):
        rec = np.zeros((nq, limit), api.RECORD_DTYPE)
        mc = np.zeros(nq, np.uint32)
        api.memcpy_d2h(rec, merged[r])
        api.memcpy_d2h(mc, cnts[r])
        assert np.array_equal(mc, counts)
        for i in range(nq):
            c = int(counts[i])
            assert np.array_equal(rec["scoreBits"][i, :c], sc[i, :c].view(np.uint32))
            assert np.array_equal(rec["row"][i, :c], o["firstRowExt"][ext[i, :c]])
    for ix in shards:
        ix.close()
    for p in recv + merged + cnts:
        api.device_free(p)


def test_key_range_shards_p2p_signal_wait(ssb, oracle_built):
    """the exchange without any collective: the export kernel signals an epoch into every rank's flags, the merge kernel
    waits for the flags of all shards; two exchanges on alternating receive buffers, every rank keeps only the slice it
    merged. N emulated ranks on one GPU (all buffers local, the indexing is the multi-GPU one)."""
    import torch
    from stringsearchlib_b200 import synth
    from stringsearchlib_b200 import api
    n, nq, limit, N = 40000, 128, 20, 4
    keys = synth.ascii_keys(n)
    S = oracle_built.SpecIndex(keys, 1)
    o = S.tables()
    rb = api.RECORD_DTYPE.itemsize
    part = nq // N
    recv = [[api.device_alloc(N * part * limit * rb) for _ in range(N)] for _ in range(2)]
    merged = [api.device_alloc(nq * limit * rb) for _ in range(N)]
    cnts = [api.device_alloc(nq * 4) for _ in range(N)]
    flags = [api.device_alloc(64) for _ in range(N)]
    for f in flags:
        api.device_memset(f, 0, 64)
    bounds = np.linspace(0, n, N + 1).astype(int)
    shards = [ssb.Index(keys.slice(bounds[r], bounds[r + 1]), 1, row_base=int(bounds[r])) for r in range(N)]
    for epoch in (1, 2):
        q = synth.queries_from(keys, nq, stride=epoch)
        counts, ext, sc = S.search_batch(q, 0.0, limit)
        for r in range(N):
            shards[r].search_batch(q, 0.0, limit)
            shards[r].export_records_p2p_signal(nq, limit, recv[epoch & 1], r, flags, epoch)
        for r in range(N):
            api.merge_records_p2p_wait(recv[epoch & 1][r], N, part, limit, [merged[r]], [cnts[r]], r, flags[r], epoch)
        torch.cuda.synchronize()
        for r in range(N):
            fl = np.zeros(16, np.uint32)
            api.memcpy_d2h(fl, flags[r])
            assert np.all(fl[:N] == epoch)
            rec = np.zeros((nq, limit), api.RECORD_DTYPE)
            mc = np.zeros(nq, np.uint32)
            api.memcpy_d2h(rec, merged[r])
            api.memcpy_d2h(mc, cnts[r])
            sl = slice(r * part, (r + 1) * part)
            assert np.array_equal(mc[sl], counts[sl])
            for i in range(r * part, (r + 1) * part):
                c = int(counts[i])
                assert np.array_equal(rec["scoreBits"][i, :c], sc[i, :c].view(np.uint32))
                assert np.array_equal(rec["row"][i, :c], o["firstRowExt"][ext[i, :c]])
    for ix in shards:
        ix.close()
    for p in recv[0] + recv[1] + merged + cnts + flags:
        api.device_free(p)


def test_global_sizes_over_shards(ssb, oracle_built):
    """getSize / getLibSize of a sharded index (SURVEY.md §8e "Global counters", nGramSearch.hpp:488-499): the distinct terms
    and grams over all key-range shards — digests gathered, counted per residue class, summed — equal the unsharded index's
    and the oracle's numbers. Corpora with many terms and grams shared between shards; narrow and wide; gSize 2, 3 and 5."""
    import torch
    from stringsearchlib_b200 import synth
    from stringsearchlib_b200.api import count_distinct_digests
    rng = np.random.default_rng(11)
    messy, w = random_corpus(rng, 4000, 2, weighted=True, max_len=14)
    wide = synth.wide_keys(20000)
    wide_dup = [wide.get(i) for i in list(range(12000)) + list(range(3000, 9000))]            # rows 12000.. repeat rows 3000..
    ascii_dup = [synth.ascii_keys(9000).get(i) for i in list(range(9000)) + list(range(0, 9000, 3))]
    cases = [(messy, 2, w, 3, False), (messy, 2, w, 2, False), (wide_dup, 1, None, 2, True), (ascii_dup, 1, None, 3, False),
             (ascii_dup, 1, None, 5, False)]
    for words, rs, weights, g, is_wide in cases:
        S = oracle_built.SpecIndex(words, rs, weights, g, is_wide)
        rows = len(words) // rs
        with ssb.Index(words, rs, weights, g, wide=is_wide) as whole:
            assert (whole.size(), whole.lib_size()) == (S.size(), S.lib_size())
        for shards in (2, 3):
            bounds = np.linspace(0, rows, shards + 1).astype(int)
            idx = []
            for sh in range(shards):
                lo, hi = int(bounds[sh]) * rs, int(bounds[sh + 1]) * rs
                idx.append(ssb.Index(words[lo:hi], rs, None if weights is None else weights[lo:hi], g, wide=is_wide, row_base=int(bounds[sh])))
            got = []
            for which in (0, 1):
                counts = [ix.digest_count(which) for ix in idx]
                assert counts == [ix.size() if which == 0 else ix.lib_size() for ix in idx]
                buf = torch.zeros((max(sum(counts), 1), 16), dtype=torch.uint8, device="cuda")
                at = 0
                for ix, c in zip(idx, counts):
                    if c:
                        ix.export_digests(which, buf[at:].data_ptr())
                    at += c
                torch.cuda.synchronize()
                one = count_distinct_digests(buf.data_ptr(), at)
                parts = [count_distinct_digests(buf.data_ptr(), at, p, 3) for p in range(3)]
                assert sum(parts) == one and (at < 100 or all(parts))
                got.append(one)
            for ix in idx:
                ix.close()
            assert tuple(got) == (S.size(), S.lib_size()), (rs, g, is_wide, shards, got, sum(counts))
            if shards == 2:
                assert sum(counts) > got[1]          # the shards do share grams: a sum would be wrong


def test_merge_collapses_duplicates_and_large_limit(ssb, oracle_built):
    """the same key strings in several shards collapse to one entry (best score); 8 shards x limit 1000 exercises the
    hierarchical merge (more records than one CTA's shared memory holds)."""
    import torch
    from stringsearchlib_b200 import synth
    from stringsearchlib_b200.api import RECORD_DTYPE, merge_records
    n, nq, limit, shards = 24000, 32, 1000, 8
    keys = synth.ascii_keys(n)
    words = [keys.get(i) for i in range(n)]
    # every shard also holds a copy of the first 50 keys: duplicates across shards
    per = n // shards
    q = synth.queries_from(keys, nq)
    S = oracle_built.SpecIndex(words, 1)
    o = S.tables()
    counts, ext, sc = S.search_batch(q, 0.0, limit)
    gathered = torch.empty((shards, nq, limit, RECORD_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    idx = []
    for s in range(shards):
        mine = words[s * per:(s + 1) * per] + (words[:50] if s else [])
        ix = ssb.Index(mine, 1, row_base=s * per if s == 0 else s * per)
        # rows of the duplicated tail sit after the shard's own rows, so their global row ids are larger than the originals'
        ix.search_batch(q, 0.0, limit)
        ix.export_records(nq, limit, gathered[s].data_ptr())
        idx.append(ix)
    merged = torch.empty((nq, limit, RECORD_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    mcount = torch.empty(nq, dtype=torch.int32, device="cuda")
    merge_records(gathered.data_ptr(), shards, nq, limit, merged.data_ptr(), mcount.data_ptr())
    torch.cuda.synchronize()
    rec = merged.cpu().numpy().view(RECORD_DTYPE).reshape(nq, limit)
    mc = mcount.cpu().numpy()
    assert np.array_equal(mc, counts)
    for i in range(nq):
        c = int(counts[i])
        assert np.array_equal(rec["scoreBits"][i, :c], sc[i, :c].view(np.uint32))
        assert np.array_equal(rec["keyLen"][i, :c], np.diff(o["keyOffExt"]).astype(np.uint32)[ext[i, :c]])
    for ix in idx:
        ix.close()


def test_properties_full_size(ssb):
    """size-independent properties at 1M keys (the oracle is not run): sorted scores, deterministic repeat,
    unmodified key ranks first, resident rerun identical, no dense fallback on this distribution."""
    from stringsearchlib_b200 import synth
    n, nq = 1000000, 4096
    keys = synth.ascii_keys(n)
    with ssb.Index(keys, 1) as ix:
        info = ix.info()
        assert info["K"] == info["T"] <= n and info["G"] == 26 ** 3 and info["identity"]
        q = synth.queries_from(keys, nq)
        r1 = ix.search_batch(q, 0.0, 100)
        r2 = ix.search_batch(q, 0.0, 100)
        assert np.array_equal(r1.offsets, r2.offsets) and np.array_equal(r1.kids, r2.kids)
        assert np.array_equal(r1.scores.view(np.uint32), r2.scores.view(np.uint32))
        assert np.all(np.diff(r1.offsets) == 100)
        sc = r1.scores.reshape(nq, 100)
        assert np.all(sc[:, :-1] >= sc[:, 1:])
        kid = r1.kids.reshape(nq, 100)
        same = sc[:, :-1] == sc[:, 1:]
        assert np.all(kid[:, :-1][same] < kid[:, 1:][same])          # ties: kid ascending
        ms, msk, st = ix.search_resident(nq, 0.0, 100)
        # the default path chooses per batch: ~4 600 postings per query at this size -> the register merge (k_fast_search)
        assert st["nFast"] == nq and st["nDense"] == 0 and st["nOverflow"] == 0, st
        assert st["nSieve"] == 0 and st["nSieveHandover"] == nq, st
        # k_sieve_search for every query it can take (path 1) and k_fast_search alone (path 0) must answer the batch identically
        for path in (1, 0):
            try:
                ssb.set_search_path(path)
                r3 = ix.search_batch(q, 0.0, 100)
                _, _, st3 = ix.search_resident(nq, 0.0, 100)
            finally:
                ssb.set_search_path(-1)
            assert np.array_equal(r1.offsets, r3.offsets) and np.array_equal(r1.kids, r3.kids), path
            assert np.array_equal(r1.scores.view(np.uint32), r3.scores.view(np.uint32)), path
            assert st3["nFast"] == nq and st3["nDense"] == 0, (path, st3)
            if path == 1:
                assert st3["nSieve"] + st3["nSieveHandover"] == nq and st3["nSieve"] > nq // 2, st3
            else:
                assert st3["nSieve"] == 0, st3
        exact = ix.search_batch(keys.slice(1000, 1064), 0.0, 5, strings=True)
        for i in range(64):
            assert exact.query(i)[3][0] == keys.get(1000 + i)


def test_c4_full_size_both_kernels_agree(ssb):
    """BASELINE.json's headline configuration at full size (10M keys, batch 4096, top-100; the oracle is not run): the
    default choice there is k_sieve_search, and the register merge of k_fast_search — a different algorithm over the same
    postings — must give bit-identical lists; plus the size-independent properties."""
    from stringsearchlib_b200 import synth
    n, nq = 10_000_000, 4096
    keys = synth.ascii_keys(n)
    with ssb.Index(keys, 1) as ix:
        info = ix.info()
        assert info["identity"] and info["uniformW"] and info["G"] == 26 ** 3
        q = synth.ascii_queries(nq, n)
        r1 = ix.search_batch(q, 0.0, 100)
        _, _, st = ix.search_resident(nq, 0.0, 100)
        assert st["nFast"] == nq and st["nDense"] == 0 and st["nSieve"] + st["nSieveHandover"] == nq and st["nSieve"] > nq // 2, st
        assert np.all(np.diff(r1.offsets) == 100)
        sc = r1.scores.reshape(nq, 100)
        kid = r1.kids.reshape(nq, 100)
        assert np.all(sc[:, :-1] >= sc[:, 1:])
        same = sc[:, :-1] == sc[:, 1:]
        assert np.all(kid[:, :-1][same] < kid[:, 1:][same])          # ties: kid ascending
        try:
            ssb.set_search_path(0)
            r0 = ix.search_batch(q, 0.0, 100)
            _, _, st0 = ix.search_resident(nq, 0.0, 100)
        finally:
            ssb.set_search_path(-1)
        assert st0["nSieve"] == 0 and st0["nFast"] == nq, st0
        assert np.array_equal(r1.offsets, r0.offsets) and np.array_equal(r1.kids, r0.kids)
        assert np.array_equal(r1.scores.view(np.uint32), r0.scores.view(np.uint32))
        # under a threshold (C5-like: 0.5, top-1000) the two kernels must agree as well
        ra = ix.search_batch(q.slice(0, 512), 0.5, 1000)
        try:
            ssb.set_search_path(0)
            rb = ix.search_batch(q.slice(0, 512), 0.5, 1000)
        finally:
            ssb.set_search_path(-1)
        assert np.array_equal(ra.offsets, rb.offsets) and np.array_equal(ra.kids, rb.kids)
        assert np.array_equal(ra.scores.view(np.uint32), rb.scores.view(np.uint32))


def test_index_save_load_round_trip(ssb, oracle_built, tmp_path):
    """a saved and re-loaded index has bit-identical tables and answers bit-identically (narrow with
    descriptions + weights + short terms, and wide gSize=2)."""
    from stringsearchlib_b200 import synth
    words, w = synth.key_desc_rows(3000)
    words = [words.get(i) for i in range(len(words))] + [b"LWMS", b"lw", b"LWM", b"x", None, b"  ", b"ab#c d", b"q"]
    w = np.concatenate([w, np.linspace(0.5, 1.5, 8).astype(np.float32)])
    S = oracle_built.SpecIndex(words, 2, w)
    qs = [words[i] for i in range(0, 200, 2) if words[i]] + [b"LWMS", b"*", b"", b"l", b"lw", b"abc d"]
    path = str(tmp_path / "narrow.ssbidx")
    with ssb.Index(words, 2, w) as ix:
        ix.set_valid_char(b".%$ @0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ#")
        t0 = ix.tables()
        r0 = ix.search_batch(qs, 0.0, 50, strings=True)
        ix.save(path)
    with ssb.Index.load(path) as ix2:
        assert ix2.size() == S.size() and ix2.lib_size() == S.lib_size()
        t1 = ix2.tables()
        for k in t0:
            if k == "info":
                assert {a: b for a, b in t0[k].items() if a not in ("windows", "hashRetries")} == \
                       {a: b for a, b in t1[k].items() if a not in ("windows", "hashRetries")}
            else:
                assert np.array_equal(t0[k], t1[k]), k
        r1 = ix2.search_batch(qs, 0.0, 50, strings=True)
        assert np.array_equal(r0.offsets, r1.offsets) and np.array_equal(r0.kids, r1.kids)
        assert np.array_equal(r0.scores.view(np.uint32), r1.scores.view(np.uint32))
        assert r0.strings == r1.strings
    wk = synth.wide_keys(4000)
    path2 = str(tmp_path / "wide.ssbidx")
    with ssb.Index(wk, 1, None, 2) as ix:
        q = synth.queries_from(wk, 64)
        r0 = ix.search_batch(q, 0.0, 20, strings=True)
        ix.save(path2)
    with ssb.Index.load(path2) as ix2:
        assert ix2.wide
        r1 = ix2.search_batch(q, 0.0, 20, strings=True)
        assert np.array_equal(r0.kids, r1.kids) and np.array_equal(r0.scores.view(np.uint32), r1.scores.view(np.uint32))
        assert r0.strings == r1.strings
    with pytest.raises(RuntimeError):
        ssb.Index.load(str(tmp_path / "missing.ssbidx"))


def test_concurrent_single_queries_coalesce(ssb, oracle_built):
    """many threads calling the single-query entry point at once (the reference's intended usage,
    dllmain.cpp:63) get exactly the single-threaded answers, and their calls are answered in shared batches — through
    ssb_search_one directly and through score() of the as-shipped handle ABI."""
    import threading
    import time
    from stringsearchlib_b200 import synth
    n, nthreads, per = 20000, 16, 64
    keys = synth.ascii_keys(n)
    q = synth.queries_from(keys, nthreads * per)
    queries = [q.get(i) for i in range(len(q))]
    S = oracle_built.SpecIndex(keys, 1)
    counts, ext, sc = S.search_batch(q, 0.0, 20)
    kid_of_ext = S.tables()["kidOfExt"]
    with ssb.Index(keys, 1) as ix:
        got = [None] * len(queries)
        errors = []

        def work(t):
            try:
                for i in range(t * per, (t + 1) * per):
                    got[i] = ix.search_one(queries[i], 0.0, 20)
            except Exception as e:      # noqa: BLE001 - reported below
                errors.append(e)

        th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        assert not errors, errors
        for i in range(len(queries)):
            c = int(counts[i])
            kids, scores = got[i]
            assert np.array_equal(kids, kid_of_ext[ext[i, :c]]), i
            assert np.array_equal(scores.view(np.uint32), sc[i, :c].view(np.uint32)), i
        st = ix.search_one_stats()
        assert st["calls"] == len(queries) and st["batches"] <= st["calls"]
        print(f"\n{len(queries)} concurrent single queries from {nthreads} threads: {len(queries) / dt:.0f} q/s, "
              f"{st['calls'] / st['batches']:.1f} queries per GPU batch")
    # the as-shipped ABI from several threads
    words = [keys.get(i) for i in range(2000)]
    H = ssb.HandleLibrary()
    h = H.indexN(words, 1)
    single = [H.score(h, queries[i], 0.0, 10) for i in range(64)]
    out = [None] * 64

    def work2(t):
        for i in range(t * 8, (t + 1) * 8):
            out[i] = H.score(h, queries[i], 0.0, 10)

    th = [threading.Thread(target=work2, args=(t,)) for t in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert out == single
    H.lib.dispose(h)
