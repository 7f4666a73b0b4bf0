#!/usr/bin/env python
"""Generates tests/golden/reference_golden.json by running the UNMODIFIED reference (oracle/_ref/libref.so, built from
/root/reference by `make -C oracle ref`) through its own C ABI (indexN / score / getSize / getLibSize,
dllmain.cpp:37-140) and by dumping its private tables (oracle/ref_dump.cpp). Run where /root/reference exists:

    python tests/golden/make_golden.py

The fixture travels to the GPU box; nothing at test time needs /root/reference. Inputs avoid the zones where the
reference itself is iteration-order dependent: weights > 100 or < 0, wildcard over keys that have
several edges with different weights.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import loader  # noqa: E402
from util import f32_bits, random_corpus  # noqa: E402

SEVEN = [b"LWMS", b"LWM", b"LWMA", b"LWYY", b"L", b"I", b"GHRSDGSDGS Egdsrtg g"]   # SearchTest/test.cpp:7-9


def enc(b):
    return None if b is None else b.hex()


def run_case(R, name, words, row_size, weights, queries):
    h = R.indexN(words, row_size, weights)
    terms, grams, counts = loader.ref_tables(words, row_size, weights)
    case = dict(name=name, words=[enc(w) for w in words], rowSize=row_size,
                weights=None if weights is None else [float(x) for x in weights],
                getSize=int(R.L.getSize(h)), getLibSize=int(R.L.getLibSize(h)),
                terms={k.hex(): {kk.hex(): f32_bits(v) for kk, v in d.items()} for k, d in terms.items()},
                grams={str(g): sorted(t.hex() for t in s) for g, s in grams.items()},
                queries=[])
    for q, th, lim in queries:
        res = R.score(h, q, th, lim)
        case["queries"].append(dict(q=q.hex(), threshold=th, limit=lim, results=[[k.hex(), f32_bits(s)] for k, s in res]))
    R.dispose(h)
    return case


def main():
    loader.build(ref=True)
    R = loader.RefLib()
    cases = []
    # 1. the reference's own fixture (SearchTest/test.cpp:7-18) with rowSize=1 and rowSize=7
    q7 = [(b"LWMS", 0.5, 0), (b"LWMS", 0.0, 0), (b"LWM", 0.0, 0), (b"ghrsdgsdgs", 0.0, 0), (b"SDGSDG", 0.0, 0), (b"*", 0.0, 0),
          (b"", 0.0, 0), (b"   ", 0.0, 0), (b"LWMS", 0.5, 2), (b"lw", 0.0, 0), (b"L", 0.0, 0), (b"Egdsrtg", 0.3, 0),
          (b"GHRSDGSDGS Egdsrtg g", 0.0, 0), (b"ghrsdgsdgs egdsrtg g", 0.0, 0)]
    cases.append(run_case(R, "searchtest_row1", SEVEN, 1, None, q7))
    cases.append(run_case(R, "searchtest_row7", SEVEN, 7, None, q7))
    cases.append(run_case(R, "one_word", [b"LWMS"], 1, None, [(b"LWMS", 0.0, 0)]))
    # 2. seeded random corpora: duplicates, blanks, invalid characters, NULL cells, per-word weights (incl. 0)
    rng = np.random.default_rng(20240607)
    for ci, (rows, rs, weighted) in enumerate([(60, 1, False), (80, 2, True), (50, 3, True), (120, 1, True), (40, 4, False)]):
        words, weights = random_corpus(rng, rows, rs, weighted)
        qs = []
        pool = [w for w in words if w]
        for _ in range(24):
            w = bytearray(pool[int(rng.integers(len(pool)))])
            if w and rng.random() < 0.7:
                w[int(rng.integers(len(w)))] = int(rng.choice(list(b"abcABC12 .#")))
            if rng.random() < 0.3:
                w = w + bytes(rng.choice(list(b"abc12"), size=int(rng.integers(1, 6))).astype(np.uint8))
            qs.append((bytes(w), float(rng.choice([0.0, 0.0, 0.3, 0.5, 0.75])), int(rng.choice([0, 0, 3, 10]))))
        if not weighted:
            qs.append((b"*", 0.0, 0))
        cases.append(run_case(R, f"random_{ci}", words, rs, weights, qs))
    # 3. longer lower-case keys (the benchmark distribution), every term on the n-gram path
    from stringsearchlib_b200 import synth
    keys = synth.ascii_keys(400)
    qs = synth.queries_from(keys, 40)
    words = [keys.get(i) for i in range(len(keys))]
    cases.append(run_case(R, "synthetic_400", words, 1, None,
                          [(qs.get(i), th, lim) for i in range(len(qs)) for th, lim in ((0.0, 100), (0.3, 100), (0.0, 5))][:90]))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")
    with open(out, "w") as f:
        json.dump(dict(source="oracle/_ref/libref.so (unmodified reference, g++ -O2)", cases=cases), f, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes,", len(cases), "cases,", sum(len(c["queries"]) for c in cases), "queries")


if __name__ == "__main__":
    main()
