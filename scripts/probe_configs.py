"""kernel time of the batched fast path on the BASELINE.json configs C1-C3 (and C4 with a threshold), both search paths."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import stringsearchlib_b200 as ssb
from stringsearchlib_b200 import synth

def run(name, ix, q, th, limit):
    out = []
    for path in (0, 1):
        ssb.set_search_path(path)
        ix.search_batch(q, th, limit)
        best = 1e9
        for _ in range(5):
            ms, msk, st = ix.search_resident(len(q), th, limit)
            best = min(best, msk)
        out.append((path, round(best, 4), st["nSieve"], st["nSieveHandover"], st["nDense"]))
    ssb.set_search_path(-1)
    print(name, out, flush=True)

words, w = synth.key_desc_rows(1_000_000)
with ssb.Index(words, 2, w) as ix:
    q = synth.queries_from(words, 4096, stride=2)
    run("C2 1M rows x2, weights, th=0.3, limit=100", ix, q, 0.3, 100)
    run("C2 same, th=0", ix, q, 0.0, 100)
wk = synth.wide_keys(1_000_000)
with ssb.Index(wk, 1, None, 2) as ix:
    run("C3 wide 1M gSize=2 limit=50", ix, synth.queries_from(wk, 4096), 0.0, 50)
keys = synth.ascii_keys(10_000_000)
with ssb.Index(keys, 1) as ix:
    q = synth.ascii_queries(4096, 10_000_000)
    run("C4 10M th=0", ix, q, 0.0, 100)
    run("C4 10M th=0.3", ix, q, 0.3, 100)
    run("C4 10M th=0.5 limit=500", ix, q, 0.5, 500)
