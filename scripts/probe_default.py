"""kernel time of the batched fast path on the BASELINE.json configs with the DEFAULT choice of kernel, next to both forced paths."""
import sys
sys.path.insert(0, '/root/repo')
import stringsearchlib_b200 as ssb
from stringsearchlib_b200 import synth

def run(name, ix, q, th, limit):
    out = []
    for path in (-1, 0, 1):
        ssb.set_search_path(path)
        ix.search_batch(q, th, limit)
        best = 1e9
        for _ in range(5):
            ms, msk, st = ix.search_resident(len(q), th, limit)
            best = min(best, msk)
        out.append((path, round(best, 4), st["nSieve"]))
    ssb.set_search_path(-1)
    print(name, out, flush=True)

words, w = synth.key_desc_rows(1_000_000)
with ssb.Index(words, 2, w) as ix:
    q = synth.queries_from(words, 4096, stride=2)
    run("C2 1M rows x2, weights, th=0.3, limit=100", ix, q, 0.3, 100)
for n in (1_250_000, 2_500_000):
    keys = synth.ascii_keys(n)
    with ssb.Index(keys, 1) as ix:
        q = synth.ascii_queries(4096, n)
        run(f"{n} keys th=0.5 limit=1000", ix, q, 0.5, 1000)
        run(f"{n} keys th=0.3 limit=100", ix, q, 0.3, 100)
