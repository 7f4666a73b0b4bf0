"""k_fast_search time vs index size and batch size (device-resident), to separate fixed per-query cost from posting work."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import stringsearchlib_b200 as ssb
from stringsearchlib_b200 import synth
for nkeys in (1_000_000, 2_500_000, 5_000_000, 10_000_000):
    keys = synth.ascii_keys(nkeys)
    ix = ssb.Index(keys, 1)
    for nq in (1024, 4096, 16384):
        q = synth.ascii_queries(nq, nkeys)
        ix.search_batch(q, 0.0, 100)
        ts = []
        for _ in range(5):
            ms, msk, st = ix.search_resident(nq, 0.0, 100)
            ts.append(msk)
        print(f"keys={nkeys:9d} nq={nq:6d} k_fast={np.median(ts):7.3f} ms  per-query={1e3*np.median(ts)/nq:6.3f} us  postings/q={st['postings']/nq:8.0f}", flush=True)
    ix.close()
