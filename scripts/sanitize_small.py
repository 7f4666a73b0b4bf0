"""small end-to-end run for compute-sanitizer: build + fast path + generic fast path + dense path + shard merge."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import stringsearchlib_b200 as ssb
from stringsearchlib_b200 import synth, api
keys = synth.ascii_keys(30000)
q = synth.queries_from(keys, 256)
with ssb.Index(keys, 1) as ix:
    for path in (1, 0, -1):                      # k_sieve_search for every query it takes, k_fast_search alone, the default choice
        ssb.set_search_path(path)
        r = ix.search_batch(q, 0.0, 100)
        r = ix.search_batch(q, 0.3, 1000)
        r = ix.search_batch(q, 0.0, 2000)
        print("identity", ix.info()["identity"], path, len(r.kids))
ssb.set_search_path(1)
words, w = synth.key_desc_rows(8000)
with ssb.Index(words, 2, w) as ix:
    r = ix.search_batch(synth.queries_from(words, 128, stride=2), 0.0, 50)
    print("generic", len(r.kids))
wk = synth.wide_keys(5000)
with ssb.Index(wk, 1, None, 2) as ix:
    r = ix.search_batch(synth.queries_from(wk, 64), 0.0, 50)
    print("wide", len(r.kids))
msg = [b"LWMS", b"LWM", b"LWMA", b"LWYY", b"L", b"I", b"GHRSDGSDGS Egdsrtg g", None, b"  ", b"ab#c d"] * 20
with ssb.Index(msg, 2) as ix:
    r = ix.search_batch([b"LWMS", b"*", b"", b"L", b"lw", b"ghrsdgsdgs egdsrtg"], 0.0, 0)
    print("dense", len(r.kids))
rb = api.RECORD_DTYPE.itemsize
N, nq, limit = 2, 64, 10
recv = [api.device_alloc(N * (nq // N) * limit * rb) for _ in range(N)]
merged = [api.device_alloc(nq * limit * rb) for _ in range(N)]
cnts = [api.device_alloc(nq * 4) for _ in range(N)]
shards = []
for r_ in range(N):
    ix = ssb.Index(keys.slice(r_ * 15000, (r_ + 1) * 15000), 1, row_base=r_ * 15000)
    ix.search_batch(synth.queries_from(keys, nq), 0.0, limit)
    ix.export_records_p2p(nq, limit, recv, r_)
    shards.append(ix)
for r_ in range(N):
    api.merge_records_p2p(recv[r_], N, nq // N, limit, merged, cnts, r_)
mc = np.zeros(nq, np.uint32)
api.memcpy_d2h(mc, cnts[0])
print("merge", int(mc.sum()))
