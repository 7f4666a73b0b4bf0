"""per-query duration lines of a -DSV_TIMING=1 build (SVQ ... printed by k_sieve_search) on the headline workload."""
import sys
sys.path.insert(0, '/root/repo')
import stringsearchlib_b200 as ssb
from stringsearchlib_b200 import synth
n = 10_000_000
keys = synth.ascii_keys(n)
with ssb.Index(keys, 1) as ix:
    q = synth.ascii_queries(4096, n)
    ix.search_batch(q, 0.0, 100)
    print("SVQ ---- second run", flush=True)
    ms, msk, st = ix.search_resident(len(q), 0.0, 100)
    print("kernel ms", msk, flush=True)
