"""C1: 10k synthetic ASCII words, rowSize=1, gSize=3, single search() threshold=0 limit=100 — per-call latency through the
as-shipped handle ABI (score/release), next to the compiled reference on the host (if oracle/_ref is present)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import stringsearchlib_b200 as ssb
from stringsearchlib_b200 import synth
keys = synth.ascii_keys(10000)
words = [keys.get(i) for i in range(10000)]
qs = synth.queries_from(keys, 1000)
queries = [qs.get(i) for i in range(1000)]
H = ssb.HandleLibrary()
h = H.indexN(words, 1)
for q in queries[:50]:
    H.score(h, q, 0.0, 100)
t0 = time.perf_counter()
n = 0
for q in queries:
    n += len(H.score(h, q, 0.0, 100))
dt = time.perf_counter() - t0
print(f"B200 handle ABI: {1000/dt:.0f} q/s single-call ({1e6*dt/1000:.1f} us/call), {n/1000:.1f} results/query")
seven = [b"LWMS", b"LWM", b"LWMA", b"LWYY", b"L", b"I", b"GHRSDGSDGS Egdsrtg g"]
h2 = H.indexN(seven, 1)
t0 = time.perf_counter()
for i in range(200):
    H.score(h2, "LWMS", 0.5, 0)
dt = time.perf_counter() - t0
print(f"B200 dense path (SearchTest fixture): {200/dt:.0f} q/s ({1e6*dt/200:.1f} us/call)")
try:
    from oracle import loader
    R = loader.RefLib()
    hr = R.indexN(words, 1)
    t0 = time.perf_counter()
    for q in queries:
        R.score(hr, q, 0.0, 100)
    dt = time.perf_counter() - t0
    print(f"reference CPU: {1000/dt:.0f} q/s single-thread ({1e6*dt/1000:.1f} us/call)")
except Exception as e:
    print("reference unavailable", e)
