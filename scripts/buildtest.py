import sys, time
sys.path.insert(0, '/root/repo')
import stringsearchlib_b200 as ssb
from stringsearchlib_b200 import synth
keys = synth.ascii_keys(10_000_000)
ptrs = keys.pointers()
for it in range(4):
    t0 = time.perf_counter()
    ix = ssb.Index.from_pointers(ptrs, len(keys), 1, None, 3, False)
    dt = time.perf_counter() - t0
    print(it, round(dt, 3), {k: round(v, 1) for k, v in ix.build_ms().items()}, flush=True)
    ix.close()
