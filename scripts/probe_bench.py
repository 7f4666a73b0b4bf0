"""10M-key (or argv[1]) index, a few resident batches; prints kernel ms + stats. Used under ncu for per-kernel captures."""
import sys
sys.path.insert(0, '/root/repo')
import stringsearchlib_b200 as ssb
from stringsearchlib_b200 import synth
nkeys = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 100
keys = synth.ascii_keys(nkeys)
ix = ssb.Index(keys, 1)
q = synth.ascii_queries(4096, nkeys)
ix.search_batch(q, 0.0, limit)
for _ in range(reps):
    ms, msk, st = ix.search_resident(4096, 0.0, limit)
print(ms, msk, st)
