import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import stringsearchlib_b200 as ssb
from stringsearchlib_b200 import synth
nkeys = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
keys = synth.ascii_keys(nkeys)
ix = ssb.Index(keys, 1)
q = synth.ascii_queries(4096, nkeys)
ix.search_batch(q, 0.0, 100)
for _ in range(4):
    ms, msk, st = ix.search_resident(4096, 0.0, 100)
print(msk, st)
