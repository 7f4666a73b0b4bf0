#!/bin/bash
# block size of the sieve (ids per block = 2^shift) on the headline workload: kernel time per batch
for sh in 5 6 7; do
  SSB_SIEVE_SHIFT=$sh SSB_DEBUG=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-micro 2>/tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('shift $sh', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']), d['index_build']['device_bytes'])"
  grep "sieve_search per query" /tmp/err.txt | tail -1
done
