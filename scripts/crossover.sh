#!/bin/bash
# k_sieve_search vs k_fast_search (SSB_MERGE=1) by index size, single GPU, batch 4096
for k in 1250000 2500000 5000000; do
  for m in sieve merge; do
    if [ $m = merge ]; then export SSB_MERGE=1; else unset SSB_MERGE; fi
    python bench.py --keys $k --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$k $m', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']))"
  done
done
