#!/bin/bash
# k_sieve_search (forced: SSB_SIEVE_MIN_POSTINGS=0) vs k_fast_search (SSB_MERGE=1) by index size, single GPU, batch 4096
for k in ${KEYS:-1250000 2500000 5000000}; do
  for m in sieve merge; do
    if [ $m = merge ]; then export SSB_MERGE=1; unset SSB_SIEVE_MIN_POSTINGS; else unset SSB_MERGE; export SSB_SIEVE_MIN_POSTINGS=0; fi
    python bench.py --keys $k --steps 10 --warmup 3 --no-cpu-baseline --no-micro 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$k $m', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']), d['queries_by_kernel'])"
  done
done
