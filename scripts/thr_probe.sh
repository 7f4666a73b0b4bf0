#!/bin/bash
# the headline index under thresholds: kernel time per batch
for cfg in "0.3 100" "0.5 500"; do
  set -- $cfg
  python bench.py --threshold $1 --limit $2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threshold $1 limit $2', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']), d['queries_by_kernel'])"
done
