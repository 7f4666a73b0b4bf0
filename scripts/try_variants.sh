#!/bin/bash
# runs the short bench with every libssb_core variant under build/variants (each replaces the in-tree library for its run)
cp stringsearchlib_b200/lib/libssb_core.so /tmp/core_orig.so
for f in build/variants/*.so; do
  cp $f stringsearchlib_b200/lib/libssb_core.so
  SSB_DEBUG=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value']))"
  grep sieve /tmp/err.txt | tail -2
done
cp /tmp/core_orig.so stringsearchlib_b200/lib/libssb_core.so
