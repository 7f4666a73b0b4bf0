#!/bin/bash
# A/B of the libssb_core variants under build/variants: alternating runs of the short bench, kernel time of each
cp stringsearchlib_b200/lib/libssb_core.so /tmp/core_orig.so
for rep in 1 2 3; do
  for f in build/variants/*.so; do
    cp $f stringsearchlib_b200/lib/libssb_core.so
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$f', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']))"
  done
done
cp /tmp/core_orig.so stringsearchlib_b200/lib/libssb_core.so
