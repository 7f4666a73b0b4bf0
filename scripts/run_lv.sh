#!/bin/bash
cp stringsearchlib_b200/lib/libssb_core.so /tmp/core_orig.so
for rep in 1 2; do for v in default lv3; do
  if [ $v != default ]; then cp build/variants/$v.so stringsearchlib_b200/lib/libssb_core.so; else cp /tmp/core_orig.so stringsearchlib_b200/lib/libssb_core.so; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('10M $v', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']))"
done; done
cp build/variants/lv3.so stringsearchlib_b200/lib/libssb_core.so
python bench.py --keys 100000000 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('100M lv3', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'])"
cp /tmp/core_orig.so stringsearchlib_b200/lib/libssb_core.so
