#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('10M default', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']))"
cp stringsearchlib_b200/lib/libssb_core.so /tmp/core_orig.so
for v in default nodyn; do
  if [ $v != default ]; then cp build/variants/$v.so stringsearchlib_b200/lib/libssb_core.so; fi
  SSB_DEBUG=1 python bench.py --keys 100000000 --steps 5 --warmup 3 --no-cpu-baseline 2>/tmp/err.txt > gpurun_out/100m_$v.json
  tail -1 gpurun_out/100m_$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('100M $v', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['roofline']['kernel_ms'], d['roofline']['frac'])"
  grep "sieve_search per query" /tmp/err.txt | tail -1
done
cp /tmp/core_orig.so stringsearchlib_b200/lib/libssb_core.so
