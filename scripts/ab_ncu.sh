#!/bin/bash
# per-launch durations of k_sieve_search under ncu for every variant, with and without ncu's cache flush between kernels
cp stringsearchlib_b200/lib/libssb_core.so /tmp/core_orig.so
for f in build/variants/*.so; do
  cp $f stringsearchlib_b200/lib/libssb_core.so
  for cc in all none; do
    ncu --metrics gpu__time_duration.sum --clock-control none --cache-control $cc -k regex:k_sieve_search -c 8 --csv --log-file /tmp/l.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
    echo "$f cache-control=$cc: $(grep k_sieve_search /tmp/l.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
  done
done
cp /tmp/core_orig.so stringsearchlib_b200/lib/libssb_core.so
