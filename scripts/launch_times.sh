#!/bin/bash
# per-launch durations of the search kernels of one bench step (ncu --metrics gpu__time_duration.sum; serialised, cold-cache)
SSB_TILES=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tile|k_sieve_search|k_fast_search|k_query_prep|k_order" -c 64 --csv --log-file gpurun_out/tile_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/tile_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
out=[(r[ki].split('::')[-1][:28], float(r[vi])/1e3) for r in rows[1:]]
for k,v in out[-32:]: print(f"{k:30s} {v:8.1f} us")
PY
