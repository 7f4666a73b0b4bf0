#!/bin/bash
# one GPU iteration: parity tests, a short bench line, optional ncu capture of the sieve kernel (arg 1 = report name)
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
SSB_DEBUG=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b1.log 2> gpurun_out/b1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b1.log").read().strip().splitlines()[-1])
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"kernel_ms",round(d["roofline"]["kernel_ms"],4),"frac",round(d["roofline"]["frac"],4),d["queries_by_kernel"],d["index_build"]["rank0_ms"])
PY
grep -v "^$" gpurun_out/b1.err | grep sieve | tail -2
if [ -n "$1" ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_sieve_search -s 3 -c 1 -o gpurun_out/$1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  tail -2 gpurun_out/ncu.log | cut -c1-200
fi
