#!/bin/bash
# round-end evidence: parity tests, the default bench line, the reference arm, launch list and one full ncu capture of k_sieve_search
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; tail -1 gpurun_out/r02_bench_final_n1.json | cut -c1-400
python bench.py --impl reference > gpurun_out/r02_bench_final_reference_arm.json 2> gpurun_out/ref.err; tail -1 gpurun_out/r02_bench_final_reference_arm.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_sieve_search|k_fast_search|k_query_prep|k_order_fast|k_clamp_counts|k_finalize" -c 60 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sieve_search -s 3 -c 1 -o gpurun_out/r02_sieve_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log | cut -c1-120
./stringsearchlib_b200/lib/random_sectors > gpurun_out/r02_random_sectors.txt 2>&1; tail -3 gpurun_out/r02_random_sectors.txt
