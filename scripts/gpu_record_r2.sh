#!/bin/bash
# Refresh of the N=1 records on the last build: default bench line, ncu launch list, full ncu capture of both hot kernels.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2rec_bench_n1.json 2> gpurun_out/r2rec_bench_n1.err; tail -c 900 gpurun_out/r2rec_bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2rec_launches_bench_n1.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/r2rec_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel|eval_fused_kernel" -s 4 -c 2 -f \
    -o gpurun_out/prof_r2rec python scripts/ncu_target.py > gpurun_out/r2rec_ncu_full.log 2>&1; tail -1 gpurun_out/r2rec_ncu_full.log
python __graft_entry__.py smoke 2>&1 | tail -1
