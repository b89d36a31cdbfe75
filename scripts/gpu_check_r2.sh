#!/bin/bash
# One gpurun call: GPU test suite, kernel timings, default bench line, DRAM traffic of the scan kernel.
#   usage: gpu_check_r2.sh <tag>
tag=${1:-chk}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/r2${tag}_pytest.log; tail -2 gpurun_out/r2${tag}_pytest.log
timeout 200 python scripts/k1_time.py 2>&1 | tail -1 | tee gpurun_out/r2${tag}_ktime.log
timeout 600 python bench.py > gpurun_out/r2${tag}_bench_n1.json 2> gpurun_out/r2${tag}_bench_n1.err; tail -c 1500 gpurun_out/r2${tag}_bench_n1.json
bash scripts/ncu_traffic.sh 2>&1 | tail -1
