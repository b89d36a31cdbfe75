#!/bin/bash
# One gpurun call of round 2: GPU test suite, kernel timings, one full ncu capture of the scan and
# frame-evaluation kernels.
#   usage: gpu_batch_r2.sh <tag> [pytest -k expression]
tag=${1:-a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2${tag}_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} 2>&1 | tail -15 ) > gpurun_out/r2${tag}_pytest.log
tail -3 gpurun_out/r2${tag}_pytest.log
timeout 200 python scripts/k1_time.py 2>&1 | tail -1 | tee gpurun_out/r2${tag}_ktime.log
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:"scan_kernel|eval_serial_kernel" -s 2 -c 2 -f -o gpurun_out/prof_r2${tag} python scripts/ncu_target.py > gpurun_out/r2${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/r2${tag}_ncu_full.log
