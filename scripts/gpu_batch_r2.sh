#!/bin/bash
# One gpurun call of round 2: GPU test suite (optionally under a non-default scan kernel), kernel
# timings per variant, one full ncu capture of the scan and frame-evaluation kernels.
#   usage: gpu_batch_r2.sh <tag> [scan variants to time, default "1 2"] [pytest -k expression]
tag=${1:-a}
variants=${2:-"1 2"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2${tag}_smi.txt 2>&1
( MODES_SCAN_VARIANT=2 timeout 900 python -m pytest tests -m gpu -x -q ${3:+-k "$3"} 2>&1 | tail -15 ) > gpurun_out/r2${tag}_pytest_v2.log
tail -3 gpurun_out/r2${tag}_pytest_v2.log
for sv in $variants; do
  MODES_SCAN_VARIANT=$sv MODES_EVAL_VARIANT=lean timeout 200 python scripts/k1_time.py 2>&1 | tail -1
done | tee gpurun_out/r2${tag}_ktime.log
MODES_SCAN_VARIANT=2 MODES_EVAL_VARIANT=lean timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:"scan2_kernel|eval_serial_kernel" -s 2 -c 2 -f -o gpurun_out/prof_r2${tag} python scripts/ncu_target.py > gpurun_out/r2${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/r2${tag}_ncu_full.log
