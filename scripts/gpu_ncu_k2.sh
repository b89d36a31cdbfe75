#!/bin/bash
# One gpurun call: full ncu capture of the frame-evaluation kernel selected by $1 (MODES_EVAL_VARIANT), third launch.
v=${1:-fused}
mkdir -p gpurun_out
MODES_EVAL_VARIANT=$v timeout 500 ncu --set full --clock-control none --import-source on -k regex:"eval_" -s 2 -c 1 -f \
    -o gpurun_out/prof_k2_$v python scripts/ncu_target.py > gpurun_out/ncu_k2_$v.log 2>&1; tail -2 gpurun_out/ncu_k2_$v.log
