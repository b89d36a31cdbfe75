#!/bin/bash
# One gpurun call: GPU parity of the frame-evaluation kernels, their timings (scripts/exp_list.txt),
# and a full ncu capture of the variant named by $2 (optional).
tag=${1:-k2}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused" 2>&1 | tail -8 ) > gpurun_out/r2${tag}_pytest.log; tail -3 gpurun_out/r2${tag}_pytest.log
bash scripts/gpu_exp.sh ${tag} < scripts/exp_list.txt
if [ -n "$2" ]; then bash scripts/gpu_ncu_k2.sh $2; fi
