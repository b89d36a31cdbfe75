#!/bin/bash
# One gpurun call: the GPU test suite, kernel timings under scripts/exp_list.txt, and a full ncu capture of
# the scan kernel (third launch).
tag=${1:-k1}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/r2${tag}_pytest.log; tail -2 gpurun_out/r2${tag}_pytest.log
bash scripts/gpu_exp.sh ${tag} < scripts/exp_list.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel" -s 2 -c 1 -f \
    -o gpurun_out/prof_k1_${tag} python scripts/ncu_target.py > gpurun_out/ncu_k1_${tag}.log 2>&1; tail -1 gpurun_out/ncu_k1_${tag}.log
