#!/bin/bash
# Round-2 record run, final build (one GPU): GPU tests, default bench and reference arm as the driver runs
# them, ncu launch list of the bench, full ncu capture of both hot kernels, DRAM traffic tied to the scan
# kernel's sources, compute-sanitizer memcheck and racecheck.  Everything under gpurun_out/r2fin_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2fin_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2fin_pytest_gpu.log; tail -2 gpurun_out/r2fin_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2fin_bench_n1.json 2> gpurun_out/r2fin_bench_n1.err; tail -c 700 gpurun_out/r2fin_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2fin_bench_reference_arm.json 2> gpurun_out/r2fin_bench_reference_arm.err; tail -c 400 gpurun_out/r2fin_bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2fin_launches_bench_n1.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/r2fin_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel|eval_fused_kernel" -s 4 -c 2 -f \
    -o gpurun_out/prof_r2fin python scripts/ncu_target.py > gpurun_out/r2fin_ncu_full.log 2>&1; tail -1 gpurun_out/r2fin_ncu_full.log
bash scripts/ncu_traffic.sh 2>&1 | tail -1
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > gpurun_out/r2fin_sanitizer_memcheck.log 2>&1; tail -2 gpurun_out/r2fin_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python scripts/sanitize_target.py > gpurun_out/r2fin_sanitizer_racecheck.log 2>&1; tail -2 gpurun_out/r2fin_sanitizer_racecheck.log
