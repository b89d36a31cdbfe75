#!/bin/bash
# Round-2 record run (one GPU): GPU tests, default bench and reference arm as the driver runs them,
# ncu launch list of the bench, full ncu capture of both hot kernels, DRAM traffic tied to the build,
# compute-sanitizer memcheck.  Everything under gpurun_out/r2final_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2final_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2final_pytest_gpu.log; tail -2 gpurun_out/r2final_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2final_bench_n1.json 2> gpurun_out/r2final_bench_n1.err; tail -c 600 gpurun_out/r2final_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2final_bench_reference_arm.json 2> gpurun_out/r2final_bench_reference_arm.err; tail -c 900 gpurun_out/r2final_bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2final_launches_bench_n1.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/r2final_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel|eval_serial_kernel" -s 2 -c 2 -f \
    -o gpurun_out/prof_r2final python scripts/ncu_target.py > gpurun_out/r2final_ncu_full.log 2>&1; tail -1 gpurun_out/r2final_ncu_full.log
bash scripts/ncu_traffic.sh 2>&1 | tail -1
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > gpurun_out/r2final_sanitizer_memcheck.log 2>&1; tail -2 gpurun_out/r2final_sanitizer_memcheck.log
