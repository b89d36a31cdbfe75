#!/bin/bash
# One gpurun call: default bench, ncu launch list of the bench, one full ncu capture of each hot
# kernel, host-side resolve timing.  Outputs under gpurun_out/ with the given tag.
tag=${1:-x}
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err
cat gpurun_out/bench_n1_$tag.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench_$tag.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel|eval_serial_kernel" -s 2 -c 2 \
    -f -o gpurun_out/prof_r1$tag python scripts/ncu_target.py > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
for tool in memcheck racecheck; do
  timeout 200 compute-sanitizer --tool $tool python scripts/sanitize_target.py > gpurun_out/sanitizer_${tool}_$tag.log 2>&1
  tail -3 gpurun_out/sanitizer_${tool}_$tag.log
done
if [ -n "$2" ]; then
  for t in 1 8 16 32; do
    echo "== MODES_BUILD_THREADS=$t"
    MODES_BUILD_THREADS=$t MODES_RESOLVE_TIMING=1 timeout 200 python tests/resolve_probe.py 1024 8 2>&1 | tail -5
  done > gpurun_out/resolve_probe_$tag.log 2>&1
  cat gpurun_out/resolve_probe_$tag.log
fi
