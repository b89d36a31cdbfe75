#!/bin/bash
# experiment batch: kernel timings under a list of "ENV=VAL ..." settings (one per line on stdin)
tag=${1:-x}
mkdir -p gpurun_out
while read -r line; do
  [ -z "$line" ] && continue
  echo "== $line"
  env $line timeout 200 python scripts/k1_time.py 2>&1 | tail -1
done | tee gpurun_out/r2${tag}_exp.log
