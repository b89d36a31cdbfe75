#!/bin/bash
# One gpurun call: the whole GPU test suite (incl. the receiver pool), then the receivers workload at several pool sizes.
tag=${1:-pool}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2${tag}_pytest.log; tail -3 gpurun_out/r2${tag}_pytest.log
for r in 64 256 1024; do
  timeout 600 python bench.py --workload receivers --receivers $r --steps 20 --warmup 3 > gpurun_out/r2${tag}_receivers_$r.json 2> gpurun_out/r2${tag}_receivers_$r.err
  echo "== receivers $r rc=$?"; tail -c 400 gpurun_out/r2${tag}_receivers_$r.json; tail -2 gpurun_out/r2${tag}_receivers_$r.err
done
