#!/bin/bash
# One gpurun call: the receiver pool on the GPU (parity tests, then the receivers workload at several pool sizes).
tag=${1:-pool}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_pool.py tests/test_gpu_parity.py -m gpu -x -q -k "pool or fused" 2>&1 | tail -6 ) > gpurun_out/r2${tag}_pytest.log; tail -3 gpurun_out/r2${tag}_pytest.log
for r in 64 256 1024; do
  timeout 600 python bench.py --workload receivers --receivers $r --steps 20 --warmup 3 > gpurun_out/r2${tag}_receivers_$r.json 2> gpurun_out/r2${tag}_receivers_$r.err
  echo "== receivers $r rc=$?"; tail -c 900 gpurun_out/r2${tag}_receivers_$r.json; tail -2 gpurun_out/r2${tag}_receivers_$r.err
done
