#!/bin/bash
# bench.py on N GPUs (default 1) + multi-GPU parity script; outputs under gpurun_out/
tag=${1:-a}; n=${2:-1}; shift; shift
mkdir -p gpurun_out
if [ "$n" = "1" ]; then
  timeout 600 python bench.py "$@" > gpurun_out/r2${tag}_bench_n1.json 2> gpurun_out/r2${tag}_bench_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $n "$@" > gpurun_out/r2${tag}_bench_n$n.json 2> gpurun_out/r2${tag}_bench_n$n.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 \
      scripts/multi_gpu_parity.py > gpurun_out/r2${tag}_parity_n$n.log 2>&1
  tail -4 gpurun_out/r2${tag}_parity_n$n.log
fi
tail -c 3000 gpurun_out/r2${tag}_bench_n$n.json; tail -5 gpurun_out/r2${tag}_bench_n$n.err
