#!/bin/bash
# N-GPU check of the final build: the default bench as the driver launches it and the multi-GPU parity script.
tag=${1:-fin}; n=${2:-2}
mkdir -p gpurun_out
tr() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
tr 600 29631 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2${tag}_bench_n$n.json 2> gpurun_out/r2${tag}_bench_n$n.err
tail -c 2500 gpurun_out/r2${tag}_bench_n$n.json; tail -2 gpurun_out/r2${tag}_bench_n$n.err
tr 600 29634 scripts/multi_gpu_parity.py > gpurun_out/r2${tag}_parity_n$n.log 2>&1; tail -2 gpurun_out/r2${tag}_parity_n$n.log
