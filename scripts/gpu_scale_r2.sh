#!/bin/bash
# N-GPU validation: default bench, the three other workloads, the multi-GPU parity script
tag=${1:-a}; n=${2:-4}
mkdir -p gpurun_out
tr() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
tr 600 29631 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2${tag}_bench_n$n.json 2> gpurun_out/r2${tag}_bench_n$n.err
tail -c 1800 gpurun_out/r2${tag}_bench_n$n.json; tail -2 gpurun_out/r2${tag}_bench_n$n.err
tr 600 29632 bench.py --gpus $n --workload snr_sweep --frames 10000 > gpurun_out/r2${tag}_snr_sweep_n$n.json 2> gpurun_out/r2${tag}_snr_sweep_n$n.err
tail -c 600 gpurun_out/r2${tag}_snr_sweep_n$n.json; tail -2 gpurun_out/r2${tag}_snr_sweep_n$n.err
tr 900 29633 bench.py --gpus $n --workload tiled_64g --steps 3 --warmup 1 > gpurun_out/r2${tag}_tiled_64g_n$n.json 2> gpurun_out/r2${tag}_tiled_64g_n$n.err
tail -c 1500 gpurun_out/r2${tag}_tiled_64g_n$n.json; tail -2 gpurun_out/r2${tag}_tiled_64g_n$n.err
tr 600 29634 scripts/multi_gpu_parity.py > gpurun_out/r2${tag}_parity_n$n.log 2>&1; tail -2 gpurun_out/r2${tag}_parity_n$n.log
