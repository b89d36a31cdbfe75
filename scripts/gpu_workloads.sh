#!/bin/bash
# the non-default bench workloads (BASELINE.json configs[2..4]) on N GPUs; JSON lines under gpurun_out/
tag=${1:-a}; n=${2:-1}
mkdir -p gpurun_out
run() {  # name, extra args...
  name=$1; shift
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --workload $name "$@" > gpurun_out/r2${tag}_${name}_n$n.json 2> gpurun_out/r2${tag}_${name}_n$n.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29621 \
        bench.py --gpus $n --workload $name "$@" > gpurun_out/r2${tag}_${name}_n$n.json 2> gpurun_out/r2${tag}_${name}_n$n.err
  fi
  echo "== $name rc=$?"; tail -c 1500 gpurun_out/r2${tag}_${name}_n$n.json; tail -3 gpurun_out/r2${tag}_${name}_n$n.err
}
run df17_aggressive --steps 10 --warmup 3
run snr_sweep --frames ${3:-10000}
run tiled_64g --steps 3 --warmup 1
