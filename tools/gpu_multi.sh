#!/bin/bash
# multi-GPU bench (one process per GPU under torchrun).  usage: tools/gpu_multi.sh <ngpus> <tag>
N=${1:-2}; TAG=${2:-m}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "rc=$?"; tail -c 2500 gpurun_out/bench_${TAG}_n$N.json; tail -5 gpurun_out/bench_${TAG}_n$N.err
