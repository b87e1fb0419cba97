#!/bin/bash
# multi-GPU bench (one process per GPU under torchrun).  usage: tools/gpu_multi.sh <ngpus> <tag> [extra bench args]
N=${1:-2}; TAG=${2:-m}; shift; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "rc=$?"; python - <<P
import json
for l in open('gpurun_out/bench_${TAG}_n$N.json'):
    if l.startswith('{'):
        j=json.loads(l); print($N, 'QPS', round(j['value']), 'ms', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), 'recall', j['config']['recall_at_10'], j['config']['parallelism']); print(j['stages_ms'])
P
tail -3 gpurun_out/bench_${TAG}_n$N.err | grep -v "^\*\|OMP_NUM\|^$"
