#!/bin/bash
# N GPUs: the bench line under torchrun (replica grid, graph replay on), configs[1] only.  usage: tools/gpu_grid_bench.sh <tag> <N>
TAG=${1:-r3n8}; N=${2:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 4 --no-cpu-baseline --lean --extra none > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "rc=$?"; python - <<P
import json
for l in open('gpurun_out/bench_${TAG}.json'):
    if l.startswith('{'):
        j=json.loads(l); print('QPS', round(j['value']), 'ms', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), j['details'].get('grid')); print(j['stages_ms']); print(j['roofline']['frac'], j['roofline']['ms_per_launch'])
P
tail -3 gpurun_out/bench_${TAG}.err | grep -v "^\*\|OMP_NUM\|^$" || true
