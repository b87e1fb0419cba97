#!/bin/bash
# Two GPUs: the NCCL test (graph replay captures the all-gathers), then the bench under torchrun with and without graph replay.
TAG=${1:-r3m}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -p no:cacheprovider -k "two_ranks or one_rank" 2>&1 | tail -6) > gpurun_out/pytest_$TAG.log; tail -3 gpurun_out/pytest_$TAG.log
for G in 1 0; do
echo "== VIDX_GRAPH_MULTI=$G"
VIDX_GRAPH_MULTI=$G timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus 2 --steps 20 --warmup 4 --no-cpu-baseline --lean --extra none > gpurun_out/bench_${TAG}_g$G.json 2> gpurun_out/bench_${TAG}_g$G.err
echo "rc=$?"; python - <<P
import json
for l in open('gpurun_out/bench_${TAG}_g$G.json'):
    if l.startswith('{'):
        j=json.loads(l); print('QPS', round(j['value']), 'ms', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), j['details'].get('grid'), j['details'].get('parity')); print(j['stages_ms'])
P
tail -3 gpurun_out/bench_${TAG}_g$G.err | grep -v "^\*\|OMP_NUM\|^$"
done
