#!/bin/bash
# quick GPU check: smoke, parity tests, short bench.  usage: tools/gpu_quick.sh <tag> [bench args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; tail -8 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; python - <<P
import json
for l in open('gpurun_out/bench_$TAG.json'):
    if l.startswith('{'):
        j=json.loads(l); print('QPS',j['value'],'e2e',j['e2e']['value'],'recall',j['config']['recall_at_10']); print(j['roofline']); print(j['roofline_hbm']); print(j['stages_ms'])
P
tail -5 gpurun_out/bench_$TAG.err
