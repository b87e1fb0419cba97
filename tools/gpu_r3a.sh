#!/bin/bash
# Round-2 re-entry, first GPU visit: the streamed-query-tile scan (D > 512) against the oracle, smoke, the D = 768 rows of the
# reference's bench.yaml grid, and a short bench line (did the default kernel keep its time?).  Outputs under gpurun_out/.
TAG=${1:-r3a}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -q -x --timeout 240 -p no:cacheprovider \
    -k "large_dimensions or streamed or coarse_tensor or pipeline_shapes" 2>&1 | tail -25) > gpurun_out/pytest_$TAG.log; tail -8 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 400 python tools/bench_yaml_grid.py --dims 768 --counts 100000 > gpurun_out/grid768_$TAG.jsonl 2> gpurun_out/grid768_$TAG.err; echo "grid rc=$?"; cat gpurun_out/grid768_$TAG.jsonl; tail -3 gpurun_out/grid768_$TAG.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lean --nprobe 8 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<P
import json
for l in open('gpurun_out/bench_$TAG.json'):
    if l.startswith('{'):
        j=json.loads(l); print('QPS',j['value'],'e2e',j['e2e']['value']); print(j['roofline']['ms_per_launch'], j['roofline']['frac']); print(j['stages_ms'])
P
tail -3 gpurun_out/bench_$TAG.err
