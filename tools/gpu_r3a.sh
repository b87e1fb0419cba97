#!/bin/bash
# Round-2 re-entry, first GPU visit: (1) the streamed-query-tile scan (D > 512) against the oracle with graph replay off,
# (2) graph replay of repeated searches (parity tests that repeat searches, the threads test), (3) smoke, (4) the D = 768 rows
# of the reference's bench.yaml grid, (5) a short bench line with and without graph replay.  Outputs under gpurun_out/.
TAG=${1:-r3a}
mkdir -p gpurun_out
(VIDX_GRAPH=0 timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -q -x --timeout 240 -p no:cacheprovider \
    -k "large_dimensions or streamed or coarse_tensor or pipeline_shapes" 2>&1 | tail -25) > gpurun_out/pytest_sa_$TAG.log; tail -8 gpurun_out/pytest_sa_$TAG.log
(timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_multi.py -m gpu -q --timeout 240 -p no:cacheprovider \
    -k "replayed or repeated or sparse_regime or dense_regime or threads or streams or large_k or overflow or one_rank_communicator or nan_query" 2>&1 | tail -25) > gpurun_out/pytest_graph_$TAG.log; tail -8 gpurun_out/pytest_graph_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 400 python tools/bench_yaml_grid.py --dims 768 --counts 100000 > gpurun_out/grid768_$TAG.jsonl 2> gpurun_out/grid768_$TAG.err; echo "grid rc=$?"; cat gpurun_out/grid768_$TAG.jsonl; tail -3 gpurun_out/grid768_$TAG.err
for G in 1 0; do
VIDX_GRAPH=$G timeout 400 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --lean --nprobe 8 > gpurun_out/bench_g${G}_$TAG.json 2> gpurun_out/bench_g${G}_$TAG.err; echo "bench graph=$G rc=$?"
python - <<P
import json
for l in open('gpurun_out/bench_g${G}_$TAG.json'):
    if l.startswith('{'):
        j=json.loads(l); print('QPS',j['value'],'e2e',j['e2e']['value'], 'launches', j['gpu_launches']); print(j['roofline']['ms_per_launch'], j['roofline']['frac'], 'hbm', j['roofline_hbm']['frac'], j['roofline_hbm'].get('qps')); print(j['stages_ms'])
P
tail -3 gpurun_out/bench_g${G}_$TAG.err
done
