#!/bin/bash
# Epilogue-side norms (VIDX_TC_NB): parity tests of the search paths, then the bench line with and without.
TAG=${1:-r3c}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_multi.py tests/test_gpu_build.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_$TAG.log; tail -6 gpurun_out/pytest_$TAG.log
for NB in 1 0; do
VIDX_TC_NB=$NB timeout 400 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --lean --nprobe 8 > gpurun_out/bench_nb${NB}_$TAG.json 2> gpurun_out/bench_nb${NB}_$TAG.err; echo "bench NB=$NB rc=$?"
python - <<P
import json
for l in open('gpurun_out/bench_nb${NB}_$TAG.json'):
    if l.startswith('{'):
        j=json.loads(l); print('QPS',j['value'],'e2e',j['e2e']['value'], 'launches', j['gpu_launches']); print(j['roofline']['ms_per_launch'], j['roofline']['frac'], 'surv', j['roofline']['filter_survivors'], 'hbm', j['roofline_hbm']['frac'], j['roofline_hbm'].get('qps'), j['details'].get('parity')); print(j['stages_ms'])
P
tail -3 gpurun_out/bench_nb${NB}_$TAG.err
done
