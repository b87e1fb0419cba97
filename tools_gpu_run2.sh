#!/bin/bash
# scratch driver for one gpurun call: bench + ncu evidence
mkdir -p gpurun_out
set -o pipefail
timeout 800 python bench.py --full-curve > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
NP=$(python -c "import json;print(json.load(open('gpurun_out/bench_full.json'))['config']['nprobe'])")
echo "headline nprobe=$NP"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nprobe $NP"
KRE='regex:coarse_dist|select_topk|group_|scan_|merge_slots|exclusive|fill_u32|scan_apply|scan_of|scan_block|pad_'
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:scan_dense|scan_sparse|coarse_dist' -s 30 -c 9 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
