#!/bin/bash
# Closing visit of round 2 (second session) on one GPU: every GPU test, smoke, the bench line and the reference arm, the
# reference's bench.yaml grid, then ncu: launch lists of one step with graph replay on (the kernels are graph nodes) and off, the
# --set full capture of the scan of the same step, and the --set full capture of the streamed-query-tile scan at D = 768.
TAG=${1:-f3}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6) > gpurun_out/pytest_$TAG.log; tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/benchref_$TAG.json 2> gpurun_out/benchref_$TAG.err; echo "ref rc=$?"
timeout 600 python tools/bench_yaml_grid.py > gpurun_out/grid_$TAG.jsonl 2> gpurun_out/grid_$TAG.err; echo "grid rc=$?"; tail -2 gpurun_out/grid_$TAG.err
CMD="python bench.py --steps 1 --warmup 4 --no-cpu-baseline --lean --nprobe 8 --profile-window"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph_$TAG.csv $CMD > gpurun_out/ncu_launches_graph_$TAG.log 2>&1; echo "launch list (graph replay) rc=$?"
VIDX_GRAPH=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "launch list rc=$?"
VIDX_GRAPH=0 timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:scan_tc_kernel' -s 2 -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1; echo "scan capture rc=$?"
CMD768="python tools/bench_yaml_grid.py --dims 768 --counts 100000 --nprobes 8 --reps 1"
VIDX_GRAPH=0 timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:scan_tc_kernel|tc_atile_kernel' -c 7 -o gpurun_out/prof768_$TAG $CMD768 > gpurun_out/ncu_full768_$TAG.log 2>&1; echo "D=768 capture rc=$?"
ls -la gpurun_out | grep $TAG
