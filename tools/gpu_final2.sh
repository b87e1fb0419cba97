#!/bin/bash
# Closing visit of round 2 on one GPU: every GPU test, smoke, the bench line and the reference arm, the reference-harness sweep,
# then the ncu captures of the final build (launch lists + --set full of the scan, 10 000- and 128-query steps).
TAG=${1:-fin2}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/pytest_$TAG.log; tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/benchref_$TAG.json 2> gpurun_out/benchref_$TAG.err; echo "ref rc=$?"
timeout 300 python -m vector_indexer_py.bench_harness --output-dir gpurun_out/harness_$TAG > gpurun_out/harness_$TAG.log 2>&1; tail -9 gpurun_out/harness_$TAG.log
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --lean --nprobe 8 --profile-window"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:scan_tc_kernel' -s 2 -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "10k-query captures rc=$?"
CMD128="$CMD --profile-nq 128"
timeout 600 $CMD128 > gpurun_out/plain128_$TAG.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches128_$TAG.csv $CMD128 > gpurun_out/ncu_launches128_$TAG.log 2>&1 && \
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:scan_tc_kernel' -c 2 -o gpurun_out/prof128_$TAG $CMD128 > gpurun_out/ncu_full128_$TAG.log 2>&1
echo "128-query captures rc=$?"; ls -la gpurun_out | grep $TAG
