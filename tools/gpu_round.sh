#!/bin/bash
# One GPU-box visit: parity tests, the bench line, then (only if both exited 0) the ncu launch list of one
# timed step and one full capture of the scan kernel.  Outputs under gpurun_out/.
# usage: tools/gpu_round.sh <tag> [--no-ncu]
TAG=${1:-x}; NONCU=$2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; tail -5 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
[ "$NONCU" = "--no-ncu" ] && exit 0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --nprobe 8 --profile-window"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1 && \
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:scan_tc_kernel' -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log; ls -la gpurun_out | tail -5
