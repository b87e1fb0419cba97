#!/bin/bash
# ncu captures of one warmed-up step (launch list + full capture of the scan kernel).  usage: tools/gpu_ncu.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --nprobe 8 --profile-window"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1 && \
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:scan_tc_kernel' -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log; ls -la gpurun_out | tail -5
