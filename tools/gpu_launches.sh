#!/bin/bash
# launch list of one warmed-up step only.  usage: tools/gpu_launches.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --nprobe 8 --profile-window"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu rc=$?"; grep "scan_tc\|finalize" gpurun_out/launches_$TAG.csv | awk -F'","' '{print $5, $NF}'
