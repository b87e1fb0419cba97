#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --nprobe 8"
timeout 600 $CMD > gpurun_out/plain_tc.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:scan_tc_kernel' -s 9 -c 2 -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tc.log; ls -la gpurun_out | tail -4
