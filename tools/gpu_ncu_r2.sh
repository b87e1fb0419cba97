#!/bin/bash
# Round-2 ncu captures on one B200 (each only after the same command exited 0 without ncu).  usage: tools/gpu_ncu_r2.sh <tag>
#   1. launch list of one warmed-up 10 000-query step      -> gpurun_out/launches_<tag>.csv
#   2. --set full of the two scan_tc_kernel launches of it  -> gpurun_out/prof_<tag>.ncu-rep
#   3. launch list + --set full of the 128-query step (the HBM-bound operating point of bench.py's roofline_hbm)
#   4. --set full of coarse quantization at the configs[4] shape (nlist = 65 536, D = 96): exact FP32 kernel and tensor-core filter
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --lean --nprobe 8 --profile-window"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:scan_tc_kernel' -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "10k-query captures rc=$?"
CMD128="$CMD --profile-nq 128"
timeout 600 $CMD128 > gpurun_out/plain128_$TAG.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches128_$TAG.csv $CMD128 > gpurun_out/ncu_launches128_$TAG.log 2>&1 && \
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:scan_tc_kernel' -c 2 -o gpurun_out/prof128_$TAG $CMD128 > gpurun_out/ncu_full128_$TAG.log 2>&1
echo "128-query captures rc=$?"
CMDC="python tools/coarse_ncu.py"
timeout 600 $CMDC > gpurun_out/plainc_$TAG.log 2>&1 && \
timeout 1500 ncu --profile-from-start off --set full --clock-control none -k 'regex:coarse_dist_kernel|scan_tc_kernel|select_topk_kernel|select_small_kernel|finalize_kernel' -c 8 -o gpurun_out/profc_$TAG $CMDC > gpurun_out/ncu_fullc_$TAG.log 2>&1
echo "coarse captures rc=$?"; tail -2 gpurun_out/ncu_fullc_$TAG.log; ls -la gpurun_out | grep $TAG
