#!/bin/bash
# Every GPU test, then the per-rank batch sizes of a replica grid on one GPU with and without graph replay.
TAG=${1:-r3b}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_$TAG.log; tail -6 gpurun_out/pytest_$TAG.log
for G in 1 0; do
echo "== VIDX_GRAPH=$G"
VIDX_GRAPH=$G NQS=10000,5000,2500,1250,625,128 SCAN_MODES=0 timeout 300 python tools/nq_sweep.py 2>&1 | tee gpurun_out/nq_sweep_g${G}_$TAG.log | tail -8
done
