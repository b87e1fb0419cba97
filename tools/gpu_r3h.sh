#!/bin/bash
# Warp-private hit rings (-DVIDX_TC_WARPQ build): parity tests, then timing against the shipped queue (ablation builds of both).
TAG=${1:-r3h}
mkdir -p gpurun_out
(VIDX_B200_LIB=$PWD/vector-indexer_b200/lib_warpq/libvidx_b200.so timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -8) > gpurun_out/pytest_$TAG.log; tail -4 gpurun_out/pytest_$TAG.log
LIBS="ablate warpq" FLAGS=0,4096 bash tools/gpu_r3d.sh $TAG
