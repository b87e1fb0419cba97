#!/bin/bash
# A variant build of the rare path (lib_$1): parity tests, then timing against the shipped path (ablation builds of both).
V=${1:-xp}; TAG=${2:-r3x}
mkdir -p gpurun_out
(VIDX_B200_LIB=$PWD/vector-indexer_b200/lib_$V/libvidx_b200.so timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -8) > gpurun_out/pytest_$TAG.log; tail -4 gpurun_out/pytest_$TAG.log
LIBS="ablate $V" FLAGS=0,4096 bash tools/gpu_ablate.sh $TAG
