#!/bin/bash
# Where does the epilogue's time go?  Timing ablations (wrong answers) of the main launch on the bench workload, ablation builds.
TAG=${1:-r3d}
mkdir -p gpurun_out
for LIB in ${LIBS:-ablate}; do
export VIDX_B200_LIB=$PWD/vector-indexer_b200/lib_$LIB/libvidx_b200.so
for NB in 0 1; do
echo "== lib_$LIB VIDX_TC_NB=$NB"
VIDX_TC_NB=$NB VARIANTS=0 FLAGS=${FLAGS:-0,4096,8192,2} timeout 300 python tools/ablate.py 2>&1 | tee gpurun_out/ablate_${LIB}_nb${NB}_$TAG.log | tail -5
done
done
