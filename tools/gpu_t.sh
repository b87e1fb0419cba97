#!/bin/bash
# run selected GPU tests a few times.  usage: tools/gpu_t.sh "<pytest -k expr>" [repeats]
K="$1"; N=${2:-3}
mkdir -p gpurun_out
for i in $(seq 1 $N); do
  timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider -k "$K" 2>&1 | tail -12 > gpurun_out/t_$i.log; tail -4 gpurun_out/t_$i.log
done
