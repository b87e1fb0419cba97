#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/pytest3.log; tail -25 gpurun_out/pytest3.log
timeout 600 python bench.py > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"; tail -c 4500 gpurun_out/bench3.json; tail -5 gpurun_out/bench3.err
