#!/bin/bash
# One-GPU closing visit: every GPU test, smoke, the bench line, the 10M and 100M configs on one GPU, the reference-harness sweep.
TAG=${1:-fin}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/pytest_$TAG.log; tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/benchref_$TAG.json 2> gpurun_out/benchref_$TAG.err; echo "ref rc=$?"
timeout 400 python bench.py --config c4 --extra none --lean --no-cpu-baseline --steps 10 > gpurun_out/c4_n1_$TAG.json 2> gpurun_out/c4_n1_$TAG.err; echo "c4 rc=$?"
timeout 600 python bench.py --config c5 --extra none --lean --no-cpu-baseline --steps 5 > gpurun_out/c5_n1_$TAG.json 2> gpurun_out/c5_n1_$TAG.err; echo "c5 rc=$?"
timeout 300 python -m vector_indexer_py.bench_harness --output-dir gpurun_out/harness_$TAG > gpurun_out/harness_$TAG.log 2>&1; tail -9 gpurun_out/harness_$TAG.log
