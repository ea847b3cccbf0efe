#!/bin/bash
# Round-2 GPU call: pytest -m gpu, then a bench line (tag = $1).  Outputs -> gpurun_out/.
TAG=${1:-a}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
