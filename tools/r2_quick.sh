#!/bin/bash
# quick GPU check: selected tests ($2, default routing + rollout) and a bench line (tag = $1) [+ extra bench args $3]
TAG=${1:-q}
SEL=${2:-"tests/test_gpu_routing.py tests/test_gpu_rollout.py"}
mkdir -p gpurun_out
timeout 1500 python -m pytest $SEL -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -8 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline $3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
try:
    j=json.load(open('gpurun_out/bench_$TAG.json'))
    print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'],'launches',j['gpu_launches'])
    print({k:(round(v,4) if isinstance(v,float) else v) for k,v in j['stage_ms'].items()})
    print('env frac',j['roofline_env_step']['frac'],'agg',j['roofline_aggregate']['frac'],'gemm',j['roofline_gemm']['frac'])
except Exception as e:
    print('no bench line',e)
PY
tail -5 gpurun_out/bench_$TAG.err
