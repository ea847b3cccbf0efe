#!/bin/bash
# side measurements for profiles/: env kernel alone at 4096 / 2048 envs, the other workloads and math modes on one GPU
mkdir -p gpurun_out
for B in 4096 2048; do timeout 300 python tools/env_only.py $B 60 2>&1 | head -1 | tee -a gpurun_out/env_only.txt; done
for W in cfg2ln cfg3 cfg4; do
  timeout 900 python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "$W rc=$?"
done
timeout 900 python bench.py --math bf16 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bf16 rc=$?"
python - <<PY
import json
for t in ("cfg2ln","cfg3","cfg4","bf16"):
    try:
        j=json.load(open(f'gpurun_out/bench_{t}.json'))
        print(t,'value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'],'gemm_ms',j['stage_ms']['gemm_ms'],'envfrac',j['roofline_env_step']['frac'], j['config'].get('envs_per_gpu'))
    except Exception as e:
        print(t,'no line',e)
PY
