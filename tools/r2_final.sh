#!/bin/bash
# Round-2 closing call on one GPU: the whole `-m gpu` suite, then the measurement files of profiles/ (tools/r2_measure.sh)
# and the other workloads' bench lines.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_final.log
tail -3 gpurun_out/pytest_final.log
bash tools/r2_measure.sh final
for W in cfg2ln cfg3 cfg4; do
  timeout 900 python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "$W rc=$?"
done
for B in 4096 2048; do timeout 300 python tools/env_only.py $B 60 2>&1 | head -1 | tee -a gpurun_out/env_only_final.txt; done
python - <<PY
import json
for t in ("cfg2ln","cfg3","cfg4"):
    try:
        j=json.load(open(f'gpurun_out/bench_{t}.json'))
        print(t,'value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'],'gemm_ms',j['stage_ms']['gemm_ms'],'envfrac',j['roofline_env_step']['frac'], j['config'].get('envs_per_gpu'))
    except Exception as e:
        print(t,'no line',e)
PY
