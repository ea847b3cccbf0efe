#!/bin/bash
# Multi-GPU call of round 2 (run under `gpurun --gpus N`): weak- and strong-scaled config 2, BASELINE config 3,
# and the learner's collectives on NCCL.  Outputs -> gpurun_out/.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29501 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_weak_${N}gpu.json 2> gpurun_out/bench_weak_${N}gpu.err; echo "weak rc=$?"
$TR --master-port 29502 bench.py --gpus $N --steps 30 --warmup 5 --scaling strong > gpurun_out/bench_strong_${N}gpu.json 2> gpurun_out/bench_strong_${N}gpu.err; echo "strong rc=$?"
$TR --master-port 29503 bench.py --gpus $N --steps 50 --warmup 5 --workload cfg3 > gpurun_out/bench_cfg3_${N}gpu.json 2> gpurun_out/bench_cfg3_${N}gpu.err; echo "cfg3 rc=$?"
$TR --master-port 29504 tools/allreduce_check.py > gpurun_out/allreduce_${N}gpu.json 2> gpurun_out/allreduce_${N}gpu.err; echo "allreduce rc=$?"
for f in weak strong cfg3; do python - <<PY
import json
try:
    j=json.load(open('gpurun_out/bench_${f}_${N}gpu.json'))
    print('$f', 'value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], 'envs/gpu', j['config']['envs_per_gpu'], 'scaling', j['scaling'])
except Exception as e:
    print('$f failed', e)
PY
done
cat gpurun_out/allreduce_${N}gpu.json; tail -3 gpurun_out/allreduce_${N}gpu.err
