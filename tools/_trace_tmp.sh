timeout 900 python -m pytest tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); e=j['e2e']; print('value', round(j['value']), 'ms', round(j['ms_per_step'],4), 'e2e', round(e['value']), 'ms', round(e['ms_per_step'],4), 'ratio', round(e['value']/j['value'],4))"
done
