timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_routing.py -m gpu -x -q 2>&1 | tail -5
for i in 1 2; do
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); e=j['e2e']; s=j['stage_ms']; print('value', round(j['value']), 'ms', round(j['ms_per_step'],4), 'e2e', round(e['value']), 'env_ms', round(s['env_kernel_ms'],4), 'env frac', round(j['roofline_env_step']['frac'],3), 'replay', s['replay_kernel_ms'], s['replay_kernel_launches_per_step'], 'launches', j['gpu_launches'], 'mhz', j['clocks']['sm_mhz'])"
done
