set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_rollout.py tests/test_sl_config5.py -m gpu -x -q > gpurun_out/pytest_var.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_var.log
tail -4 gpurun_out/pytest_var.log
grep -q "rc=0" gpurun_out/pytest_var.log || exit 1
for cfg in 0 1 0 1; do
  GM_AGG_PIPE_GENERIC=$cfg timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/aggs_g${cfg}_$RANDOM.json 2>/dev/null
done
