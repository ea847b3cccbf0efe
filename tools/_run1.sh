set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_routing.py tests/test_gpu_rollout.py -m gpu -x -q > gpurun_out/pytest_env.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_env.log
tail -4 gpurun_out/pytest_env.log
grep -q "rc=0" gpurun_out/pytest_env.log || exit 1
for i in 1 2; do
python bench.py --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/envb_$i.json 2> gpurun_out/envb_$i.err
done
