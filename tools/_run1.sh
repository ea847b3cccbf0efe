set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "variants" > gpurun_out/pytest_var.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_var.log
for cfg in "4 160 0" "4 256 0" "8 160 0" "4 160 80" "4 320 80" "4 128 0" "4 160 0" "8 160 0"; do
  set -- $cfg
  GM_AGG_MAP=$1 GM_AGG_THREADS=$2 GM_AGG_ROWS=$3 python bench.py --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/aggm_$1_$2_$3.json 2> gpurun_out/aggm_$1_$2_$3.err
done
tail -5 gpurun_out/pytest_var.log
