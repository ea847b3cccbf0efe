set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_s3.log
for cfg in "1 256" "0 256" "1 320" "1 160" "1 256"; do
  set -- $cfg
  GM_AGG_STAGE_LISTS=$1 GM_AGG_THREADS=$2 python bench.py --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/aggs_$1_$2.json 2> gpurun_out/aggs_$1_$2.err
done
tail -3 gpurun_out/pytest_s3.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/aggs_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['roofline_aggregate']['ms_per_launch'], d['roofline_aggregate']['frac'])
    except Exception as e: print(f, 'ERR', e)
P
