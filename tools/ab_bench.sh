# A/B of two library builds in one call: tools/ab_bench.sh <libA> <libB> [rounds] [bench args]
A=$1; B=$2; R=${3:-2}; shift 3
for r in $(seq 1 $R); do
for L in $A $B; do
GM_LIB_PATH=$L timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); s=j['stage_ms']; print('$L', 'ms', round(j['ms_per_step'],4), 'value', round(j['value']), 'gemm', round(s['gemm_ms'],4), 'agg', round(s['aggregate_kernel_ms'],4), 'readout', round(s['readout_kernel_ms'],4), 'env', round(s['env_kernel_ms'],4), 'mhz', j['clocks']['sm_mhz'])"
done; done
