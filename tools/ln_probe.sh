for F in ${LN_FLAGS:-0 2 4 8 16 24 28}; do
  GM_LN_DEBUG=$F GM_LIB_PATH=build_variants/libtcprobe.so timeout 300 python bench.py --workload cfg2ln --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('LN_DEBUG=$F', round(j['ms_per_step'],4), 'gemm_ms', round(j['stage_ms']['gemm_ms'],4))"
done
