set -x
mkdir -p gpurun_out
for m in 2 1; do
GM_AGG_MAP=$m ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:aggregate_pk" -s 30 -c 2 -f -o gpurun_out/agg_map$m \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --graph-steps 0 > gpurun_out/ncu_agg_map$m.log 2>&1
tail -3 gpurun_out/ncu_agg_map$m.log
done
ls -la gpurun_out/agg_map*
