#!/bin/bash
# One-GPU measurement call of round 2 (run under gpurun): bench line, reference-arm line, ncu launch list of the same
# program kernel by kernel (--graph-steps 0) and ONE `ncu --set full` pass over one rollout step's kernels.
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --graph-steps 0 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --graph-steps 0 > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:linear_tc|enc_fused|routing_kernel|aggregate_pk|replay_insert|readout_agents" -s 300 -c 32 -f -o gpurun_out/step_$TAG \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --graph-steps 0 > gpurun_out/ncu_step_$TAG.log 2>&1
ncu -i gpurun_out/step_$TAG.ncu-rep --page raw --csv > gpurun_out/step_${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/step_${TAG}_raw.csv gpurun_out/step_${TAG}_summary.json gpurun_out/ncu_traffic_$TAG.json | tail -40
rm -f gpurun_out/step_$TAG.ncu-rep
tail -c 400 gpurun_out/bench_$TAG.json; echo
