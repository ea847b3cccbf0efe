#!/bin/bash
# One-GPU round-end measurement (run under gpurun): bench line, reference-arm line, ncu launch list of the same program
# kernel by kernel (--graph-steps 0) and one `ncu --set full` pass over the HBM-bound kernels.  Outputs -> gpurun_out/.
TAG=${1:-final}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --graph-steps 0 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --graph-steps 0 > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:routing_kernel|aggregate_pk" -s 8 -c 8 -f -o gpurun_out/envagg_$TAG \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --graph-steps 0 > gpurun_out/ncu_envagg_$TAG.log 2>&1
ncu -i gpurun_out/envagg_$TAG.ncu-rep --page raw --csv > gpurun_out/envagg_${TAG}_raw.csv 2>/dev/null
tail -c 600 gpurun_out/bench_$TAG.json; echo; ls -la gpurun_out/*$TAG*
