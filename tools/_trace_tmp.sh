mkdir -p gpurun_out
bash tools/ncu_kernel.sh 'linear_tc_kernel.*128.*3.*3.*4' 5 gpurun_out/ln_cell_after -- python tools/tc_trace.py cfg2ln
