timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -3
echo "== LSTM"; GM_LIB_PATH=build_variants/libtcprobe.so GM_TC_TRACE_EPI=1 python tools/tc_trace.py 2>&1 | sed -n 5,9p | cut -c1-100
echo "== DQN L1"; GM_LIB_PATH=build_variants/libtcprobe.so GM_TC_TRACE_EPI=0 GM_TC_TRACE_KP=672 python tools/tc_trace.py 2>&1 | sed -n 5,9p | cut -c1-100
echo "== DQN L2"; GM_LIB_PATH=build_variants/libtcprobe.so GM_TC_TRACE_EPI=2 python tools/tc_trace.py 2>&1 | sed -n 3,7p | cut -c1-100
bash tools/_ab.sh build_variants/libprev.so graph_marl_b200/lib/libgraphmarl_b200.so 3
