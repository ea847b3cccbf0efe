for L in build_variants/libtcprobe.so build_variants/libpfprobe.so; do
echo "== LSTM cell $L"; GM_LIB_PATH=$L GM_TC_TRACE_EPI=1 python tools/tc_trace.py 2>&1 | sed -n 5,9p | cut -c1-100
done
bash tools/_ab.sh graph_marl_b200/lib/libgraphmarl_b200.so build_variants/libpf.so 2
