timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_netmon.py -m gpu -x -q -k "ln or layernorm or LayerNorm or golden" 2>&1 | tail -4
bash tools/_ab.sh build_variants/libbase.so graph_marl_b200/lib/libgraphmarl_b200.so 2 --workload cfg2ln
