mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_s4_8gpu.json 2> gpurun_out/bench_s4_8gpu.err
tail -c 300 gpurun_out/bench_s4_8gpu.json
