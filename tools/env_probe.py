"""Times the routing env step kernel alone (CUDA events) at BASELINE config-2 size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graph_marl_b200.env.network import Network
from graph_marl_b200.env.routing import Routing

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = A = 20
net = Network(N, random_topology=False, topology_init_seed=923430603)
env = Routing(net, A, 1, num_envs=B, seed=1, batched=True, store_mode=int(os.environ.get("STORE_MODE", "0")))
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
acts = [torch.randint(0, 4, (B, A), device="cuda", generator=g, dtype=torch.int32) for _ in range(8)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(5):
    env.step(acts[i % 8])
ts = []
for i in range(30):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step(acts[i % 8]); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
bytes_step = 20772 * B
print(f"B={B} store_mode={os.environ.get('STORE_MODE','0')}: median {ts[len(ts)//2]*1e3:.1f} us  min {ts[0]*1e3:.1f} us  -> {bytes_step/ts[len(ts)//2]/1e6:.0f} GB/s (L2 flushed between steps; includes python launch overhead)")
