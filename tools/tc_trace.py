"""Per-tile timeline of one tcgen05 kernel launch (CTA 0): needs a build with GM_NVCC_EXTRA=-DGM_TC_PROBES=1.
usage: GM_TC_TRACE_EPI=<0 linear|1 lstm|2 qhead|9 fused encoder L1+L2 (columns 3-7 then hold one producer thread's clocks per tile:
gathers, activation + stores, tile end, wait for the chunk, wait for a free A slot)> [GM_TC_TRACE_KP=<packed K of the layer, e.g. 96 encoder L1, 512 encoder L2,
672 DQN L1>] [GM_LIB_PATH=build_variants/libtcprobe.so] python tools/tc_trace.py [cfg]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
os.environ["GM_TC_TRACE_PTR"] = hex(buf.data_ptr())
from graph_marl_b200.rollout import Rollout  # noqa: E402

ro = Rollout(sys.argv[1] if len(sys.argv) > 1 else "cfg2", num_envs=4096, math="bf16x3", with_replay=False)
ro.reset()
for _ in range(3):
    ro.step()
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(64, 8)
t0 = t[0, 0]
names = ["mma:acc_free", "mma:first_full", "mma:issued", "epi:acc_full", "epi:done", "copy:issued", "epi:step0_loaded", "epi:step0_done"]
print("tile " + " ".join(f"{n:>15s}" for n in names) + "   (SM clocks relative to tile 0 acc_free; last traced launch)")
fused = os.environ.get("GM_TC_TRACE_EPI") == "9"  # fused encoder: slots 3, 4, 6, 7 = clocks one producer thread spent in the gathers, in activation + stores, waiting (chunk, A slot)
for i in range(64):
    if t[i, 0] == 0:
        break
    print(f"{i:4d} " + " ".join(f"{int(t[i, k] - (0 if fused and k in (3, 4, 5, 6, 7) else t0)):15d}" for k in range(8)))
