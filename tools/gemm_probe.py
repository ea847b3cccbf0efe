"""Times gm_linear (tcgen05 path) over a few shapes: mainloop / epilogue balance probe."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import graph_marl_b200._lib as L

def run(M, N, K, math, iters=10):
    mm = L.MATH_MODES[math]
    A = torch.randn((M, K), device="cuda"); W = torch.randn((N, K), device="cuda") * 0.05; b = torch.zeros(N, device="cuda")
    C = torch.empty((M, N), device="cuda")
    ws = torch.empty(int(L.lib().gm_linear_workspace_bytes(M, N, K, mm)) + 256, dtype=torch.uint8, device="cuda")
    f = lambda: L.check(L.lib().gm_linear(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), C.data_ptr(), N, M, N, K, 0, mm, ws.data_ptr(), ws.numel(), L.current_stream()))
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    passes = 3 if math == "bf16x3" else 1
    print(f"M={M} N={N} K={K} {math}: {ms*1e3:8.1f} us  useful {2*M*N*K/ms/1e9:7.1f} TF/s  tensor-executed {passes*2*M*N*K/ms/1e9:7.1f} TF/s  A+C bytes {(M*K+M*N)*4/ms/1e6:6.0f} GB/s")

for math in ("bf16x3", "bf16"):
    for (M, N, K) in [(81920, 256, 64), (81920, 256, 512), (81920, 256, 2048), (18944, 256, 8192), (81920, 512, 512), (81920, 128, 512), (18944, 512, 8192)]:
        run(M, N, K, math)
