"""Write-only, read-only and copy bandwidth of this GPU's HBM (torch kernels, CUDA events, best of 10): the env-step and
readout kernels mostly WRITE, and the copy figure in MEASURED_PEAKS.json counts read + write bytes of a copy."""
import json
import torch

n = 1 << 30  # 1 Gi bf16 elements = 2 GiB
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")


def best(fn, nbytes, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return nbytes / (min(t) * 1e-3) / 1e9


out = dict(write_only_gbs=best(lambda: a.fill_(1.0), 2 * n), read_only_gbs=best(lambda: a.view(torch.int32).max(), 2 * n),
           copy_gbs=best(lambda: b.copy_(a), 4 * n),
           how="torch fill_ / max / copy_ over 1 Gi bf16 elements (2 GiB), best of 10, CUDA events")
print(json.dumps(out))
