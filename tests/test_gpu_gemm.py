"""GPU parity of the fused linear layer (gm_linear through the C ABI) in its three arithmetic
modes against an fp64 matmul: fp32 FFMA, bf16x3 (tcgen05, fp32-accurate split) and bf16 (tcgen05,
single pass).  Stated tolerances (max abs error / max |C|): fp32 2e-6, bf16x3 2e-5, bf16 2e-2."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-6, "bf16x3": 2e-5, "bf16": 2e-2}
ACT = {None: -1, "leaky_relu": 0}


def _linear(A, W, bias, act, math, accumulate_into=None):
    import graph_marl_b200._lib as L

    M, K = A.shape
    N = W.shape[0]
    mm = L.MATH_MODES[math]
    C = torch.empty((M, N), device="cuda") if accumulate_into is None else accumulate_into
    nbytes = max(int(L.lib().gm_linear_workspace_bytes(M, N, K, mm)), 16)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    L.check(L.lib().gm_linear(A.data_ptr(), A.stride(0), W.data_ptr(), L.ptr(bias), C.data_ptr(), C.stride(0), M, N, K,
                              ACT[act], mm, ws.data_ptr(), ws.numel(), L.current_stream()))
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize("math", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1000, 512, 88), (257, 256, 512), (300, 512, 642), (4096, 128, 256),
                                   (77, 4, 256), (20000, 512, 128)])
def test_linear_modes_against_fp64(M, N, K, math):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn((M, K), generator=g) * (torch.rand((M, K), generator=g) < 0.7)).cuda()
    W = (torch.randn((N, K), generator=g) / np.sqrt(K)).cuda()
    b = torch.randn((N,), generator=g).cuda()
    ref = torch.nn.functional.leaky_relu(A.double() @ W.double().T + b.double(), 0.01)
    out = _linear(A, W, b, "leaky_relu", math)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < TOL[math], (math, M, N, K, err)
    if math == "bf16x3":  # the split must be far better than a single bf16 pass
        assert err < 1e-4


@pytest.mark.parametrize("math", ["fp32", "bf16x3"])
def test_linear_no_bias_identity_and_strided_views(math):
    """No bias / no activation, A a column slice of a wider matrix (row stride != K, 8-byte aligned only)."""
    g = torch.Generator().manual_seed(5)
    big = torch.randn((500, 642), generator=g).cuda()
    A = big[:, :130]
    W = (torch.randn((512, 130), generator=g) / 12).cuda()
    ref = A.double() @ W.double().T
    out = _linear(A, W, None, None, math)
    assert (out.double() - ref).abs().max().item() / ref.abs().max().item() < TOL[math]
    A2 = big[:, 130:]
    W2 = (torch.randn((512, 512), generator=g) / 22).cuda()
    ref2 = A2.double() @ W2.double().T
    out2 = _linear(A2, W2, None, None, math)
    assert (out2.double() - ref2).abs().max().item() / ref2.abs().max().item() < TOL[math]


def test_linear_exact_small_integers_bf16x3():
    """Small integers are exact in bf16: the tensor-core path must reproduce the integer product
    exactly, which pins the smem core-matrix layout, descriptors and TMEM row/column mapping."""
    g = torch.Generator().manual_seed(11)
    A = torch.randint(-4, 5, (300, 192), generator=g).float().cuda()
    W = torch.randint(-3, 4, (272, 192), generator=g).float().cuda()
    ref = (A.double() @ W.double().T).float()
    for math in ("bf16x3", "bf16"):
        out = _linear(A, W, None, None, math)
        assert torch.equal(out, ref), math
