"""Device-side backward of NetMon / DQN (csrc/train.cu behind torch.autograd) against torch autograd of
the reference's own math (model.py composed from torch ops, `_forward_autograd` / nn.Linear chains) on identical
weights and inputs.

Stated tolerance: every parameter gradient and the state / input gradients within rtol 1e-4 of torch autograd
(relative to the gradient tensor's max magnitude) for fp32 forward arithmetic; 5e-3 when the forward GEMMs run the
tcgen05 bf16x3 split (the backward GEMMs are fp32 in both cases; a pre-activation within ~1e-6 of zero can land on the
other side of leaky_relu's kink, which changes that unit's gradient by a finite step -- measured 2.3e-3 at worst).
"""
import copy
import time

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("math,tol", [("fp32", 1e-4), ("bf16x3", 5e-3)])
@pytest.mark.parametrize("rows,D,hidden", [(640, 642, (512, 256)), (37, 13, (8,)), (300, 130, (64, 48, 32))])
def test_dqn_backward_matches_torch_autograd(rows, D, hidden, math, tol):
    import graph_marl_b200.model as M

    torch.manual_seed(1)
    dqn = M.DQN(D, hidden, 4, F.leaky_relu, math=math).cuda()
    ref = copy.deepcopy(dqn)
    x = torch.randn(rows // 20 or 1, 20 if rows >= 20 else rows, D, device="cuda")
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    wgt = torch.randn(*x.shape[:-1], 4, device="cuda")
    n0 = M.GRAD_PATH_CALLS["device"]
    q = dqn(x1, None)
    assert M.GRAD_PATH_CALLS["device"] == n0 + 1
    q_ref = ref.q_net(ref.encoder(x2))
    assert _rel(q, q_ref) < tol
    (q * wgt).sum().backward()
    (q_ref * wgt).sum().backward()
    assert _rel(x1.grad, x2.grad) < tol
    for (n, p), (_, r) in zip(dqn.named_parameters(), ref.named_parameters()):
        assert p.grad is not None and _rel(p.grad, r.grad) < tol, n


def _netmon_pair(Dn, H, enc, K, agg, nbr, math, rnn="lstm"):
    import graph_marl_b200.model as M

    torch.manual_seed(2)
    nm = M.NetMon(Dn, H, enc, K, F.leaky_relu, rnn_type=rnn, agg_type=agg, output_neighbor_hidden=nbr, math=math).cuda()
    if rnn == "lnlstm":  # LayerNorm parameters away from their (1, 0) initial values
        with torch.no_grad():
            for cell in (nm.rnn_obs, nm.rnn_update):
                for ln in (cell.ln_input, cell.ln_hidden, cell.ln_cell):
                    ln.weight.add_(0.2 * torch.randn_like(ln.weight))
                    ln.bias.add_(0.1 * torch.randn_like(ln.bias))
    return nm, copy.deepcopy(nm)


@pytest.mark.parametrize("math,tol", [("fp32", 1e-4), ("bf16x3", 5e-3)])
@pytest.mark.parametrize("H,enc,K,agg,nbr,steps,rnn", [(128, (512, 256), 2, "sum", True, 3, "lstm"), (32, (48,), 1, "mean", True, 2, "lstm"),
                                                       (16, (24, 8), 3, "sum", False, 2, "lstm"),
                                                       (128, (512, 256), 2, "sum", True, 3, "lnlstm"), (32, (48,), 2, "mean", True, 2, "lnlstm")])
def test_netmon_sequence_backward_matches_torch_autograd(H, enc, K, agg, nbr, steps, rnn, math, tol):
    """BASELINE config 5 inputs (sl.py: 32 evaluation graphs, identity node-agent matrix), a `steps`-long sequence with
    the NetMon state carried (and masked, main.py:861-864) from step to step: outputs, final state and every parameter
    gradient of the device path equal torch autograd of the composed reference math."""
    import graph_marl_b200.model as M

    g = load_golden("sl_netmon")
    x = torch.from_numpy(g["node_obs"][:12]).cuda()
    adj = torch.from_numpy(g["node_adj"][:12]).float().cuda()
    B, N, Dn = x.shape
    nam = (torch.rand(B, N, 7, device="cuda") < 0.2).float()  # an arbitrary (not one-hot) node-agent matrix
    nm, ref = _netmon_pair(Dn, H, enc, K, agg, nbr, math, rnn)
    if rnn == "lnlstm":  # LayerNorm over the gate rows amplifies rounding: looser bars
        tol = 2e-3 if math == "fp32" else 3e-2
    torch.manual_seed(4)
    s0 = torch.randn(B, N, 2 * H, device="cuda") * 0.3
    keep = (torch.rand(B, device="cuda") > 0.3).float().view(-1, 1, 1)
    wts = [torch.randn(B, 7, nm.get_out_features(), device="cuda") for _ in range(steps)]
    ws = torch.randn(B, N, 2 * H, device="cuda")

    def run(mod, fwd):
        mod.state = s0.clone()
        loss = 0.0
        outs = []
        for t in range(steps):
            if t:
                mod.state = mod.state * keep
            xt = x + 0.01 * t
            out = fwd(mod, xt)
            outs.append(out)
            loss = loss + (out * wts[t]).sum()
        loss = loss + (mod.state * ws).sum()
        loss.backward()
        return outs, mod.state

    n0 = M.GRAD_PATH_CALLS["device"]
    o1, st1 = run(nm, lambda m, xt: m(xt, adj, nam))
    assert M.GRAD_PATH_CALLS["device"] == n0 + steps
    o2, st2 = run(ref, lambda m, xt: m._forward_autograd(xt, adj, nam, None, False))
    for a, b in zip(o1, o2):
        assert a.shape == b.shape and _rel(a, b) < tol
    assert _rel(st1, st2) < tol
    for (n, p), (_, r) in zip(nm.named_parameters(), ref.named_parameters()):
        assert p.grad is not None and r.grad is not None, n
        assert _rel(p.grad, r.grad) < tol, (n, _rel(p.grad, r.grad))


def test_first_step_without_state_and_unsupported_cells_take_the_torch_path():
    import warnings

    import graph_marl_b200.model as M

    g = load_golden("sl_netmon")
    x = torch.from_numpy(g["node_obs"][:4]).cuda()
    adj = torch.from_numpy(g["node_adj"][:4]).float().cuda()
    eye = torch.eye(20, device="cuda").repeat(4, 1, 1)
    nm, ref = _netmon_pair(x.shape[-1], 32, (40,), 2, "sum", True, "fp32")
    nm.state = ref.state = None  # zeros (model.py:480-484): no state gradient, W_hh of rnn_obs gets a zero gradient
    nm(x, adj, eye).square().sum().backward()
    ref._forward_autograd(x, adj, eye, None, False).square().sum().backward()
    for (n, p), (_, r) in zip(nm.named_parameters(), ref.named_parameters()):
        assert _rel(p.grad, r.grad) < 1e-4, n
    assert float(nm.rnn_obs.weight_hh.grad.abs().max()) == 0.0
    ln = M.NetMon(x.shape[-1], 32, (40,), 1, F.leaky_relu, rnn_type="gru", output_neighbor_hidden=True).cuda()
    n_t = M.GRAD_PATH_CALLS["torch"]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        M._warned_torch_path.clear()
        ln(x, adj, eye).sum().backward()
    assert M.GRAD_PATH_CALLS["torch"] == n_t + 1 and any("torch-composed" in str(i.message) for i in w)


@pytest.mark.parametrize("reps", [1, 16])
def test_sl_shape_training_step_timed(capsys, reps):
    """sl.py-shape training step (32 graphs x 20 nodes -- and the same batch tiled to 512 graphs --, paper dims, K = 2,
    sequence of 8 NetMon steps, loss on every node output, forward + backward): device path vs the torch-composed path,
    same gradients; both times are printed."""
    g = load_golden("sl_netmon")
    x = torch.from_numpy(g["node_obs"]).cuda().repeat(reps, 1, 1)
    adj = torch.from_numpy(g["node_adj"]).float().cuda().repeat(reps, 1, 1)
    B, N, Dn = x.shape
    eye = torch.eye(N, device="cuda").repeat(B, 1, 1)
    nm, ref = _netmon_pair(Dn, 128, (512, 256), 2, "sum", True, "fp32")

    def step(mod, fwd):
        mod.zero_grad(set_to_none=True)
        mod.state = None
        loss = 0.0
        for t in range(8):
            loss = loss + fwd(mod, x).square().mean()
        loss.backward()
        return loss

    times = {}
    for name, mod, fwd in (("device", nm, lambda m, xt: m(xt, adj, eye)),
                           ("torch", ref, lambda m, xt: m._forward_autograd(xt, adj, eye, None, False))):
        for _ in range(3):
            step(mod, fwd)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            loss = step(mod, fwd)
        torch.cuda.synchronize()
        times[name] = (time.perf_counter() - t0) / 10 * 1e3
    for (n, p), (_, r) in zip(nm.named_parameters(), ref.named_parameters()):
        assert _rel(p.grad, r.grad) < 2e-4, n
    with capsys.disabled():
        print(f"\n[sl-shape training step, {B} graphs = {B * N} node rows x 8 steps] device backward {times['device']:.2f} ms, torch-composed {times['torch']:.2f} ms")
