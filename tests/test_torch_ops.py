"""The torch.library binding (gigl_b200/csrc/torch_ops.cpp -> lib/libgigl_b200_torch.so): SURVEY.md 8(b) row B4 asks for
`torch.ops.gigl_b200.*` from a C++ extension so that DDP (training_process.py:298-303) and graph capture see dispatcher
ops.  CPU: the library loads, the ops exist with their schemas, shapes propagate through the Meta kernels, CPU tensors are
refused (no fallback).  GPU: the ops are what gigl_b200.nn runs, opcheck accepts their registrations, gradients flow."""
import numpy as np
import pytest


def test_ops_are_registered_and_refuse_cpu_tensors():
    import torch

    import gigl_b200.nn  # noqa: F401  (loads the extension)

    ops = torch.ops.gigl_b200
    for name in ("csr_from_coo", "sage_conv", "gcn_conv", "sage_conv_fwd", "sage_conv_bwd", "gcn_conv_fwd", "gcn_conv_bwd"):
        assert hasattr(ops, name), name
    assert "Tensor? bl" in str(ops.sage_conv.default._schema) and "int m" in str(ops.sage_conv.default._schema)
    # Meta kernels: shapes without a device
    x = torch.empty(10, 6, device="meta")
    rp, col = torch.empty(11, dtype=torch.int64, device="meta"), torch.empty(20, dtype=torch.int32, device="meta")
    Wl, Wr = torch.empty(4, 6, device="meta"), torch.empty(4, 6, device="meta")
    out, saved = ops.sage_conv_fwd(x, rp, col, Wl, None, Wr, False, 7)
    assert out.shape == (7, 4) and saved.shape == (7, 16)  # saved = [mean | self] with F padded to a multiple of 4
    assert ops.gcn_conv_fwd(x, rp, col, Wl, None, True).shape == (10, 4)
    r2, c2 = ops.csr_from_coo(torch.empty(20, dtype=torch.int64, device="meta"), torch.empty(20, dtype=torch.int64, device="meta"), 10)
    assert r2.shape == (11,) and c2.dtype == torch.int32
    # no CPU implementation: the dispatcher refuses, nothing falls back
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.sage_conv_fwd(torch.zeros(10, 6), torch.zeros(11, dtype=torch.int64), torch.zeros(20, dtype=torch.int32), torch.zeros(4, 6), None,
                          torch.zeros(4, 6), False, 7)
    m = gigl_b200.nn.GraphSAGE(6, 8, 2, 4)
    with pytest.raises(RuntimeError):
        m(torch.zeros(5, 6), torch.zeros(2, 3, dtype=torch.int64))


@pytest.mark.gpu
def test_ops_run_the_layers_and_pass_opcheck():
    import torch

    import gigl_b200.nn as gnn
    from oracle import oracle as O

    ops = torch.ops.gigl_b200
    rng = np.random.default_rng(0)
    n, e, F, Oc = 300, 2500, 20, 12
    ei = torch.from_numpy(np.stack([rng.integers(0, n, e), rng.integers(0, n, e)])).cuda()
    x = torch.from_numpy(rng.standard_normal((n, F)).astype(np.float32)).cuda().requires_grad_(True)
    conv = gnn.SAGEConv(F, Oc).cuda()
    gi = gnn.GraphIndex(ei, n)
    out_mod = conv(x, gi, relu=True)
    t_rowptr, t_col = gi.transposed
    out_op = ops.sage_conv(x, gi.rowptr, gi.col, t_rowptr, t_col, conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight, True, n)
    assert torch.equal(out_mod, out_op)  # the module IS the op
    ref = O.c_sage_conv(x.detach().cpu().numpy(), ei.cpu().numpy(), conv.lin_l.weight.detach().cpu().numpy(),
                        conv.lin_l.bias.detach().cpu().numpy(), conv.lin_r.weight.detach().cpu().numpy(), relu=True, f64=True)
    assert np.abs(out_op.detach().cpu().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    out_op.sum().backward()
    assert x.grad is not None and conv.lin_l.weight.grad is not None and conv.lin_r.weight.grad is not None and conv.lin_l.bias.grad is not None
    args = (x.detach(), gi.rowptr, gi.col, conv.lin_l.weight.detach(), conv.lin_l.bias.detach(), conv.lin_r.weight.detach(), True, n)
    torch.library.opcheck(ops.sage_conv_fwd.default, args, test_utils=("test_schema", "test_faketensor"))
    gcn = gnn.GCNConv(F, Oc).cuda()
    y = gcn(x.detach().requires_grad_(True), gi, relu=False)
    y.square().mean().backward()
    assert gcn.lin.weight.grad is not None and gcn.bias.grad is not None
