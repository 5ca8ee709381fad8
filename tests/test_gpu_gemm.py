"""GPU: the tensor-core projection (3xTF32 on tcgen05) against an fp64 F.linear, within the 1e-5
relative bound of the north star; shapes cover M tails, N not a multiple of 16, K tails, N > 256."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def ctx():
    from gigl_b200 import Context

    c = Context.on_torch_stream(0)
    yield c
    c.close()


@pytest.mark.parametrize("M,N,K", [(128, 256, 200), (1, 16, 8), (1000, 47, 512), (130, 64, 32), (4097, 256, 200),
                                   (300, 300, 70), (77, 7, 5), (50000, 128, 256), (257, 512, 96)])
@pytest.mark.parametrize("relu,bias", [(False, True), (True, False)])
def test_linear_tc_matches_fp64(ctx, M, N, K, relu, bias):
    import torch

    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    out = ctx.linear(A, W, b, relu=relu)
    ctx.sync()
    ref = A.double() @ W.double().T
    if bias:
        ref = ref + b.double()
    if relu:
        ref = ref.clamp(min=0)
    err = float((out.double() - ref).abs().max() / max(1.0, float(ref.abs().max())))
    assert err < RTOL, err
    # per-element relative check away from cancellation: 3xTF32 + fp32 tensor-core accumulation stays ~1e-5, 1xTF32 would be ~1e-3
    big = ref.abs() > 0.5
    if bool(big.any()):
        assert float(((out.double() - ref).abs() / ref.abs())[big].max()) < 2e-4


def test_linear_tc_strided_inputs_and_identity(ctx):
    import torch

    g = torch.Generator(device="cuda").manual_seed(0)
    big = torch.randn(500, 300, device="cuda", generator=g)
    A = big[:, 10:110]  # pitch 300, offset 10 floats (not 16-byte aligned: the split pass re-pitches)
    W = torch.eye(100, device="cuda")
    out = ctx.linear(A, W)
    ctx.sync()
    # hi + lo reassembles the value up to the TF32 rounding of lo (2^-21 relative)
    assert float(((out - A).abs() / A.abs().clamp(min=1e-30)).max()) < 2e-6
    # huge / tiny magnitudes keep their relative accuracy
    A2 = torch.randn(256, 64, device="cuda", generator=g) * 1e6
    W2 = torch.randn(32, 64, device="cuda", generator=g) * 1e-6
    o2 = ctx.linear(A2, W2)
    ref = A2.double() @ W2.double().T
    assert float((o2.double() - ref).abs().max() / ref.abs().max()) < RTOL
