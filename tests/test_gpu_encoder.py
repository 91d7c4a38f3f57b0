"""Encoder conv5 + max-pool fused kernel (tcgen05) against a plain PyTorch fp32 reference of the
same op.  Operands are rounded to bf16 before the tensor cores (fp32 accumulation): against an fp32
reference fed the SAME bf16-rounded operands the kernel must agree to fp32 summation-order noise
(1e-4 of the output scale); against the un-rounded fp32 layer the tolerance is bf16's (2e-2)."""
import numpy as np
import pytest
import torch

from pointnet_autoencoder_b200 import ops
from pointnet_autoencoder_b200.encoder import PointNetEncoder

pytestmark = pytest.mark.gpu


def scaled_err(a, ref):
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("b,n,k,c", [(2, 256, 128, 1024), (3, 2048, 128, 1024), (2, 300, 128, 256), (1, 1000, 64, 128),
                                     (4, 37, 128, 128)])
def test_conv_pool_stats_vs_torch(b, n, k, c):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(b, n, k, device="cuda", generator=g)
    w = torch.randn(k, c, device="cuda", generator=g) / k ** 0.5
    xb = x.to(torch.bfloat16); wb = w.to(torch.bfloat16)
    vmax, vmin, vsum, vsq = ops.encoder_conv_pool(xb, wb.t().contiguous())
    y = xb.float() @ wb.float()                      # fp32 reference on the same bf16-rounded operands
    assert scaled_err(vmax, y.amax(1)) < 1e-4
    assert scaled_err(vmin, y.amin(1)) < 1e-4
    assert scaled_err(vsum, y.sum(1)) < 1e-4 * n ** 0.5
    assert scaled_err(vsq, (y * y).sum(1)) < 1e-4
    y32 = x @ w                                      # the un-rounded fp32 layer: bf16 operand rounding only
    assert scaled_err(vmax, y32.amax(1)) < 2e-2


@pytest.mark.parametrize("training", [True, False])
def test_encoder_fused_matches_unfused(training):
    torch.manual_seed(0)
    enc = PointNetEncoder(fused=True).cuda()
    ref = PointNetEncoder(fused=False).cuda()
    ref.load_state_dict(enc.state_dict())
    with torch.no_grad():       # non-trivial BN parameters, including negative scales (the min path)
        enc.conv5.gamma.copy_(torch.randn(1024, device="cuda")); enc.conv5.beta.copy_(0.1 * torch.randn(1024, device="cuda"))
        enc.conv5.moving_var.copy_(torch.rand(1024, device="cuda") + 0.5)
        ref.load_state_dict(enc.state_dict())
    enc.train(training); ref.train(training)
    pc = torch.randn(4, 1024, 3, device="cuda")
    a = enc(pc); r = ref(pc)
    assert a.shape == (4, 1024)
    assert scaled_err(a, r) < 2e-2                   # bf16 operands vs the fp32 library path
    if training:                                     # moving statistics were updated the same way
        assert scaled_err(enc.conv5.moving_mean, ref.conv5.moving_mean) < 2e-2
        assert scaled_err(enc.conv5.moving_var, ref.conv5.moving_var) < 2e-2


def test_encoder_backward_runs_and_matches_unfused():
    torch.manual_seed(0)
    enc = PointNetEncoder(fused=True).cuda().train()
    ref = PointNetEncoder(fused=False).cuda().train()
    ref.load_state_dict(enc.state_dict())
    pc = torch.randn(2, 512, 3, device="cuda")
    enc(pc).square().sum().backward()
    ref(pc).square().sum().backward()
    ga = enc.conv5.weight.grad; gr = ref.conv5.weight.grad
    assert scaled_err(ga, gr) < 5e-2
    assert scaled_err(enc.layers[0].weight.grad, ref.layers[0].weight.grad) < 5e-2
