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


def test_conv_pool_argext():
    g = torch.Generator(device="cuda").manual_seed(2)
    b, n, k, c = 3, 700, 128, 256
    x = torch.randn(b, n, k, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(k, c, device="cuda", generator=g) / k ** 0.5).to(torch.bfloat16)
    sign = torch.randn(c, device="cuda", generator=g)
    vmax, vmin, _, _, arg = ops.encoder_conv_pool(x, w.t().contiguous(), sign=sign)
    assert arg.dtype == torch.int32 and arg.shape == (b, c) and int(arg.min()) >= 0 and int(arg.max()) < n
    y = x.float() @ w.float()
    # the reported point really attains the reported extremum (to accumulation-order noise)
    picked = y.gather(1, arg.long().unsqueeze(1)).squeeze(1)
    want = torch.where(sign >= 0, vmax, vmin)
    assert float((picked - want).abs().max()) < 1e-4 * float(y.abs().max())


@pytest.mark.parametrize("training", [True, False])
def test_conv5_pool_backward_vs_autograd(training):
    """The structured backward (Gram-matrix form, no (B,N,C) tensor) against autograd through plain fp32
    library ops.  Operands are pre-rounded to bf16 so both paths see the same values and select the same
    arg-max points (with un-rounded operands bf16 can flip a near-tie arg-max, a legitimate sub-gradient
    but not comparable entry by entry)."""
    from pointnet_autoencoder_b200.encoder import _Conv5Pool, BN_EPS
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(3)
    b, n, k, c = 3, 640, 128, 256
    rnd = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    x = rnd(b, n, k).to(torch.bfloat16).float(); w = (rnd(k, c) / k ** 0.5).to(torch.bfloat16).float()
    bias = 0.1 * rnd(c); gamma = rnd(c); beta = 0.1 * rnd(c)
    rm = 0.1 * rnd(c); rv = torch.rand(c, device="cuda", generator=g) + 0.5
    tgt = rnd(b, c)

    leaves = [t.clone().requires_grad_(True) for t in (x, w, bias, gamma, beta)]
    out = _Conv5Pool.apply(*leaves, rm.clone(), rv.clone(), training, 0.9)
    (out * tgt).sum().backward()

    ref = [t.clone().requires_grad_(True) for t in (x, w, bias, gamma, beta)]
    y = ref[0] @ ref[1] + ref[2]
    if training:
        mean = y.mean(dim=(0, 1)); var = y.var(dim=(0, 1), unbiased=False)
    else:
        mean, var = rm, rv
    rout = F.relu((y - mean) * torch.rsqrt(var + BN_EPS) * ref[3] + ref[4]).amax(dim=1)
    (rout * tgt).sum().backward()

    assert scaled_err(out.detach(), rout.detach()) < 1e-4
    for name, a, r in zip(("x", "w", "bias", "gamma", "beta"), leaves, ref):
        if name == "bias" and training:        # BN removes the mean: analytically zero, autograd leaves rounding noise
            assert float(a.grad.abs().max()) == 0.0 and float(r.grad.abs().max()) < 1e-3 * float(ref[1].grad.abs().max())
            continue
        assert scaled_err(a.grad, r.grad) < 2e-3, name


def test_encoder_backward_end_to_end():
    torch.manual_seed(0)
    enc = PointNetEncoder(fused=True).cuda().train()
    pc = torch.randn(2, 512, 3, device="cuda")
    enc(pc).square().sum().backward()
    for p_ in enc.parameters():
        assert p_.grad is not None and torch.isfinite(p_.grad).all()
