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


# the SMs split the (channel block, element, point tile) stream evenly, so these shapes exercise: whole elements per SM,
# elements cut in two (32, 2048), in many one-tile parts (1, 4096), ranges that cross a channel-block boundary (40, 512),
# ragged last tiles, and fewer tiles than SMs
@pytest.mark.parametrize("b,n,k,c", [(2, 256, 128, 1024), (3, 2048, 128, 1024), (2, 300, 128, 256), (1, 1000, 64, 128),
                                     (4, 37, 128, 128), (32, 2048, 128, 1024), (1, 4096, 128, 1024), (5, 2125, 128, 1024),
                                     (40, 512, 64, 256), (7, 3000, 128, 384)])
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
    # the merge of an element's parts is in point order whichever SM finishes last: bitwise repeatable, and the tickets
    # in the cached workspace are left ready for the next call
    again = ops.encoder_conv_pool(xb, wb.t().contiguous())
    for a, r in zip(again, (vmax, vmin, vsum, vsq)):
        assert torch.equal(a, r)


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


@pytest.mark.parametrize("b,n,k,c", [(3, 700, 128, 256), (8, 2300, 128, 1024), (1, 4096, 128, 1024)])
def test_conv_pool_argext(b, n, k, c):
    g = torch.Generator(device="cuda").manual_seed(2)
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


# ---- layers 1-4 (csrc/shared_mlp.cu) against an fp64 PyTorch evaluation of the same layer ---------------------------
@pytest.mark.parametrize("b,n", [(2, 256), (1, 1000), (3, 2048), (32, 2048), (1, 37)])
def test_mlp_first_layer_vs_torch(b, n):
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(b, n, 3, device="cuda", generator=g)
    w = torch.randn(3, 64, device="cuda", generator=g); bias = torch.randn(64, device="cuda", generator=g)
    y, st = ops.mlp_first(x, w, bias)
    ref = (x.double().reshape(-1, 3) @ w.double() + bias.double())
    assert y.shape == (b * n, 64) and scaled_err(y, ref.float()) < 1e-6
    assert scaled_err(st[0].double(), ref.sum(0)) < 1e-5 and scaled_err(st[1].double(), (ref * ref).sum(0)) < 1e-5


@pytest.mark.parametrize("npts,kout,training", [(256, 64, True), (1000, 64, False), (65536, 64, True), (65536, 128, True), (37, 128, False), (4133, 128, True)])
def test_mlp_layer_vs_torch(npts, kout, training):
    """previous layer's BatchNorm (batch or moving statistics) + ReLU on the way in, 3xTF32 product (fp32 accuracy, NOT
    TF32's 1e-3), bias, statistics of the output, TF-style moving-average update of the previous layer's statistics"""
    g = torch.Generator(device="cuda").manual_seed(5)
    rnd = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    yprev = rnd(npts, 64) * 2 + 0.3
    gamma = rnd(64); beta = 0.2 * rnd(64)
    mm = 0.3 * rnd(64); mv = torch.rand(64, device="cuda", generator=g) + 0.5
    w = rnd(64, kout) / 8; bias = rnd(kout)
    eps, decay = 1e-3, 0.9
    yd = yprev.double()
    if training:
        mean = yd.mean(0); var = yd.var(0, unbiased=False)
        st_prev = torch.stack([yd.sum(0), (yd * yd).sum(0)]).float().contiguous()
    else:
        mean, var, st_prev = mm.double(), mv.double(), None
    mm_k, mv_k = mm.clone(), mv.clone()
    y, st = ops.mlp_layer(yprev, st_prev, gamma, beta, mm_k, mv_k, training, decay, eps, w, bias)
    a = torch.relu((yd - mean) * torch.rsqrt(var + eps) * gamma.double() + beta.double())
    ref = a @ w.double() + bias.double()
    assert y.shape == (npts, kout)
    assert scaled_err(y.double(), ref) < 1e-5, "the 3xTF32 split must keep fp32 accuracy"
    assert scaled_err(st[0].double(), ref.sum(0)) < 1e-4 and scaled_err(st[1].double(), (ref * ref).sum(0)) < 1e-5
    if training:      # moving = decay * moving + (1 - decay) * batch, biased variance (tf.contrib.layers.batch_norm)
        assert scaled_err(mm_k.double(), decay * mm.double() + (1 - decay) * mean) < 1e-5
        assert scaled_err(mv_k.double(), decay * mv.double() + (1 - decay) * var) < 1e-4
    else:
        assert torch.equal(mm_k, mm) and torch.equal(mv_k, mv)


def test_mlp_apply_bf16():
    g = torch.Generator(device="cuda").manual_seed(6)
    y = torch.randn(4133, 128, device="cuda", generator=g)
    gamma = torch.randn(128, device="cuda", generator=g); beta = 0.1 * torch.randn(128, device="cuda", generator=g)
    mm = 0.1 * torch.randn(128, device="cuda", generator=g); mv = torch.rand(128, device="cuda", generator=g) + 0.5
    out = ops.mlp_apply_bf16(y, None, gamma, beta, mm, mv, False, 0.9, 1e-3)
    s = gamma * torch.rsqrt(mv + 1e-3); t = beta - mm * s
    ref = torch.relu(y * s + t)
    # bf16 rounding of an fp32 value that may itself differ in the last fp32 bit: one bf16 ulp (2^-8 relative)
    assert out.dtype == torch.bfloat16 and float((out.float() - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())


def test_conv5_finish_vs_torch():
    g = torch.Generator(device="cuda").manual_seed(7)
    b, c, n = 5, 256, 300
    rnd = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    y0 = rnd(b, n, c)
    vmax, vmin, vsum, vsq = y0.amax(1), y0.amin(1), y0.sum(1), (y0 * y0).sum(1)
    bias = 0.1 * rnd(c); gamma = rnd(c); beta = 0.1 * rnd(c); mm = 0.1 * rnd(c); mv = torch.rand(c, device="cuda", generator=g) + 0.5
    for training in (True, False):
        mm_k, mv_k = mm.clone(), mv.clone()
        pooled, inv, mean0, ext0, z = ops.conv5_finish(vmax, vmin, vsum, vsq, b * n, bias, gamma, beta, mm_k, mv_k, training, 0.9, 1e-3)
        y = y0.double() + bias.double()
        if training:
            mean, var = y.mean((0, 1)), y.var((0, 1), unbiased=False)
        else:
            mean, var = mm.double(), mv.double()
        ref = torch.relu((y - mean) * torch.rsqrt(var + 1e-3) * gamma.double() + beta.double()).amax(1)
        assert scaled_err(pooled.double(), ref) < 1e-5
        if training:
            assert scaled_err(mm_k.double(), 0.9 * mm.double() + 0.1 * mean) < 1e-5 and scaled_err(mv_k.double(), 0.9 * mv.double() + 0.1 * var) < 1e-4


@pytest.mark.parametrize("training", [True, False])
def test_encoder_chain_layers_1_to_4_match_the_library_path(training):
    """PointNetEncoder(fused=True) (every layer a hand-written kernel) against fused="conv5" (library layers 1-4, the
    same conv5 kernel): layers 1-4 keep fp32 accuracy, so the two agree far inside conv5's bf16 tolerance, and the
    moving statistics of layers 1-4 are updated identically."""
    torch.manual_seed(1)
    a = PointNetEncoder(fused=True).cuda(); r = PointNetEncoder(fused="conv5").cuda()
    r.load_state_dict(a.state_dict())
    a.train(training); r.train(training)
    pc = torch.randn(4, 1024, 3, device="cuda")
    ya = a(pc); yr = r(pc)
    assert scaled_err(ya, yr) < 5e-3
    if training:
        for la, lr in zip(a.layers, r.layers):
            assert scaled_err(la.moving_mean, lr.moving_mean) < 1e-4 and scaled_err(la.moving_var, lr.moving_var) < 1e-3
    # backward: gradients of every parameter against the library-layer path
    pa = pc.clone().requires_grad_(True); pr = pc.clone().requires_grad_(True)
    tgt = torch.randn(4, 1024, device="cuda")
    (a(pa) * tgt).sum().backward(); (r(pr) * tgt).sum().backward()
    assert scaled_err(pa.grad, pr.grad) < 5e-2
    for (na, p_a), (_, p_r) in zip(a.named_parameters(), r.named_parameters()):
        if p_r.grad is None or float(p_r.grad.abs().max()) == 0.0 or (training and na.endswith(".bias")):
            continue          # (training-mode BN removes the mean: bias gradients are rounding noise around zero)
        assert scaled_err(p_a.grad, p_r.grad) < 5e-2, na


@pytest.mark.parametrize("b,n,kout,training", [(32, 2048, 64, True), (3, 1000, 64, True), (2, 300, 128, True), (4, 512, 64, False)])
def test_layer1_folded_into_layer2_matches_the_two_kernel_path(b, n, kout, training):
    """pnae_xyz_moments + pnae_mlp_layer_xyz (layer 1 computed inside layer 2's producers, its BatchNorm statistics from
    the moments of xyz) against pnae_mlp_first + pnae_mlp_layer: same output, same statistics, same moving averages."""
    g = torch.Generator(device="cuda").manual_seed(11)
    rnd = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    xyz = rnd(b, n, 3) * 0.6 + 0.3                      # off-centre: the variance is a difference of second moments
    w1, b1 = rnd(3, 64), 0.5 * rnd(64)
    gamma1, beta1 = 1.0 + 0.2 * rnd(64), 0.1 * rnd(64)
    w2, b2 = rnd(64, kout) / 8, 0.1 * rnd(kout)
    mm = [0.1 * rnd(64), None]; mv = [0.5 + torch.rand(64, device="cuda", generator=g), None]
    mm[1], mv[1] = mm[0].clone(), mv[0].clone()
    y1, st1 = ops.mlp_first(xyz, w1, b1)
    ref, ref_st = ops.mlp_layer(y1, st1, gamma1, beta1, mm[0], mv[0], training, 0.9, 1e-3, w2, b2)
    mom = ops.xyz_moments(xyz) if training else None
    if training:                                        # the moments themselves, against float64 sums
        x64 = xyz.double().reshape(-1, 3)
        want = torch.cat([x64.sum(0), (x64[:, 0:1] * x64).sum(0), (x64[:, 1:2] * x64[:, 1:]).sum(0), (x64[:, 2] ** 2).sum().view(1)])
        assert float((mom - want).abs().max() / want.abs().max()) < 1e-12
    out, st = ops.mlp_layer_xyz(xyz, mom, w1, b1, gamma1, beta1, mm[1], mv[1], training, 0.9, 1e-3, w2, b2)
    assert scaled_err(out, ref) < 2e-5
    assert scaled_err(st, ref_st) < 2e-5
    assert scaled_err(mm[1], mm[0]) < 1e-5 and scaled_err(mv[1], mv[0]) < 1e-5
