"""GPU parity of the backward kernels (K3w weight gradient on tcgen05, dgrad through K3 with transposed weights,
BatchNorm/activation backward, resampling backward, head and stem backward) against torch autograd in fp32 evaluated on
the same bf16-rounded operands.  Tolerances are relative to the largest reference magnitude of each tensor:
fp32-accumulated quantities (weight / bias / BN parameter gradients) 2e-3; bf16 activation gradients 1.5e-2
(one bf16 rounding of the result plus accumulation-order noise)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import cooperative_training_and_latent_space_data_augmentation_b200 as pkg
    return pkg.ops


def _bf(*shape, gen, scale=1.0):
    return (torch.randn(*shape, device="cuda", generator=gen) * scale).to(torch.bfloat16)


def _cmp(got, want, rel, what=""):
    got, want = got.float(), want.float()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    ref = float(want.abs().max()) + 1e-30
    err = float((got - want).abs().max())
    assert err <= rel * ref, "%s: max err %.3e vs max |ref| %.3e (ratio %.2e > %.1e)" % (what, err, ref, err / ref, rel)


WG_LAYERS = [  # (Cin, Cout, k, N, H, W)
    (16, 16, 3, 2, 224, 224), (16, 32, 3, 2, 112, 112), (32, 32, 3, 2, 112, 112), (32, 64, 3, 2, 56, 56),
    (64, 64, 3, 2, 56, 56), (64, 128, 3, 2, 28, 28), (128, 128, 3, 3, 28, 28), (128, 128, 3, 3, 14, 14),
    (128, 64, 3, 2, 28, 28), (64, 32, 3, 2, 56, 56), (32, 16, 3, 2, 112, 112),
    (16, 16, 1, 2, 224, 224), (16, 32, 1, 2, 112, 112), (64, 128, 1, 2, 28, 28), (128, 128, 1, 2, 14, 14),
    (128, 64, 1, 2, 28, 28), (32, 16, 1, 2, 112, 112), (64, 32, 1, 2, 56, 56),
    (16, 16, 3, 1, 20, 36), (32, 48, 3, 2, 17, 9), (64, 16, 1, 1, 5, 40), (16, 16, 3, 5, 16, 16), (128, 128, 3, 1, 7, 7),
]


@pytest.mark.parametrize("cin,cout,k,N,H,W", WG_LAYERS)
def test_wgrad_matches_torch(ops, cin, cout, k, N, H, W):
    g = torch.Generator(device="cuda").manual_seed(cin * 100 + cout + k + H)
    x = _bf(N, cin, H, W, gen=g)
    dy = _bf(N, cout, H, W, gen=g, scale=0.1)
    got = ops.wgrad_to_conv_weight(ops.conv_wgrad_c8(ops.nchw_to_c8(x), ops.nchw_to_c8(dy), k * k), k)
    want = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, k, k), dy.float(), padding=k // 2)
    _cmp(got, want, 2e-3, "wgrad %d->%d k%d" % (cin, cout, k))
    # straight into nn.Conv2d's layout, accumulating into a caller-provided (arena) slice
    arena = torch.zeros(cout * cin * k * k + 8, device="cuda")
    direct = ops.conv_wgrad_c8(ops.nchw_to_c8(x), ops.nchw_to_c8(dy), k * k, out=arena[4:4 + cout * cin * k * k].view(cout, cin, k, k),
                               layout='conv')
    _cmp(direct, want, 2e-3, "wgrad (conv layout)")
    assert float(arena[:4].abs().sum()) == 0 and float(arena[-4:].abs().sum()) == 0


@pytest.mark.parametrize("cin,cout,k,N,H,W", [(16, 16, 3, 2, 64, 48), (16, 32, 3, 2, 40, 24), (128, 64, 3, 2, 28, 28),
                                               (64, 128, 1, 2, 28, 28), (32, 16, 1, 2, 30, 20), (128, 128, 3, 2, 14, 14)])
def test_dgrad_through_k3(ops, cin, cout, k, N, H, W):
    g = torch.Generator(device="cuda").manual_seed(cin + cout * 7 + k)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) * (2.0 / (cin * k * k)) ** 0.5
    dy = _bf(N, cout, H, W, gen=g)
    res = _bf(N, cin, H, W, gen=g)
    got = ops.conv2d_c8(ops.nchw_to_c8(dy), ops.pack_conv_weight_dgrad(w), cin, k * k, res=ops.nchw_to_c8(res))
    want = torch.nn.grad.conv2d_input((N, cin, H, W), w.to(torch.bfloat16).float(), dy.float(), padding=k // 2) + res.float()
    _cmp(ops.c8_to_nchw(got), want, 1.5e-2, "dgrad")


@pytest.mark.parametrize("C,N,H,W", [(16, 2, 64, 48), (32, 2, 28, 36), (128, 2, 14, 14)])
def test_stride2_conv_backward_via_zero_stuffing(ops, C, N, H, W):
    g = torch.Generator(device="cuda").manual_seed(C + H)
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) * (2.0 / (C * 9)) ** 0.5
    x = _bf(N, C, H, W, gen=g)
    dy = _bf(N, C, H // 2, W // 2, gen=g)
    dyz = ops.zero_stuff2x_c8(ops.nchw_to_c8(dy))
    assert tuple(dyz.shape) == (N, C // 8, H, W, 8)
    z = ops.c8_to_nchw(dyz)
    assert torch.equal(z[:, :, ::2, ::2], dy.float()) and int((z != 0).sum()) == int((dy != 0).sum())
    xf = x.float().requires_grad_(True)
    wf = w.to(torch.bfloat16).float().requires_grad_(True)
    F.conv2d(xf, wf, None, stride=2, padding=1).backward(dy.float())
    dx = ops.conv2d_c8(dyz, ops.pack_conv_weight_dgrad(w), C, 9)
    _cmp(ops.c8_to_nchw(dx), xf.grad, 1.5e-2, "stride-2 dgrad")
    dw = ops.wgrad_to_conv_weight(ops.conv_wgrad_c8(ops.nchw_to_c8(x), dyz, 9), 3)
    _cmp(dw, wf.grad, 2e-3, "stride-2 wgrad")
    _cmp(ops.channel_sum_c8(ops.nchw_to_c8(dy)), dy.float().sum(dim=(0, 2, 3)), 2e-3, "channel sum")


@pytest.mark.parametrize("C,N,H,W", [(128, 2, 14, 14), (64, 2, 28, 20), (16, 2, 56, 56)])
def test_convtranspose_backward_via_parity_split(ops, C, N, H, W):
    g = torch.Generator(device="cuda").manual_seed(C * 3 + H)
    w = torch.randn(C, C, 2, 2, device="cuda", generator=g) * (1.0 / C) ** 0.5
    x = _bf(N, C, H, W, gen=g)
    dy = _bf(N, C, 2 * H, 2 * W, gen=g)
    xf = x.float().requires_grad_(True)
    wf = w.to(torch.bfloat16).float().requires_grad_(True)
    F.conv_transpose2d(xf, wf, None, stride=2).backward(dy.float())
    parts = ops.split_parity2x2_c8(ops.nchw_to_c8(dy))
    for d in range(4):
        assert torch.equal(ops.c8_to_nchw(parts[d]), dy.float()[:, :, d // 2::2, d % 2::2])
    xc = ops.nchw_to_c8(x)
    dW = torch.stack([ops.conv_wgrad_c8(xc, parts[d], 1)[0] for d in range(4)], dim=2).reshape(C, C, 2, 2)
    _cmp(dW, wf.grad, 2e-3, "convT wgrad")
    dW2 = torch.zeros(C, C, 2, 2, device="cuda")
    for d in range(4):
        ops.conv_wgrad_c8(xc, parts[d], 1, out=dW2, layout=('convT', d))
    _cmp(dW2, wf.grad, 2e-3, "convT wgrad (strided taps)")
    dx = None
    for d in range(4):
        wp = ops.pack_conv_weight(w[:, :, d // 2, d % 2].reshape(C, C, 1, 1))
        dx = ops.conv2d_c8(parts[d], wp, C, 1, res=dx)
    _cmp(ops.c8_to_nchw(dx), xf.grad, 2e-2, "convT dgrad")


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("C,N,H,W", [(16, 3, 40, 24), (128, 4, 14, 14), (32, 2, 112, 112)])
def test_bn_act_backward(ops, act, C, N, H, W):
    g = torch.Generator(device="cuda").manual_seed(C + act)
    a = _bf(N, C, H, W, gen=g) * 1.5 + 0.3
    dy = _bf(N, C, H, W, gen=g, scale=0.1)
    gamma = 1 + 0.2 * torch.randn(C, device="cuda", generator=g)
    beta = 0.3 * torch.randn(C, device="cuda", generator=g)
    ac = ops.nchw_to_c8(a)
    scale, shift, mean, var = ops.bn_batch_affine_c8(ac, gamma, beta, 1e-5, want_stats=True)
    h = ops.scale_shift_act_c8(ac, scale, shift, act)
    af = a.float().requires_grad_(True)
    gf, bf_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    v = F.batch_norm(af, None, None, gf, bf_, True, 0.1, 1e-5)
    hf = F.leaky_relu(v, 0.2) if act == 1 else F.relu(v) if act == 2 else v
    _cmp(ops.c8_to_nchw(h), hf.detach(), 1.5e-2, "forward")
    # use the kernel's own (bf16) h for the activation mask, as the training path does
    hk = ops.c8_to_nchw(h)
    slope = torch.where(hk > 0, 1.0, 0.2) if act == 1 else (hk > 0).float() if act == 2 else torch.ones_like(hk)
    v.backward(dy.float() * slope)
    da, dg, db, dv = ops.bn_act_bwd_c8(ops.nchw_to_c8(dy), h if act else None, ac, act, mean, var, 1e-5, gamma,
                                       want_dv=True)
    _cmp(dg, gf.grad, 3e-3, "dgamma")
    _cmp(db, bf_.grad, 3e-3, "dbeta")
    _cmp(ops.c8_to_nchw(da), af.grad, 2e-2, "da")
    _cmp(ops.c8_to_nchw(dv), dy.float() * slope, 1e-2, "dv")
    da2, _, _, _ = ops.bn_act_bwd_c8(ops.nchw_to_c8(dy), h if act else None, ac, act, mean, var, 1e-5, gamma,
                                     want_param_grads=False)
    _cmp(ops.c8_to_nchw(da2), af.grad, 2e-2, "da (no dv)")
    if act:
        # act' from sign(a*scale + shift) instead of reading h: bit-identical to the h-reading form
        da3, dg3, db3, _ = ops.bn_act_bwd_c8(ops.nchw_to_c8(dy), None, ac, act, mean, var, 1e-5, gamma,
                                             act_affine=(scale, shift))
        da4, dg4, db4, _ = ops.bn_act_bwd_c8(ops.nchw_to_c8(dy), h, ac, act, mean, var, 1e-5, gamma)
        assert torch.equal(da3, da4) and torch.equal(dg3, dg4) and torch.equal(db3, db4)
    if act:
        _cmp(ops.c8_to_nchw(ops.act_bwd_c8(ops.nchw_to_c8(dy), h, act)), dy.float() * slope, 1e-2, "act_bwd")


def test_downsample_sum(ops):
    g = torch.Generator(device="cuda").manual_seed(0)
    dy = _bf(3, 24, 14, 22, gen=g)
    got = ops.c8_to_nchw(ops.downsample2x_sum_c8(ops.nchw_to_c8(dy)))
    want = F.avg_pool2d(dy.float(), 2) * 4
    _cmp(got, want, 1e-2, "downsample2x_sum")


@pytest.mark.parametrize("N,H,W", [(3, 30, 44), (2, 64, 96), (1, 33, 20), (2, 9, 30)])
@pytest.mark.parametrize("cout,act", [(4, 0), (1, 3)])
def test_head_backward(ops, cout, act, N, H, W):
    """Head backward (logits head without activation, sigmoid reconstruction head) against torch autograd."""
    g = torch.Generator(device="cuda").manual_seed(cout)
    x = _bf(N, 16, H, W, gen=g)
    w = torch.randn(cout, 16, 1, 1, device="cuda", generator=g) * 0.3
    b = torch.randn(cout, device="cuda", generator=g)
    dy = torch.randn(N, cout, H, W, device="cuda", generator=g) * 0.01
    xf, wf, bf_ = x.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = F.conv2d(xf, wf, bf_)
    y = torch.sigmoid(y) if act == 3 else y
    y.backward(dy)
    xc = ops.nchw_to_c8(x)
    yk = ops.head_conv_c8(xc, w, b, act)
    dx, dW, db = ops.head_bwd_c8(dy, yk if act else None, xc, w, act)
    _cmp(dW, wf.grad, 2e-3, "head dW")
    _cmp(db, bf_.grad, 2e-3, "head db")
    _cmp(ops.c8_to_nchw(dx), xf.grad, 1e-2, "head dx")


@pytest.mark.parametrize("N,H,W", [(3, 37, 50), (2, 64, 96)])
@pytest.mark.parametrize("cin,in_mode", [(1, 0), (1, 1), (4, 0), (4, 1), (4, 2)])
def test_stem_backward(ops, cin, in_mode, N, H, W):
    """Stem weight gradient (CUDA-core kernel) and input gradient (warp-level tensor path, csrc/stem_dgrad_small.cuh: dy
    tiles by TMA, fp32 weights as a bf16 hi | lo fragment pair, softmax chain in registers) against torch autograd."""
    g = torch.Generator(device="cuda").manual_seed(cin * 10 + in_mode)
    w = torch.randn(16, cin, 3, 3, device="cuda", generator=g) * 0.3
    dy = _bf(N, 16, H, W, gen=g, scale=0.1)
    wf = w.clone().requires_grad_(True)
    if in_mode == 2:
        lab = torch.randint(0, 4, (N, H, W), device="cuda", generator=g)
        F.conv2d(F.one_hot(lab, 4).permute(0, 3, 1, 2).float(), wf, None, padding=1).backward(dy.float())
        got_w = ops.stem_wgrad_c8(ops.nchw_to_c8(dy), lab, cin, in_mode=2)
        _cmp(got_w, wf.grad, 2e-3, "stem wgrad (labels)")
        return
    x = torch.randn(N, cin, H, W, device="cuda", generator=g) * 2
    xf = x.clone().requires_grad_(True)
    xin = torch.softmax(xf / 2.0, dim=1) if in_mode == 1 else xf
    F.conv2d(xin, wf, None, padding=1).backward(dy.float())
    dyc = ops.nchw_to_c8(dy)
    _cmp(ops.stem_wgrad_c8(dyc, x, cin, in_mode=in_mode, temperature=2.0), wf.grad, 2e-3, "stem wgrad")
    _cmp(ops.stem_dgrad_c8(dyc, x, w, in_mode=in_mode, temperature=2.0), xf.grad, 2e-3, "stem dgrad")


@pytest.mark.parametrize("act", [0, 1])
def test_scale_shift_upadd_act(ops, act):
    """act(x*scale + shift + nearest_up2(low)) against torch on the same bf16 operands (one bf16 rounding of the result)."""
    g = torch.Generator(device="cuda").manual_seed(21 + act)
    x = _bf(3, 24, 10, 12, gen=g)
    low = _bf(3, 24, 5, 6, gen=g)
    scale = 1 + 0.2 * torch.randn(24, device="cuda", generator=g)
    shift = 0.3 * torch.randn(24, device="cuda", generator=g)
    want = x.float() * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + F.interpolate(low.float(), scale_factor=2, mode='nearest')
    if act:
        want = F.leaky_relu(want, 0.2)
    got = ops.scale_shift_upadd_act_c8(ops.nchw_to_c8(x), scale, shift, ops.nchw_to_c8(low), act)
    _cmp(ops.c8_to_nchw(got), want, 1e-2, "upadd")
    got1 = ops.scale_shift_upadd_act_c8(ops.nchw_to_c8(x), None, None, ops.nchw_to_c8(low), act)
    want1 = x.float() + F.interpolate(low.float(), scale_factor=2, mode='nearest')
    _cmp(ops.c8_to_nchw(got1), F.leaky_relu(want1, 0.2) if act else want1, 1e-2, "upadd (no affine)")


@pytest.mark.parametrize("cin,in_mode", [(1, 0), (4, 0), (4, 1), (4, 2)])
def test_stem_input_c8_and_tensor_core_stem(ops, cin, in_mode):
    """ctl_stem_input_c8: the stem input as 16 C8 channels (image / softmax(x/T) / one-hot, zero padding), and the stem
    convolution + its weight gradient through K3 / K3w on it against torch fp32 on the same bf16-rounded operands."""
    g = torch.Generator(device="cuda").manual_seed(31 + cin + in_mode)
    N, H, W = 3, 24, 40
    if in_mode == 2:
        x = torch.randint(0, cin, (N, H, W), device="cuda", generator=g)
        want_in = F.one_hot(x, cin).permute(0, 3, 1, 2).float()
    else:
        x = torch.randn(N, cin, H, W, device="cuda", generator=g)
        want_in = torch.softmax(x / 2.0, dim=1) if in_mode == 1 else x
    xin = ops.stem_input_c8(x, cin, in_mode, 2.0)
    got_in = ops.c8_to_nchw(xin)
    assert float(got_in[:, 3 * cin:].abs().max()) == 0.0
    assert torch.equal(got_in[:, :cin], got_in[:, 2 * cin:3 * cin])
    _cmp(got_in[:, :cin] + got_in[:, cin:2 * cin], want_in, 2e-5, "stem input (hi + lo)")
    w = 0.2 * torch.randn(16, cin, 3, 3, device="cuda", generator=g)
    wf = w.clone().requires_grad_(True)
    y = F.conv2d(want_in, wf, padding=1)                       # fp32 reference on the UNROUNDED operands
    got = ops.conv2d_c8(xin, ops.pack_conv_weight(ops.pad_stem_weight(w)), 16, 9)
    _cmp(ops.c8_to_nchw(got), y.detach(), 5e-3, "tensor-core stem forward")      # one bf16 rounding of the output
    dy = _bf(N, 16, H, W, gen=g, scale=0.1)
    y.backward(dy.float())
    dW16 = ops.conv_wgrad_c8(xin, ops.nchw_to_c8(dy), 9, layout='conv')
    _cmp(ops.stem_weight_grad(dW16, cin), wf.grad, 1e-4, "tensor-core stem wgrad")


@pytest.mark.parametrize("C,H,W,want_dv", [(16, 40, 24, False), (32, 28, 28, True), (128, 14, 14, False), (64, 9, 7, True)])
def test_bn_backward_totals_form_equals_the_three_launch_form(ops, C, H, W, want_dv):
    """ctl_bn_bwd_c8 (reduction -> per-channel fp64 totals -> coefficients formed in the apply pass's prologue) against
    ctl_bn_bwd_reduce_c8 + finalise + ctl_bn_bwd_apply_c8: same sums in a different order; dgamma / dbeta 1e-4, da equal up to isolated bf16 roundings."""
    g = torch.Generator(device="cuda").manual_seed(C + H)
    N = 5
    a = ops.nchw_to_c8(torch.randn(N, C, H, W, device="cuda", generator=g))
    dy = ops.nchw_to_c8(torch.randn(N, C, H, W, device="cuda", generator=g) * 0.1)
    gamma = 1 + 0.2 * torch.randn(C, device="cuda", generator=g)
    beta = 0.3 * torch.randn(C, device="cuda", generator=g)
    scale, shift, mean, var = ops.bn_batch_affine_c8(a, gamma, beta, 1e-5, want_stats=True)
    h = ops.scale_shift_act_c8(a, scale, shift, ops.ACT_LRELU)
    kw = dict(want_dv=True) if want_dv else dict(act_affine=(scale, shift))
    want = ops.bn_act_bwd_c8(dy, h, a, ops.ACT_LRELU, mean, var, 1e-5, gamma, **kw)
    totals = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    got = ops.bn_act_bwd_c8(dy, h, a, ops.ACT_LRELU, mean, var, 1e-5, gamma, totals=totals, **kw)
    assert float(totals.abs().sum()) > 0
    # da is bf16: the two forms may round an element to neighbouring bf16 values (coefficients formed in fp32 vs fp64)
    torch.testing.assert_close(got[0].float(), want[0].float(), rtol=8e-3, atol=1e-5 * float(want[0].float().abs().max()))
    assert float((got[0].float() != want[0].float()).float().mean()) < 1e-3
    for w_, g_ in zip(want[1:3], got[1:3]):
        torch.testing.assert_close(g_.float(), w_.float(), rtol=1e-4, atol=1e-5 * float(w_.float().abs().max()) + 1e-9)
    if want_dv:
        assert torch.equal(got[3], want[3])


@pytest.mark.parametrize("cin,cout,H,W,act", [(16, 16, 42, 40, 1), (32, 32, 28, 24, 1), (64, 32, 14, 16, 2), (16, 32, 30, 8, 1)])
def test_dgrad_epilogue_accumulates_the_bn_backward_sums(ops, cin, cout, H, W, act):
    """ctl_conv2d_c8_bf16_bnbwd: the convolution's output is bit-identical to the plain kernel's, and the totals its
    epilogue leaves (sum dv | sum dv*a, dv = dy * act'(a*scale + shift)) equal the stand-alone reduction's on that output;
    the apply pass from those totals reproduces the three-launch backward."""
    assert ops.conv_bnbwd_fusable(cin, cout)
    g = torch.Generator(device="cuda").manual_seed(cin + cout + H)
    N = 3
    x = ops.nchw_to_c8(torch.randn(N, cin, H, W, device="cuda", generator=g) * 0.1)
    wp = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.1)
    a = ops.nchw_to_c8(torch.randn(N, cout, H, W, device="cuda", generator=g))
    gamma = 1 + 0.2 * torch.randn(cout, device="cuda", generator=g)
    beta = 0.3 * torch.randn(cout, device="cuda", generator=g)
    scale, shift, mean, var = ops.bn_batch_affine_c8(a, gamma, beta, 1e-5, want_stats=True)
    dy_plain = ops.conv2d_c8(x, wp, cout, 9)
    totals = torch.zeros(2 * cout, device="cuda", dtype=torch.float64)
    dy = ops.conv2d_c8_bnbwd(x, wp, cout, a, scale, shift, act, totals)
    assert torch.equal(dy, dy_plain)
    want_tot = torch.zeros(2 * cout, device="cuda", dtype=torch.float64)
    want = ops.bn_act_bwd_c8(dy_plain, None, a, act, mean, var, 1e-5, gamma, act_affine=(scale, shift), totals=want_tot)
    torch.testing.assert_close(totals, want_tot, rtol=1e-5, atol=1e-7 * float(want_tot.abs().max()))
    da, dg, db = ops.bn_bwd_apply_totals_c8(dy, a, act, mean, var, 1e-5, gamma, totals, (scale, shift))
    torch.testing.assert_close(da.float(), want[0].float(), rtol=8e-3, atol=1e-5 * float(want[0].float().abs().max()))
    torch.testing.assert_close(dg, want[1], rtol=1e-4, atol=1e-5 * float(want[1].abs().max()))
    torch.testing.assert_close(db, want[2], rtol=1e-4, atol=1e-5 * float(want[2].abs().max()))
    assert not ops.conv_bnbwd_fusable(128, 16) and not ops.conv_bnbwd_fusable(16, 64)


@pytest.mark.parametrize("C,H,W,with_low,act", [(16, 42, 40, False, 1), (32, 28, 24, True, 1), (64, 14, 16, True, 1),
                                                 (128, 14, 14, False, 2), (16, 224, 224, True, 1), (256, 6, 8, False, 0)])
def test_bn_finalisation_in_the_apply_prologue(ops, C, H, W, with_low, act):
    """ctl_bn_apply_from_sums_c8 against ctl_bn_affine_from_sums + ctl_scale_shift[_upadd]_act_c8 on the same sums:
    scale / shift / mean / var and the running statistics to 1e-6 relative (fp32 rsqrt instead of an fp64 one), the
    bf16 result equal up to isolated neighbouring roundings."""
    g = torch.Generator(device="cuda").manual_seed(C + H)
    N = 3
    a = ops.nchw_to_c8(torch.randn(N, C, H, W, device="cuda", generator=g) * 1.3 + 0.2)
    af = ops.c8_to_nchw(a).double()
    sums = torch.stack([af.sum((0, 2, 3)), (af * af).sum((0, 2, 3))]).contiguous()
    gamma = 1 + 0.2 * torch.randn(C, device="cuda", generator=g)
    beta = 0.3 * torch.randn(C, device="cuda", generator=g)
    low = ops.nchw_to_c8(torch.randn(N, C, H // 2, W // 2, device="cuda", generator=g)) if with_low else None
    rm0, rv0 = torch.randn(C, device="cuda", generator=g), torch.rand(C, device="cuda", generator=g) + 0.5
    rm_w, rv_w, rm_g, rv_g = rm0.clone(), rv0.clone(), rm0.clone(), rv0.clone()
    want = ops.bn_affine_from_sums(sums, N * H * W, gamma, beta, 1e-5, rm_w, rv_w, 0.1)
    h_want = ops.scale_shift_act_c8(a, want[0], want[1], act) if low is None else \
        ops.scale_shift_upadd_act_c8(a, want[0], want[1], low, act)
    got = ops.bn_apply_from_sums_c8(a, sums, gamma, beta, 1e-5, act, rm_g, rv_g, 0.1, low=low)
    for w_, g_, what in zip(want, got[1:], ("scale", "shift", "mean", "var")):
        torch.testing.assert_close(g_, w_, rtol=2e-6, atol=2e-6, msg=lambda m: what + ": " + m)
    torch.testing.assert_close(rm_g, rm_w, rtol=2e-6, atol=2e-6)
    torch.testing.assert_close(rv_g, rv_w, rtol=2e-6, atol=2e-6)
    torch.testing.assert_close(got[0].float(), h_want.float(), rtol=8e-3, atol=1e-5)
    assert float((got[0].float() != h_want.float()).float().mean()) < 1e-3
    # and against torch's own train-mode BatchNorm
    v = F.batch_norm(af.float(), None, None, gamma, beta, True, 0.1, 1e-5)
    if low is not None:
        v = v + F.interpolate(ops.c8_to_nchw(low).float(), scale_factor=2, mode="nearest")
    v = F.leaky_relu(v, 0.2) if act == 1 else F.relu(v) if act == 2 else v
    _cmp(ops.c8_to_nchw(got[0]), v, 1e-2, "against F.batch_norm")


@pytest.mark.parametrize("C,H,W", [(16, 47, 53), (16, 97, 101), (32, 64, 64), (64, 46, 46)])
@pytest.mark.parametrize("form", ["from_a", "with_h_dv"])
def test_plane_structured_bn_backward(ops, C, H, W, form):
    """The persistent, plane-structured BatchNorm-backward kernels (planes of >= 2048 positions: U = 2 below 8192, U = 4
    above; odd sizes leave ragged last items) against torch autograd in fp32 on the same bf16 operands: the reduction
    with the activation recomputed from a, and the residual-tail form (slope from h, dv materialised)."""
    g = torch.Generator(device="cuda").manual_seed(C + H)
    N = 3
    a = _bf(N, C, H, W, gen=g) * 1.5 + 0.3
    dy = _bf(N, C, H, W, gen=g, scale=0.1)
    gamma = 1 + 0.2 * torch.randn(C, device="cuda", generator=g)
    beta = 0.3 * torch.randn(C, device="cuda", generator=g)
    ac = ops.nchw_to_c8(a)
    scale, shift, mean, var = ops.bn_batch_affine_c8(ac, gamma, beta, 1e-5, want_stats=True)
    h = ops.scale_shift_act_c8(ac, scale, shift, ops.ACT_LRELU)
    af = a.float().requires_grad_(True)
    gf, bf_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    v = F.batch_norm(af, None, None, gf, bf_, True, 0.1, 1e-5)
    _cmp(ops.c8_to_nchw(h), F.leaky_relu(v, 0.2).detach(), 1.5e-2, "forward apply")
    hk = ops.c8_to_nchw(h)
    slope = torch.where(hk > 0, 1.0, 0.2)
    v.backward(dy.float() * slope)
    totals = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    if form == "from_a":
        da, dg, db, _ = ops.bn_act_bwd_c8(ops.nchw_to_c8(dy), None, ac, ops.ACT_LRELU, mean, var, 1e-5, gamma,
                                          act_affine=(scale, shift), totals=totals)
    else:
        da, dg, db, dv = ops.bn_act_bwd_c8(ops.nchw_to_c8(dy), h, ac, ops.ACT_LRELU, mean, var, 1e-5, gamma, want_dv=True,
                                           totals=totals)
        _cmp(ops.c8_to_nchw(dv), dy.float() * slope, 1e-2, "dv")
    _cmp(dg, gf.grad, 3e-3, "dgamma")
    _cmp(db, bf_.grad, 3e-3, "dbeta")
    _cmp(ops.c8_to_nchw(da), af.grad, 2e-2, "da")


@pytest.mark.parametrize("C,Hl,Wl", [(16, 47, 53), (32, 48, 48), (16, 112, 112)])
def test_plane_structured_up_block_tail(ops, C, Hl, Wl):
    """ctl_scale_shift_upadd_act_c8 / ctl_bn_apply_from_sums_c8 with `low` on planes of >= 2048 low-resolution positions."""
    g = torch.Generator(device="cuda").manual_seed(C + Hl)
    N = 2
    x = _bf(N, C, 2 * Hl, 2 * Wl, gen=g)
    low = _bf(N, C, Hl, Wl, gen=g)
    scale = 1 + 0.2 * torch.randn(C, device="cuda", generator=g)
    shift = 0.3 * torch.randn(C, device="cuda", generator=g)
    got = ops.scale_shift_upadd_act_c8(ops.nchw_to_c8(x), scale, shift, ops.nchw_to_c8(low), ops.ACT_LRELU)
    want = F.leaky_relu(x.float() * scale.view(1, C, 1, 1) + shift.view(1, C, 1, 1) +
                        F.interpolate(low.float(), scale_factor=2, mode="nearest"), 0.2)
    _cmp(ops.c8_to_nchw(got), want, 1e-2, "scale/shift + up-sampled shortcut + LReLU")
    xf = x.double()
    sums = torch.stack([xf.sum((0, 2, 3)), (xf * xf).sum((0, 2, 3))]).contiguous()
    gamma = 1 + 0.2 * torch.randn(C, device="cuda", generator=g)
    beta = 0.3 * torch.randn(C, device="cuda", generator=g)
    got2 = ops.bn_apply_from_sums_c8(ops.nchw_to_c8(x), sums, gamma, beta, 1e-5, ops.ACT_LRELU, low=ops.nchw_to_c8(low))[0]
    want2 = F.leaky_relu(F.batch_norm(x.float(), None, None, gamma, beta, True, 0.1, 1e-5) +
                         F.interpolate(low.float(), scale_factor=2, mode="nearest"), 0.2)
    _cmp(ops.c8_to_nchw(got2), want2, 1e-2, "BatchNorm from sums + up-sampled shortcut + LReLU")
