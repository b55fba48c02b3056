"""GPU parity of the tcgen05 implicit-GEMM conv kernel (K3) against torch's fp32 convolution evaluated on the
same bf16-rounded inputs and weights.  Tolerance (bf16 output, fp32 accumulation): |err| <= 1e-2*|ref| + 2e-2 on
O(1) activations, i.e. about one bf16 ulp of the result plus accumulation-order noise."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import cooperative_training_and_latent_space_data_augmentation_b200 as pkg
    return pkg.ops


def _ref(x, w, k, stride, scale, shift, res, rs, rb, act):
    y = F.conv2d(x.float(), w.to(torch.bfloat16).float(), None, stride=stride, padding=k // 2)
    C = y.shape[1]
    y = y * (scale.view(1, C, 1, 1) if scale is not None else 1.0) + (shift.view(1, C, 1, 1) if shift is not None else 0.0)
    if res is not None:
        y = y + res.float() * (rs.view(1, C, 1, 1) if rs is not None else 1.0) + (rb.view(1, C, 1, 1) if rb is not None else 0.0)
    if act == 1:
        y = F.leaky_relu(y, 0.2)
    elif act == 2:
        y = F.relu(y)
    elif act == 3:
        y = torch.sigmoid(y)
    return y


def _check(got, want):
    assert got.dtype == torch.bfloat16 and got.dim() == 5 and got.shape[-1] == 8      # C8 layout
    import cooperative_training_and_latent_space_data_augmentation_b200 as pkg
    got = pkg.ops.c8_to_nchw(got)
    err = (got.float() - want).abs()
    tol = 1e-2 * want.abs() + 2e-2
    assert bool((err <= tol).all()), "max err %.4f at ref %.4f (tol %.4f); mismatches %d / %d" % (
        float(err.max()), float(want.flatten()[err.argmax()]), float(tol.flatten()[err.argmax()]),
        int((err > tol).sum()), err.numel())


LAYERS = [  # (Cin, Cout, k, stride, N, H, W)   -- the layer classes of FCN_16_standard
    (16, 16, 3, 1, 2, 224, 224), (16, 32, 3, 1, 2, 112, 112), (32, 32, 3, 1, 2, 112, 112), (32, 64, 3, 1, 2, 56, 56),
    (64, 64, 3, 1, 2, 56, 56), (64, 128, 3, 1, 2, 28, 28), (128, 128, 3, 1, 3, 28, 28), (128, 128, 3, 1, 3, 14, 14),
    (128, 64, 3, 1, 2, 28, 28), (64, 32, 3, 1, 2, 56, 56), (32, 16, 3, 1, 2, 112, 112),
    (16, 16, 1, 1, 2, 224, 224), (16, 32, 1, 1, 2, 112, 112), (64, 128, 1, 1, 2, 28, 28), (128, 128, 1, 1, 2, 14, 14),
    (128, 64, 1, 1, 2, 28, 28), (32, 16, 1, 1, 2, 112, 112),
    (16, 16, 3, 2, 2, 224, 224), (32, 32, 3, 2, 2, 112, 112), (64, 64, 3, 2, 2, 56, 56), (128, 128, 3, 2, 2, 28, 28),
    (16, 16, 3, 1, 1, 20, 36), (32, 48, 3, 1, 2, 17, 9), (64, 16, 1, 1, 1, 5, 40), (16, 16, 3, 1, 5, 16, 16),
    # vertically packed 3x3: heights around the 14-row tile (13, 14, 15, 29), 256^2-configuration sizes, every N tile
    (16, 16, 3, 1, 2, 13, 24), (16, 16, 3, 1, 2, 15, 8), (32, 32, 3, 1, 1, 29, 33), (16, 16, 3, 1, 1, 256, 256),
    (64, 64, 3, 1, 2, 64, 64), (128, 128, 3, 1, 2, 32, 32), (128, 128, 3, 1, 2, 16, 16), (16, 64, 3, 1, 1, 30, 30),
    (32, 64, 3, 1, 1, 28, 28), (64, 16, 3, 1, 1, 28, 28), (128, 16, 3, 1, 1, 14, 14), (128, 64, 3, 1, 1, 14, 14),
]


@pytest.mark.parametrize("cin,cout,k,stride,N,H,W", LAYERS)
def test_conv_matches_torch(ops, cin, cout, k, stride, N, H, W):
    g = torch.Generator(device="cuda").manual_seed(cin * 1000 + cout + k + H)
    x = torch.randn(N, cin, H, W, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) * (2.0 / (cin * k * k)) ** 0.5
    assert ops.conv_supported(cin, cout, k)
    wp = ops.pack_conv_weight_s2(w) if (stride == 2 and k == 3) else ops.pack_conv_weight(w)
    got = ops.conv2d_c8(ops.nchw_to_c8(x), wp, cout, k * k, subsample=stride)
    _check(got, _ref(x, w, k, stride, None, None, None, None, None, 0))


@pytest.mark.parametrize("act", [0, 1, 2, 3])
@pytest.mark.parametrize("cin,cout,k", [(16, 16, 3), (64, 32, 3), (128, 64, 1), (32, 128, 3)])
def test_conv_fused_epilogue(ops, act, cin, cout, k):
    g = torch.Generator(device="cuda").manual_seed(act * 7 + cin + cout)
    N, H, W = 3, 40, 24
    x = torch.randn(N, cin, H, W, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) * (2.0 / (cin * k * k)) ** 0.5
    scale = 1 + 0.2 * torch.randn(cout, device="cuda", generator=g)
    shift = 0.3 * torch.randn(cout, device="cuda", generator=g)
    res = torch.randn(N, cout, H, W, device="cuda", generator=g).to(torch.bfloat16)
    rs = 1 + 0.2 * torch.randn(cout, device="cuda", generator=g)
    rb = 0.3 * torch.randn(cout, device="cuda", generator=g)
    wp = ops.pack_conv_weight(w)
    xc = ops.nchw_to_c8(x)
    got = ops.conv2d_c8(xc, wp, cout, k * k, scale=scale, shift=shift, res=ops.nchw_to_c8(res), res_scale=rs,
                        res_shift=rb, act=act)
    _check(got, _ref(x, w, k, 1, scale, shift, res, rs, rb, act))
    got = ops.conv2d_c8(xc, wp, cout, k * k, shift=shift, act=act)
    _check(got, _ref(x, w, k, 1, None, shift, None, None, None, act))


def test_layout_round_trip(ops):
    x = torch.randn(3, 24, 7, 9, device="cuda")
    c8 = ops.nchw_to_c8(x)
    assert tuple(c8.shape) == (3, 3, 7, 9, 8)
    assert torch.equal(ops.c8_to_nchw(c8), x.to(torch.bfloat16).float())
    assert torch.equal(c8, x.to(torch.bfloat16).view(3, 3, 8, 7, 9).permute(0, 1, 3, 4, 2).contiguous())
    assert torch.equal(ops.c8_to_nchw(ops.nchw_to_c8(x.to(torch.bfloat16)), torch.bfloat16), x.to(torch.bfloat16))


@pytest.mark.parametrize("cin,cout,H,W", [(128, 128, 14, 14), (64, 64, 28, 28), (32, 32, 56, 40), (16, 16, 112, 112)])
def test_convtranspose2x2_matches_torch(ops, cin, cout, H, W):
    g = torch.Generator(device="cuda").manual_seed(cin + H)
    x = torch.randn(2, cin, H, W, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(cin, cout, 2, 2, device="cuda", generator=g) * (1.0 / cin) ** 0.5
    b = 0.3 * torch.randn(cout, device="cuda", generator=g)
    got = ops.conv2d_c8(ops.nchw_to_c8(x), ops.pack_convtranspose2x2_weight(w), 4 * cout, 1, up2x=True,
                        shift=b.repeat(4))
    assert tuple(got.shape) == (2, cout // 8, 2 * H, 2 * W, 8)
    want = F.conv_transpose2d(x.float(), w.to(torch.bfloat16).float(), b, stride=2)
    _check(got, want)


@pytest.mark.parametrize("cin,in_mode", [(1, 0), (4, 0), (4, 1), (4, 2)])
def test_stem_conv_matches_torch(ops, cin, in_mode):
    g = torch.Generator(device="cuda").manual_seed(cin * 10 + in_mode)
    N, H, W = 3, 37, 50
    w = torch.randn(16, cin, 3, 3, device="cuda", generator=g) * 0.3
    scale = 1 + 0.2 * torch.randn(16, device="cuda", generator=g)
    shift = 0.3 * torch.randn(16, device="cuda", generator=g)
    if in_mode == 2:
        lab = torch.randint(0, 4, (N, H, W), device="cuda", generator=g)
        xin = F.one_hot(lab, 4).permute(0, 3, 1, 2).float()
        got = ops.stem_conv_c8(lab, w, scale, shift, ops.ACT_LRELU, in_mode=2)
    else:
        x = torch.randn(N, cin, H, W, device="cuda", generator=g) * 2
        xin = torch.softmax(x / 2.0, dim=1) if in_mode == 1 else x
        got = ops.stem_conv_c8(x, w, scale, shift, ops.ACT_LRELU, in_mode=in_mode, temperature=2.0)
    want = F.leaky_relu(F.conv2d(xin, w, None, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), 0.2)
    _check(got, want)


@pytest.mark.parametrize("cout,act", [(4, 0), (1, 3)])
def test_head_upsample_bn_ops(ops, cout, act):
    g = torch.Generator(device="cuda").manual_seed(cout)
    N, H, W = 2, 30, 44
    x = torch.randn(N, 16, H, W, device="cuda", generator=g).to(torch.bfloat16)
    xc = ops.nchw_to_c8(x)
    w = torch.randn(cout, 16, 1, 1, device="cuda", generator=g) * 0.3
    b = torch.randn(cout, device="cuda", generator=g)
    want = F.conv2d(x.float(), w, b)
    want = torch.sigmoid(want) if act == 3 else want
    got = ops.head_conv_c8(xc, w, b, act)
    assert got.dtype == torch.float32 and tuple(got.shape) == (N, cout, H, W)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    up = ops.c8_to_nchw(ops.upsample2x_c8(xc))
    assert torch.equal(up, F.interpolate(x.float(), scale_factor=2, mode="nearest"))
    # batch-norm statistics folded into (scale, shift), with the running-stat update of nn.BatchNorm2d
    bn = torch.nn.BatchNorm2d(16).cuda()
    bn.weight.data.normal_(1, 0.2, generator=g); bn.bias.data.normal_(0, 0.3, generator=g)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    want_y = bn(x.float())
    scale, shift = ops.bn_batch_affine_c8(xc, bn.weight, bn.bias, bn.eps, rm, rv, bn.momentum)
    torch.testing.assert_close(rm, bn.running_mean, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rv, bn.running_var, rtol=1e-4, atol=1e-5)
    y = ops.c8_to_nchw(ops.scale_shift_act_c8(xc, scale, shift, ops.ACT_LRELU))
    torch.testing.assert_close(y, F.leaky_relu(want_y, 0.2), rtol=2e-2, atol=2e-2)


def test_conv_rejects_unsupported(ops):
    x = torch.zeros(1, 8, 8, 8, device="cuda", dtype=torch.bfloat16)
    assert not ops.conv_supported(8, 16, 3) and not ops.conv_supported(16, 4, 1) and not ops.conv_supported(16, 16, 5)
    with pytest.raises(NotImplementedError):
        ops.pack_conv_weight(torch.zeros(16, 8, 3, 3, device="cuda"))


@pytest.mark.parametrize("cout,cin,k", [(16, 16, 3), (32, 16, 3), (128, 64, 3), (128, 128, 3), (64, 128, 1), (16, 32, 1)])
def test_pack_kernel_matches_torch_statement(ops, cout, cin, k):
    w = torch.randn(cout, cin, k, k, device="cuda")
    assert torch.equal(ops.pack_conv_weight(w), ops.pack_conv_weight_torch(w).reshape(-1))
    if k == 3:
        assert torch.equal(ops.pack_conv_weight_s2(w), ops.pack_conv_weight_torch(w, tap_major=True).reshape(-1))
        assert torch.equal(ops._pack_kernel(w, False, tap_major=False), ops.pack_conv_weight_torch(w, tap_major=False).reshape(-1))
    want = ops.pack_conv_weight_torch(w.flip(2, 3).transpose(0, 1)).reshape(-1)
    assert torch.equal(ops.pack_conv_weight_dgrad(w), want)


@pytest.mark.parametrize("N,H,W", [(1, 1, 1), (2, 7, 5), (1, 8, 16), (3, 33, 31), (2, 32, 32), (1, 40, 72), (2, 65, 17), (5, 9, 100)])
@pytest.mark.parametrize("variant", ["plain", "scale", "res", "res_affine", "stats", "relu"])
def test_small_channel_kernel_geometries_and_epilogues(ops, N, H, W, variant):
    """K3s (16 -> 16 channel 3x3, warp-level tensor path, csrc/conv_small.cuh): its 32 x 32 pixel tiles, 16 x 8 pixel warp
    blocks (bulk tensor stores clipped at the image border, residual blocks loaded into the output staging buffer) and
    every epilogue form against fp32 torch on the same bf16 operands, at sizes that are not multiples of the tile."""
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + H * 10 + W)
    x = torch.randn(N, 16, H, W, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(16, 16, 3, 3, device="cuda", generator=g) * (2.0 / 144) ** 0.5
    shift = 0.3 * torch.randn(16, device="cuda", generator=g)
    scale = 1 + 0.2 * torch.randn(16, device="cuda", generator=g)
    res = torch.randn(N, 16, H, W, device="cuda", generator=g).to(torch.bfloat16)
    rs = 1 + 0.2 * torch.randn(16, device="cuda", generator=g)
    rb = 0.3 * torch.randn(16, device="cuda", generator=g)
    xc, wp = ops.nchw_to_c8(x), ops.pack_conv_weight(w)
    if variant == "plain":
        got = ops.conv2d_c8(xc, wp, 16, 9, shift=shift, act=1)
        want = _ref(x, w, 3, 1, None, shift, None, None, None, 1)
    elif variant == "scale":
        got = ops.conv2d_c8(xc, wp, 16, 9, scale=scale, shift=shift, act=1)
        want = _ref(x, w, 3, 1, scale, shift, None, None, None, 1)
    elif variant == "relu":
        got = ops.conv2d_c8(xc, wp, 16, 9, shift=shift, act=2)
        want = _ref(x, w, 3, 1, None, shift, None, None, None, 2)
    elif variant == "res":
        got = ops.conv2d_c8(xc, wp, 16, 9, res=ops.nchw_to_c8(res))
        want = _ref(x, w, 3, 1, None, None, res, None, None, 0)
    elif variant == "res_affine":
        got = ops.conv2d_c8(xc, wp, 16, 9, shift=shift, res=ops.nchw_to_c8(res), res_scale=rs, res_shift=rb, act=1)
        want = _ref(x, w, 3, 1, None, shift, res, rs, rb, 1)
    else:
        stats = torch.zeros(2, 16, device="cuda", dtype=torch.float64)
        got = ops.conv2d_c8(xc, wp, 16, 9, shift=shift, stats=stats)
        want = _ref(x, w, 3, 1, None, shift, None, None, None, 0)
        stored = ops.c8_to_nchw(got).double()
        torch.testing.assert_close(stats[0], stored.sum((0, 2, 3)), rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(stats[1], (stored * stored).sum((0, 2, 3)), rtol=1e-5, atol=1e-4)
    _check(got, want)


@pytest.mark.parametrize("N,H,W", [(1, 2, 2), (2, 34, 30), (1, 64, 64), (3, 18, 66), (2, 224, 224)])
@pytest.mark.parametrize("with_stats", [False, True])
def test_small_channel_kernel_stride2(ops, N, H, W, with_stats):
    """K3s stride-2 form (16 -> 16 channel 3x3 pad 1, the encoder's first down-sampling convolution): 16 x 16 output
    tiles, every second halo pixel as the A fragment's rows, against fp32 torch; optional statistics of the stored values."""
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + H * 10 + W)
    x = torch.randn(N, 16, H, W, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(16, 16, 3, 3, device="cuda", generator=g) * (2.0 / 144) ** 0.5
    shift = 0.3 * torch.randn(16, device="cuda", generator=g)
    stats = torch.zeros(2, 16, device="cuda", dtype=torch.float64) if with_stats else None
    got = ops.conv2d_c8(ops.nchw_to_c8(x), ops.pack_conv_weight_s2(w), 16, 9, subsample=2, shift=shift, act=1, stats=stats)
    assert tuple(got.shape) == (N, 2, H // 2, W // 2, 8)
    _check(got, _ref(x, w, 3, 2, None, shift, None, None, None, 1))
    if with_stats:
        stored = ops.c8_to_nchw(got).double()
        torch.testing.assert_close(stats[0], stored.sum((0, 2, 3)), rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(stats[1], (stored * stored).sum((0, 2, 3)), rtol=1e-5, atol=1e-4)
