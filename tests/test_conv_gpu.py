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
    assert got.dtype == torch.bfloat16 and got.is_contiguous(memory_format=torch.channels_last)
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
]


@pytest.mark.parametrize("cin,cout,k,stride,N,H,W", LAYERS)
def test_conv_matches_torch(ops, cin, cout, k, stride, N, H, W):
    g = torch.Generator(device="cuda").manual_seed(cin * 1000 + cout + k + H)
    x = torch.randn(N, cin, H, W, device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) * (2.0 / (cin * k * k)) ** 0.5
    assert ops.conv_supported(cin, cout, k)
    wp = ops.pack_conv_weight(w)
    got = ops.conv2d_bf16(x, wp, cout, k * k, subsample=stride)
    _check(got, _ref(x, w, k, stride, None, None, None, None, None, 0))


@pytest.mark.parametrize("act", [0, 1, 2, 3])
@pytest.mark.parametrize("cin,cout,k", [(16, 16, 3), (64, 32, 3), (128, 64, 1), (32, 128, 3)])
def test_conv_fused_epilogue(ops, act, cin, cout, k):
    g = torch.Generator(device="cuda").manual_seed(act * 7 + cin + cout)
    N, H, W = 3, 40, 24
    x = torch.randn(N, cin, H, W, device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) * (2.0 / (cin * k * k)) ** 0.5
    scale = 1 + 0.2 * torch.randn(cout, device="cuda", generator=g)
    shift = 0.3 * torch.randn(cout, device="cuda", generator=g)
    res = torch.randn(N, cout, H, W, device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    rs = 1 + 0.2 * torch.randn(cout, device="cuda", generator=g)
    rb = 0.3 * torch.randn(cout, device="cuda", generator=g)
    wp = ops.pack_conv_weight(w)
    got = ops.conv2d_bf16(x, wp, cout, k * k, scale=scale, shift=shift, res=res, res_scale=rs, res_shift=rb, act=act)
    _check(got, _ref(x, w, k, 1, scale, shift, res, rs, rb, act))
    got = ops.conv2d_bf16(x, wp, cout, k * k, shift=shift, act=act)
    _check(got, _ref(x, w, k, 1, None, shift, None, None, None, act))


def test_conv_rejects_unsupported(ops):
    x = torch.zeros(1, 8, 8, 8, device="cuda", dtype=torch.bfloat16)
    assert not ops.conv_supported(8, 16, 3) and not ops.conv_supported(16, 4, 1) and not ops.conv_supported(16, 16, 5)
    with pytest.raises(NotImplementedError):
        ops.pack_conv_weight(torch.zeros(16, 8, 3, 3, device="cuda"))
