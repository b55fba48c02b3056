"""GPU parity tests of K1/K2/dropout: CUDA path (through the C ABI) vs the CPU oracle on the same
seeded inputs, vs the fixtures the unmodified reference produced, and size-independent properties
at BASELINE.json's full microbench size.  Bit-exact for masks / selected indices / applied codes."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import masking_oracle as mo
from oracle.make_golden import MASK_CASES, mask_inputs

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pkg():
    import cooperative_training_and_latent_space_data_augmentation_b200 as p
    return p


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, np.int64(-2 ** 31) - ai, ai)
    bi = np.where(bi < 0, np.int64(-2 ** 31) - bi, bi)
    return np.abs(ai - bi)


def assert_equal_or_one_ulp(got, want, max_frac=1e-5):
    bad = got != want
    if bad.any():
        assert ulp_diff(got[bad], want[bad]).max() <= 1
        assert bad.sum() <= max(1, int(max_frac * got.size)), "%d of %d differ" % (bad.sum(), got.size)


SHAPES = [(8, 128, 14, 14), (3, 64, 28, 28), (4, 128, 16, 16), (2, 128, 12, 12), (5, 24, 7, 9), (1, 128, 14, 14),
          (2, 16, 2, 2), (3, 7, 5, 5), (2, 256, 32, 32)]


def test_philox_matches_oracle(pkg):
    for seed, off, first, cnt in [(0, 0, 0, 1000), (123, 5, 7, 4097), (2 ** 63 + 11, 2 ** 40 + 3, 2 ** 33 + 1, 513)]:
        got = pkg.ops.philox_uniform(seed, off, first, cnt).cpu().numpy()
        want = mo.philox_uniform(seed, off, np.arange(cnt, dtype=np.uint64) + np.uint64(first))
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", [mo.MODE_CHANNEL, mo.MODE_SPATIAL])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_saliency_reduce_matches_oracle(pkg, shape, mode, dtype):
    rs = np.random.RandomState(hash((shape, mode)) & 0xFFFF)
    g = (1e-5 * rs.standard_normal(shape)).astype(np.float32)
    gt = dev(g)
    if dtype == "bf16":
        gt = gt.to(torch.bfloat16)
        g = gt.float().cpu().numpy()
    got = pkg.ops.saliency_reduce(gt, mode).cpu().numpy()
    want = mo.saliency_reduce(g, mode)
    assert got.shape == want.shape
    # fp64 accumulation + one rounding on both sides: order-independent up to a double-rounding
    # coincidence (probability ~1e-8 per element), so bit-exact on these seeded inputs
    assert_equal_or_one_ulp(got, want)
    # and within 2 ulp-ish of the reference's own fp32 torch.mean
    ref = mo.saliency_reduce_reference_order(g, mode)
    np.testing.assert_allclose(got, ref, rtol=3e-5, atol=1e-11)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", [mo.MODE_CHANNEL, mo.MODE_SPATIAL])
@pytest.mark.parametrize("soft", [False, True])
def test_topp_mask_apply_bit_exact(pkg, shape, mode, soft):
    N, C, H, W = shape
    n = C if mode == mo.MODE_CHANNEL else H * W
    rs = np.random.RandomState(7 + n)
    z = np.maximum(rs.standard_normal(shape), 0).astype(np.float32)
    s = (1e-5 * rs.standard_normal((N, n))).astype(np.float32)
    if n >= 8:                                   # ties straddling the threshold + signed values + zeros
        s[:, 3] = s[:, 5]
        s[0, : n // 2] = s[0, 0]
        s[-1, :] = 0.0
    rand = rs.random_sample((N, n)).astype(np.float32)
    for p in (0.0, 0.1, 1 / 3.0, 0.5, (n - 1) / n + 1e-9):
        k = min(int(n * p), n - 1)
        thr = mo.topp_threshold(s, k)
        vec = mo.build_mask(s, thr, soft, rand)
        want_z, want_m = mo.apply_mask(z, vec, mode)
        got_z, got_m, got_thr = pkg.ops.topp_mask_apply(dev(s), dev(z), mode, k, soft=soft,
                                                        rand=dev(rand) if soft else None, want_thr=True)
        np.testing.assert_array_equal(got_thr.cpu().numpy(), thr)
        np.testing.assert_array_equal(got_m.cpu().numpy(), vec)
        np.testing.assert_array_equal(got_z.cpu().numpy(), want_z)


def test_topp_native_philox_and_shard_invariance(pkg):
    N, C, H, W = 6, 64, 14, 14
    rs = np.random.RandomState(3)
    z = np.maximum(rs.standard_normal((N, C, H, W)), 0).astype(np.float32)
    for mode in (mo.MODE_CHANNEL, mo.MODE_SPATIAL):
        n = C if mode == mo.MODE_CHANNEL else H * W
        s = rs.standard_normal((N, n)).astype(np.float32)
        k = int(n * 0.3)
        rng = pkg.ops.NativeRNG(seed=99, first_sample=40)
        rng.offset = 17
        got_z, got_m, _ = pkg.ops.topp_mask_apply(dev(s), dev(z), mode, k, soft=True, rng=rng)
        rand = mo.native_rand(99, 17, N, n, first_sample=40)
        vec = mo.build_mask(s, mo.topp_threshold(s, k), True, rand)
        np.testing.assert_array_equal(got_m.cpu().numpy(), vec)
        np.testing.assert_array_equal(got_z.cpu().numpy(), mo.apply_mask(z, vec, mode)[0])
        # a 2-way batch shard reproduces the full-batch result row for row
        for lo, hi in ((0, 3), (3, 6)):
            r2 = pkg.ops.NativeRNG(seed=99, first_sample=40 + lo)
            r2.offset = 17
            _, m2, _ = pkg.ops.topp_mask_apply(dev(s[lo:hi]), dev(z[lo:hi]), mode, k, soft=True, rng=r2)
            np.testing.assert_array_equal(m2.cpu().numpy(), vec[lo:hi])


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_topp_bf16_variants(pkg, out_dtype):
    N, C, H, W = 4, 64, 28, 28
    rs = np.random.RandomState(5)
    zt = dev(np.maximum(rs.standard_normal((N, C, H, W)), 0).astype(np.float32)).to(torch.bfloat16)
    z = zt.float().cpu().numpy()
    for mode in (mo.MODE_CHANNEL, mo.MODE_SPATIAL):
        n = C if mode == mo.MODE_CHANNEL else H * W
        s = rs.standard_normal((N, n)).astype(np.float32)
        rand = rs.random_sample((N, n)).astype(np.float32)
        k = int(n * 0.4)
        vec = mo.build_mask(s, mo.topp_threshold(s, k), True, rand)
        want = torch.from_numpy(mo.apply_mask(z, vec, mode)[0]).to(out_dtype)      # RNE, like the kernel
        got_z, got_m, _ = pkg.ops.topp_mask_apply(dev(s), zt, mode, k, soft=True, rand=dev(rand), out_dtype=out_dtype)
        assert got_z.dtype == out_dtype
        np.testing.assert_array_equal(got_m.cpu().numpy(), vec)
        assert torch.equal(got_z.cpu(), want)


@pytest.mark.parametrize("case", [c[0] for c in MASK_CASES])
def test_fused_path_matches_reference_fixture(pkg, case):
    """g -> K1 -> K2 on the GPU vs what the unmodified reference produced on the same inputs."""
    f = np.load(os.path.join(GOLDEN, "masking_%s.npz" % case))
    N, C, H, W = [int(x) for x in f["shape"]]
    mode = mo.MODE_CHANNEL if str(f["mode"]) == "channel" else mo.MODE_SPATIAL
    z, g0 = mask_inputs(N, C, H, W, int(f["seed"]))
    g = ((g0 * np.float32(z.size)) * np.float32(1.0 / z.size)).astype(np.float32)
    soft = bool(f["soft"])
    got_z, got_m, got_s, _ = pkg.ops.saliency_mask_apply(dev(g), dev(z), mode, int(f["k"]), soft=soft,
                                                         rand=dev(f["rand"]) if soft else None)
    np.testing.assert_array_equal(got_m.cpu().numpy().reshape(f["mask"].shape), f["mask"])
    gz = got_z.cpu().numpy()
    np.testing.assert_array_equal(gz.reshape(-1)[:: max(1, z.size // 257)][:257], f["masked_probe"])
    assert np.float64(gz.astype(np.float64).sum()) == f["masked_checksum"]


def test_k_out_of_range_raises_index_error(pkg):
    z = torch.zeros(2, 16, 4, 4, device="cuda")
    s = torch.zeros(2, 16, device="cuda")
    with pytest.raises(IndexError):
        pkg.ops.topp_mask_apply(s, z, mo.MODE_CHANNEL, 16)
    with pytest.raises(IndexError):
        pkg.ops.saliency_mask_apply(z, z, mo.MODE_SPATIAL, 16)
    # p = 1.0 through the drop-in raises exactly like the reference (fixture: IndexError)
    assert str(np.load(os.path.join(GOLDEN, "masking_errors.npz"))["p1_error"]) == "IndexError"
    with pytest.raises(IndexError):
        pkg.mask_latent_code_channel_wise(z + 1, lambda c: c, z + 1, percentile=1.0, loss_type='corr')


@pytest.mark.parametrize("shape", [(4, 128, 14, 14), (3, 64, 28, 28), (5, 24, 7, 9), (2, 16, 4, 4)])
@pytest.mark.parametrize("p", [0.0, 0.3, 0.5, 1.0])
def test_dropout_matches_oracle(pkg, shape, p):
    N, C, H, W = shape
    rs = np.random.RandomState(11)
    z = np.maximum(rs.standard_normal(shape), 0).astype(np.float32)
    z[0, 0, 0, 0] = np.inf
    keep = (rs.random_sample((N, C)) >= p).astype(np.float32)
    want, wmask = mo.channel_dropout(z, keep, p)
    got, gmask, gkeep = pkg.ops.channel_dropout(dev(z), p, keep=dev(keep), want_keep=True)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    np.testing.assert_array_equal(gmask.cpu().numpy(), wmask)
    np.testing.assert_array_equal(gkeep.cpu().numpy(), keep)
    # native Philox draw
    rng = pkg.ops.NativeRNG(seed=5, first_sample=9)
    rng.offset = 2
    got, gmask, gkeep = pkg.ops.channel_dropout(dev(z), p, rng=rng, want_keep=True)
    nkeep = mo.native_keep(5, 2, N, C, p, first_sample=9)
    np.testing.assert_array_equal(gkeep.cpu().numpy(), nkeep)
    if 0.0 < p < 1.0:
        want, wmask = mo.channel_dropout(z, nkeep, p)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
        np.testing.assert_array_equal(gmask.cpu().numpy(), wmask)


@pytest.mark.parametrize("case", ["drop_p50", "drop_p30", "drop_p0"])
def test_dropout_matches_reference_fixture(pkg, case):
    f = np.load(os.path.join(GOLDEN, "masking_%s.npz" % case))
    N, C, H, W = [int(x) for x in f["shape"]]
    z, _ = mask_inputs(N, C, H, W, int(f["seed"]))
    got, gmask, _ = pkg.ops.channel_dropout(dev(z), float(f["p"]), keep=dev(f["keep"]))
    go = got.cpu().numpy()
    np.testing.assert_array_equal(go.reshape(-1)[:: max(1, z.size // 257)][:257], f["masked_probe"])
    assert np.float64(go.astype(np.float64).sum()) == f["masked_checksum"]
    assert np.float64(gmask.cpu().numpy().astype(np.float64).sum()) == f["mask_checksum"]


def test_dropin_functions_match_reference_fixture(pkg):
    """The Python drop-ins end to end (autograd.grad -> K1 -> K2) on the hard-mask fixtures; for soft
    masks the device generator differs from the CPU one the fixture used, so positions must match and
    values must lie in [0, 0.5)."""
    for name, N, C, H, W, mode, p, rnd, soft, seed in MASK_CASES:
        f = np.load(os.path.join(GOLDEN, "masking_%s.npz" % name))
        z, g0 = mask_inputs(N, C, H, W, seed)
        label = dev(g0) * float(z.size)
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        fn = pkg.mask_latent_code_channel_wise if mode == "channel" else pkg.mask_latent_code_spatial_wise
        masked, mask = fn(dev(z), lambda c: c, label, num_classes=4, percentile=p, random=rnd, loss_type='corr',
                          if_detach=True, if_soft=soft)
        assert tuple(mask.shape) == tuple(int(x) for x in f["mask_shape"]) and mask.dtype == torch.float32
        assert masked.dtype == torch.float32 and tuple(masked.shape) == (N, C, H, W)
        m = mask.cpu().numpy()
        np.testing.assert_array_equal(m != 1, f["mask"] != 1)
        if soft:
            assert ((m[m != 1] >= 0) & (m[m != 1] < 0.5)).all()
        else:
            np.testing.assert_array_equal(m, f["mask"])
            np.testing.assert_array_equal(masked.detach().cpu().numpy().reshape(-1)[:: max(1, z.size // 257)][:257],
                                          f["masked_probe"])
        # masked == z * mask exactly
        assert torch.equal(masked.detach(), dev(z) * mask)


def test_if_detach_false_keeps_graph(pkg):
    z = torch.rand(2, 16, 4, 4, device="cuda", requires_grad=True)
    w = torch.rand(2, 16, 4, 4, device="cuda")
    masked, mask = pkg.mask_latent_code_channel_wise(z, lambda c: c, w, percentile=0.5, loss_type='corr',
                                                     if_detach=False)
    masked.sum().backward()
    assert torch.equal(z.grad, mask.expand_as(z))


def test_full_size_properties(pkg):
    """BASELINE.json config #4: [512,64,28,28].  The oracle would take a while here, so use properties:
    exactly k masked per sample, masked code == z*mask, linearity of K1 under power-of-two scaling,
    and a float64 checksum of K1 against torch.mean."""
    N, C, H, W = 512, 64, 28, 28
    gen = torch.Generator(device="cuda").manual_seed(0)
    z = torch.relu(torch.randn(N, C, H, W, device="cuda", generator=gen))
    g = 1e-5 * torch.randn(N, C, H, W, device="cuda", generator=gen)
    for mode, n in ((mo.MODE_CHANNEL, C), (mo.MODE_SPATIAL, H * W)):
        s = pkg.ops.saliency_reduce(g, mode)
        # float64 sum, ONE division, one rounding (torch's own double mean multiplies by 1/n instead)
        ref = g.double().sum(dim=(2, 3)) / (H * W) if mode == mo.MODE_CHANNEL else \
            (g.double().sum(dim=1) / C).view(N, -1)
        # torch divides a double tensor by a scalar as a * (1/b); exact round-half-even ties of sum/n (they do
        # occur: n = 16*49) then land 1 ulp away from the correctly rounded value the kernel and numpy produce
        assert_equal_or_one_ulp(s.cpu().numpy(), ref.float().cpu().numpy(), max_frac=1e-3)
        rows = [0, 20, 169, N - 1]
        want = (g[rows].cpu().numpy().astype(np.float64).reshape(len(rows), C, H * W).sum(2) / (H * W)
                if mode == mo.MODE_CHANNEL else
                g[rows].cpu().numpy().astype(np.float64).reshape(len(rows), C, H * W).sum(1) / C).astype(np.float32)
        np.testing.assert_array_equal(s[rows].cpu().numpy(), want)
        assert torch.equal(pkg.ops.saliency_reduce(g * 4, mode), s * 4)
        for p in (0.1, 0.3, 0.5):
            k = int(n * p)
            zt, mask, s2, thr = pkg.ops.saliency_mask_apply(g, z, mode, k, want_thr=True)
            assert torch.equal(s2, s)
            assert int((mask == 0).sum(1).min()) == k and int((mask == 0).sum(1).max()) == k
            assert torch.equal(thr, torch.sort(s, dim=1, descending=True)[0][:, k])
            mv = mask.view(N, C, 1, 1) if mode == mo.MODE_CHANNEL else mask.view(N, 1, H, W)
            assert torch.equal(zt, z * mv)
    zd, md, keep = pkg.ops.channel_dropout(z, 0.5, rng=pkg.ops.NativeRNG(1), want_keep=True)
    assert torch.equal(zd, z * (keep * 2).view(N, C, 1, 1))
    assert torch.equal(md, (z == 0).float())
    assert 0.45 < float(keep.mean()) < 0.55
