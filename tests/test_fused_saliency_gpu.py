"""GPU: the latent saliency reduced in the epilogue of the decoder's last input-gradient convolution + the one-kernel
masking tail against the materialised chain it replaces
(dL/dz stored -> c8_to_nchw -> K1 -> select -> K2 -> nchw_to_c8), which is itself pinned bit-exactly on the reference
fixtures (tests/test_masking_gpu.py).  Bar: s, masks, thresholds and masked codes BIT-EXACT between the two forms (the
fp64 sums of bf16 values are exact, hence independent of the accumulation order)."""
import random

import numpy as np
import pytest
import torch

from oracle import masking_oracle as mo
from oracle import weights

pytestmark = pytest.mark.gpu


@pytest.fixture()
def pkg():
    import cooperative_training_and_latent_space_data_augmentation_b200 as p
    p.conv_blocks.set_precision("kernel")
    yield p
    p.model_util.set_fused_saliency(True)
    p.set_rng_mode("torch")


@pytest.mark.parametrize("cin,H,W", [(64, 14, 14), (128, 14, 14), (64, 16, 16), (128, 16, 16), (64, 12, 20)])
@pytest.mark.parametrize("mode", [0, 1], ids=["channel", "spatial"])
def test_epilogue_sums_equal_k1_on_the_stored_gradient(pkg, cin, H, W, mode):
    ops = pkg.ops
    N, cout = 6, 128
    gen = torch.Generator(device="cuda").manual_seed(cin + H + mode)
    dy = ops.nchw_to_c8(torch.randn(N, cin, H, W, device="cuda", generator=gen) * 1e-4)
    res = ops.nchw_to_c8(torch.randn(N, cout, H, W, device="cuda", generator=gen) * 1e-4)
    wp = ops.pack_conv_weight(torch.randn(cout, cin, 1, 1, device="cuda", generator=gen) * 0.1)
    n = cout if mode == 0 else H * W
    stored = ops.conv2d_c8(dy, wp, cout, 1, res=res)                       # what the unfused path materialises
    g = ops.c8_to_nchw(stored)
    want_s = ops.saliency_reduce(g, mode)
    for store_out in (False, True):
        sums = torch.zeros(N, n, device="cuda", dtype=torch.float64)
        out = ops.conv2d_c8_saliency(dy, wp, cout, res, sums, mode, store_out=store_out)
        if store_out:
            assert torch.equal(out, stored)
        else:
            assert out is None
        count = float(H * W if mode == 0 else cout)
        got_s = (sums / count).float()
        assert torch.equal(got_s, want_s), float((got_s - want_s).abs().max())
    # the numpy oracle on the same stored gradient
    want = mo.saliency_reduce(g.cpu().numpy(), mode)
    assert np.array_equal(got_s.cpu().numpy(), want)
    # masking tail on the sums vs K1 + select + K2 on the stored gradient; soft with the native Philox stream
    z = torch.relu(torch.randn(N, cout, H, W, device="cuda", generator=gen))
    for soft in (False, True):
        k = int(n * 0.3)
        za, ma, sa, ta = ops.saliency_mask_apply(g, z, mode, k, soft=soft, rng=ops.NativeRNG(5, first_sample=8), want_thr=True)
        zb, mb, sb, tb, c8 = ops.saliency_sums_mask_apply(sums, z, mode, k, soft=soft, rng=ops.NativeRNG(5, first_sample=8),
                                                          want_thr=True)
        assert torch.equal(sa, sb) and torch.equal(ma, mb) and torch.equal(ta, tb) and torch.equal(za, zb)
        assert torch.equal(c8, ops.nchw_to_c8(zb))
        rand = mo.native_rand(5, 0, N, n, first_sample=8) if soft else None
        wz, wm, ws, _ = mo.mask_given_gradient(z.cpu().numpy(), g.cpu().numpy(), mode, k, soft=soft, rand=rand)
        assert np.array_equal(zb.cpu().numpy(), wz) and np.array_equal(mb.cpu().numpy().reshape(wm.shape), wm)
    with pytest.raises(IndexError):
        ops.saliency_sums_mask_apply(sums, z, mode, n, soft=False)


@pytest.mark.parametrize("which", ["image", "shape"])
def test_fused_masking_equals_the_materialised_chain_through_the_public_api(pkg, which):
    """mask_latent_code_* on this build's decoders: fused (default) vs set_fused_saliency(False) -- same masks, same
    masked codes, same decoder output for the masked code (which the fused form hands over in the C8 layout)."""
    torch.manual_seed(0)
    solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4)
    for k, m in solver.model.items():
        m.load_state_dict(weights.synthetic_state_dict(m, 7, prefix=k + "."))
    solver.train()
    img, lab, _ = weights.synthetic_batch(8, 224, 224, seed=4)
    img, lab = img.cuda(), lab.cuda()
    gen = torch.Generator(device="cuda").manual_seed(1)
    z = torch.relu(torch.randn(8, 128, 14, 14, device="cuda", generator=gen))
    if which == "image":
        dec, label, loss, fn = solver.model['image_decoder'], img, 'mse', pkg.mask_latent_code_channel_wise
    else:
        dec, label, loss, fn = solver.model['segmentation_decoder'], lab, 'ce', pkg.mask_latent_code_spatial_wise
    pkg.model_util.set_grad(dec, requires_grad=False)
    state0 = {n: b.clone() for n, b in dec.named_buffers()}
    calls = {"c8_to_nchw": 0, "nchw_to_c8": 0}
    orig = (pkg.ops.c8_to_nchw, pkg.ops.nchw_to_c8)

    def counted(name, f):
        def g(*a, **kw):
            calls[name] += 1
            return f(*a, **kw)
        return g

    out = {}
    try:
        pkg.ops.c8_to_nchw, pkg.ops.nchw_to_c8 = counted("c8_to_nchw", orig[0]), counted("nchw_to_c8", orig[1])
        for fused in (False, True):
            pkg.model_util.set_fused_saliency(fused)
            for n, b in dec.named_buffers():
                b.copy_(state0[n])
            random.seed(3); np.random.seed(3); torch.manual_seed(3)
            for key in calls:
                calls[key] = 0
            masked, mask = fn(z, num_classes=4, decoder_function=dec, label=label, percentile=0.5, random=True,
                              loss_type=loss, if_detach=True, if_soft=True)
            y = solver.decoder_inference(dec, masked.detach(), eval=False, disable_track_bn_stats=True)
            out[fused] = (masked.detach().clone(), mask.clone(), y.detach().float().clone(), dict(calls))
    finally:
        pkg.ops.c8_to_nchw, pkg.ops.nchw_to_c8 = orig
    (za, ma, ya, ca), (zb, mb, yb, cb) = out[False], out[True]
    assert torch.equal(ma, mb) and torch.equal(za, zb), "fused masking differs from the materialised chain"
    assert torch.equal(ya, yb)
    assert 0 < float((mb != 1).float().mean()) < 0.6
    # materialised: dL/dz converted to NCHW once, the code converted to C8 twice (saliency forward + inference);
    # fused: no NCHW gradient, and the masked code reaches the decoder without a conversion
    assert ca["c8_to_nchw"] == 1 and ca["nchw_to_c8"] == 2, ca
    assert cb["c8_to_nchw"] == 0 and cb["nchw_to_c8"] == 1, cb
    assert zb.requires_grad is False and tuple(mb.shape) == ((8, 128, 1, 1) if which == "image" else (8, 1, 14, 14))


def test_saliency_pass_reuses_the_clean_pass_forward(pkg):
    """hard_example_generation differentiates decoder(code) where `code` is the latent the clean pass has just decoded:
    with set_reuse_forward(True) (default) that forward is taken from the tape -- two decoder forwards fewer per step --
    and hard examples, BatchNorm buffers (the skipped forward's running-stat update is replayed) and the training
    backward that follows are what recomputing gives."""
    from cooperative_training_and_latent_space_data_augmentation_b200 import trainpath
    torch.manual_seed(0)
    solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4)
    for k, m in solver.model.items():
        m.load_state_dict(weights.synthetic_state_dict(m, 7, prefix=k + "."))
    img, lab, noise = weights.synthetic_batch(8, 224, 224, seed=4)
    img, lab, noise = img.cuda(), lab.cuda(), noise.cuda()
    cfg_i = {"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}
    cfg_s = {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}
    state0 = {k: {n: b.clone() for n, b in m.named_buffers()} for k, m in solver.model.items()}
    calls = {"n": 0}
    orig = trainpath.decoder_fwd

    def counted(*a, **kw):
        calls["n"] += 1
        return orig(*a, **kw)

    out = {}
    try:
        trainpath.decoder_fwd = counted
        for reuse in (False, True):
            pkg.model_util.set_reuse_forward(reuse)
            for k, m in solver.model.items():
                for n, b in m.named_buffers():
                    b.copy_(state0[k][n])
            random.seed(3); np.random.seed(3); torch.manual_seed(3)
            calls["n"] = 0
            r = pkg.cooperative_step(solver, img, lab, cfg_i, cfg_s, noise=noise, optimize=False)
            torch.cuda.synchronize()
            out[reuse] = dict(
                n=calls["n"], p_img=r['perturbed_image'].float().clone(), p_seg=r['perturbed_seg'].float().clone(),
                loss=float(r['loss']), grads=solver.flat_adam.flat_grads.clone(),
                buffers={k: {n: b.clone() for n, b in m.named_buffers()} for k, m in solver.model.items()})
    finally:
        trainpath.decoder_fwd = orig
        pkg.model_util.set_reuse_forward(True)
    a, b = out[False], out[True]
    assert a["n"] - b["n"] == 2, (a["n"], b["n"])                    # image decoder + segmentation decoder
    for key in ("p_img", "p_seg"):
        d = (a[key] - b[key]).flatten(1).norm(dim=1) / a[key].flatten(1).norm(dim=1)
        assert float(d.max()) < 1e-3, (key, d.tolist())
    assert abs(a["loss"] - b["loss"]) <= 1e-4 * abs(a["loss"])
    cos = torch.nn.functional.cosine_similarity(a["grads"].double(), b["grads"].double(), dim=0)
    assert float(cos) > 0.9999, float(cos)
    for k in a["buffers"]:
        for n, v in a["buffers"][k].items():
            w = b["buffers"][k][n]
            if n.endswith("num_batches_tracked"):
                assert int(v) == int(w), (k, n, int(v), int(w))
            else:
                torch.testing.assert_close(w, v, rtol=1e-4, atol=1e-6, msg=lambda m_, k=k, n=n: "%s.%s: %s" % (k, n, m_))
