"""GPU: forward + backward of every FTN/STN sub-network on the kernels (trainpath.py, precision mode 'kernel') against
the same reference-shaped torch modules in fp32 (precision mode 'fp32').  bf16 activations / activation gradients
through ~20 layers on a tiny batch (BatchNorm over as few as 48 values): outputs relative L2 < 2e-2.  Gradients cannot
agree tightly with an fp32 run: a 1-2 % activation error flips the LeakyReLU/ReLU mask of the ~1 % of pre-activations that
sit next to zero, and every flipped element is a full-size gradient error (relative L2 ~ sqrt(flip fraction) ~ 0.1-0.2
per stage).  Bars: cosine > 0.95 per sub-network, relative L2 < 0.4 per parameter tensor, < 0.35 for input gradients;
the per-kernel tests in test_bwd_kernels_gpu.py hold the tight (1e-3 .. 2e-2) bounds (tensors whose reference gradient is numerically zero -- conv biases in front of a BatchNorm -- are
returned as exact zeros and checked to be negligible in the reference)."""
import random

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import weights

pytestmark = pytest.mark.gpu


@pytest.fixture()
def env():
    import cooperative_training_and_latent_space_data_augmentation_b200 as pkg
    from cooperative_training_and_latent_space_data_augmentation_b200 import networks
    nets = {
        'image_encoder': networks.Dual_Branch_Encoder(1, 128, 128, feature_reduce=4, norm=nn.BatchNorm2d),
        'segmentation_decoder': networks.MyDecoder(128, 4, feature_reduce=4, norm=nn.BatchNorm2d, up_type='NN'),
        'shape_encoder': networks.MyEncoder(4, 128, feature_reduce=4, norm=nn.BatchNorm2d, act=nn.ReLU()),
        'image_decoder': networks.MyDecoder(128, 1, feature_reduce=4, norm=nn.BatchNorm2d, up_type='Conv2',
                                            last_act=nn.Sigmoid()),
    }
    for k, m in nets.items():
        m.load_state_dict(weights.synthetic_state_dict(m, 7, prefix=k + "."))
        m.cuda().train()
    yield pkg, nets
    pkg.conv_blocks.set_precision("fp32")


def _run(pkg, module, fn, inputs, mode):
    """fn(module, *inputs) -> tuple of outputs; loss = sum_i <out_i, R_i>.  Returns outputs, input grads, param grads, buffers."""
    pkg.conv_blocks.set_precision(mode)
    state0 = {n: b.clone() for n, b in module.named_buffers()}
    ins = [x.clone().requires_grad_(True) if x.is_floating_point() else x for x in inputs]
    for p in module.parameters():
        p.grad = None
    outs = fn(module, *ins)
    outs = outs if isinstance(outs, tuple) else (outs,)
    g = torch.Generator(device="cuda").manual_seed(11)
    loss = sum((o.float() * torch.randn(o.shape, device="cuda", generator=g) * 1e-3).sum() for o in outs)
    loss.backward()
    res = ([o.detach().float() for o in outs], [x.grad for x in ins if x.is_floating_point()],
           {n: (p.grad.clone() if p.grad is not None else None) for n, p in module.named_parameters()},
           {n: b.clone() for n, b in module.named_buffers()})
    for n, b in module.named_buffers():
        b.copy_(state0[n])
    pkg.conv_blocks.set_precision("fp32")
    return res


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _compare(ref, got, what, in_tol=0.35, p_tol=0.4, cos_tol=0.95):
    """Collects every violation before failing, so that one GPU run shows the whole picture."""
    (o_r, gi_r, gp_r, buf_r), (o_k, gi_k, gp_k, buf_k) = ref, got
    bad, info = [], []
    for i, (a, b) in enumerate(zip(o_k, o_r)):
        assert a.shape == b.shape
        info.append(("out%d" % i, round(_rel(a, b), 4)))
        if _rel(a, b) > 2e-2:
            bad.append(info[-1])
    for i, (a, b) in enumerate(zip(gi_k, gi_r)):
        assert a is not None, "%s: missing input gradient" % what
        cos = float(torch.nn.functional.cosine_similarity(a.float().reshape(-1), b.reshape(-1), dim=0))
        info.append(("din%d" % i, round(_rel(a.float(), b), 4), round(cos, 4)))
        if _rel(a.float(), b) > in_tol:
            bad.append(info[-1])
    big = max(float(v.norm()) for v in gp_r.values() if v is not None)
    flat_r, flat_k = [], []
    for n, r in gp_r.items():
        k = gp_k[n]
        if r is not None and k is None:
            # a conv bias in front of a train-mode BatchNorm: identically zero gradient, none is produced
            assert float(r.norm()) < 1e-4 * big, (what, n, "no gradient produced but the reference has one")
            continue
        assert (r is None) == (k is None), (what, n)
        if r is None:
            continue
        assert k.shape == r.shape and k.dtype == torch.float32, (what, n, k.shape, k.dtype)
        flat_r.append(r.reshape(-1)); flat_k.append(k.reshape(-1))
        if float(r.norm()) < 1e-4 * big:
            if float(k.norm()) > 1e-3 * big:
                bad.append((n, "should be ~0", float(k.norm())))
        else:
            info.append((n, round(_rel(k, r), 3)))
            if _rel(k, r) > p_tol:
                bad.append(info[-1])
    cos = float(torch.nn.functional.cosine_similarity(torch.cat(flat_k), torch.cat(flat_r), dim=0)) if flat_r else 1.0
    info.append(("param cosine", round(cos, 4)))
    if cos < cos_tol:
        bad.append(info[-1])
    for n, b in buf_r.items():
        if n.endswith("num_batches_tracked"):
            if int(b) != int(buf_k[n]):
                bad.append((n, int(b), int(buf_k[n])))
        elif not torch.allclose(buf_k[n], b, rtol=3e-2, atol=1e-2):
            bad.append((n, "buffer", float((buf_k[n] - b).abs().max())))
    print(what, info)
    assert not bad, "%s: %s\nall metrics: %s" % (what, bad, info)


@pytest.mark.parametrize("name", ["segmentation_decoder", "image_decoder"])
def test_decoder_forward_backward(env, name):
    pkg, nets = env
    z = torch.relu(torch.randn(4, 128, 4, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)))
    fn = lambda m, zz: m(zz)
    _compare(_run(pkg, nets[name], fn, [z], "fp32"), _run(pkg, nets[name], fn, [z], "kernel"), name)


def test_dual_encoder_forward_backward(env):
    pkg, nets = env
    img, _, _ = weights.synthetic_batch(4, 64, 48, seed=3)
    fn = lambda m, x: m(x)
    enc = nets['image_encoder']
    _compare(_run(pkg, enc, fn, [img.cuda()], "fp32"), _run(pkg, enc, fn, [img.cuda()], "kernel"), "image_encoder")


@pytest.mark.parametrize("is_label", [False, True])
def test_shape_encoder_from_segmentation(env, is_label):
    pkg, nets = env
    enc = nets['shape_encoder']
    _, lab, _ = weights.synthetic_batch(3, 48, 64, seed=5)
    seg = lab.cuda() if is_label else torch.randn(3, 4, 48, 64, device="cuda") * 3

    def fn(m, s):
        code = m.forward_from_segmentation(s, is_label_map=is_label, temperature=2)
        if code is None:
            code = m(pkg.losses.construct_input(s, num_classes=4, apply_softmax=not is_label, is_labelmap=is_label,
                                                temperature=2))
        return code

    _compare(_run(pkg, enc, fn, [seg], "fp32"), _run(pkg, enc, fn, [seg], "kernel"), "shape_encoder")


def test_frozen_decoder_gives_only_the_latent_gradient(env):
    """The saliency pass: decoder parameters frozen, gradient w.r.t. the latent code only (model_util.py:212-223)."""
    pkg, nets = env
    dec = nets['image_decoder']
    z = torch.relu(torch.randn(4, 128, 4, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2)))
    pkg.model_util.set_grad(dec, False)
    try:
        fn = lambda m, zz: m(zz)
        ref = _run(pkg, dec, fn, [z], "fp32")
        got = _run(pkg, dec, fn, [z], "kernel")
    finally:
        pkg.model_util.set_grad(dec, True)
    assert all(v is None for v in got[2].values())
    assert _rel(got[1][0].float(), ref[1][0]) < 0.35, _rel(got[1][0].float(), ref[1][0])
    # ranking of the channel saliency (what the masks are built from) agrees
    s_r, s_k = ref[1][0].mean(dim=(2, 3)), got[1][0].float().mean(dim=(2, 3))
    assert float(torch.nn.functional.cosine_similarity(s_r.reshape(-1), s_k.reshape(-1), dim=0)) > 0.95


def test_untracked_bn_pass_leaves_buffers_and_affine_grads_alone(env):
    pkg, nets = env
    dec = nets['segmentation_decoder']
    z = torch.rand(4, 128, 4, 3, device="cuda")

    def fn(m, zz):
        with pkg.model_util._disable_tracking_bn_stats(m):
            return m(zz)

    before = {n: b.clone() for n, b in dec.named_buffers()}
    pkg.conv_blocks.set_precision("kernel")
    zz = z.clone().requires_grad_(True)
    out = fn(dec, zz)
    out.sum().backward()
    pkg.conv_blocks.set_precision("fp32")
    for n, b in dec.named_buffers():
        assert torch.equal(b, before[n]), n
    for n, p in dec.named_parameters():
        is_bn = p.dim() == 1 and (".conv.1." in n or ".conv.4." in n)                       # frozen inside the context
        bias_before_bn = n.endswith((".conv.0.bias", ".conv.3.bias"))                       # identically zero: not produced
        assert (p.grad is None) == (is_bn or bias_before_bn), n
        p.grad = None
    assert zz.grad is not None


def _solver_and_batch(pkg):
    torch.manual_seed(0)
    solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4, learning_rate=1e-4)
    for k, m in solver.model.items():
        m.load_state_dict(weights.synthetic_state_dict(m, 7, prefix=k + "."))
    img, lab, noise = weights.synthetic_batch(8, 64, 64, seed=3)
    return solver, img.cuda(), lab.cuda(), noise.cuda()


CFG_I = {"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.3, "random_threshold": False, "if_soft": False}
CFG_S = {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.3, "random_threshold": False, "if_soft": False}
MODULES = ('image_encoder', 'segmentation_decoder', 'shape_encoder', 'shape_decoder', 'image_decoder')


def _step(pkg, solver, state0, mode, img, lab, noise, latent_DA):
    pkg.conv_blocks.set_precision(mode)
    for k, m in solver.model.items():
        for n, b in m.named_buffers():
            b.copy_(state0[k][n])
    random.seed(1); np.random.seed(1); torch.manual_seed(1)
    r = pkg.cooperative_step(solver, img, lab, CFG_I, CFG_S, noise=noise, optimize=False, latent_DA=latent_DA)
    grads = {n: p.grad.clone() for n, p in solver.named_parameters() if p.grad is not None}
    return {k: float(v) for k, v in r.items() if k.startswith('loss')}, grads, r


@pytest.mark.parametrize("latent_DA", [False, True])
def test_cooperative_step_kernel_mode_tracks_fp32(env, latent_DA):
    """Whole step on the kernels vs the fp32 parity mode, with cuDNN's bf16 path ('bf16' mode) as the yardstick for what
    bf16 activations cost: losses within 3e-2 (clean pass) / 5e-2 (with hard examples: the hard top-30 % masks may pick a
    few different channels / positions) of fp32, and every sub-network's gradient at least as well aligned with fp32 as
    the library bf16 path's is, minus 0.1 (absolute floor 0.25 -- mask flips through ~60 layers, see the module header)."""
    pkg, _ = env
    solver, img, lab, noise = _solver_and_batch(pkg)
    state0 = {k: {n: b.clone() for n, b in m.named_buffers()} for k, m in solver.model.items()}
    try:
        la, ga, _ = _step(pkg, solver, state0, "fp32", img, lab, noise, latent_DA)
        lc, gc, _ = _step(pkg, solver, state0, "bf16", img, lab, noise, latent_DA)
        lb, gb, rb = _step(pkg, solver, state0, "kernel", img, lab, noise, latent_DA)
    finally:
        pkg.conv_blocks.set_precision("fp32")
    info, bad = [], []
    tol = 5e-2 if latent_DA else 3e-2
    for k in sorted(la):
        info.append((k, round(la[k], 5), round(lb[k], 5), round(lc[k], 5)))
        if abs(la[k] - lb[k]) > tol * abs(la[k]) + 1e-4:
            bad.append(info[-1])
    bias_before_bn = (".conv.0.bias", ".conv.3.bias", ".inc.0.bias", ".inc.3.bias", ".final_conv.0.bias",
                      ".code_decoupler.0.bias", ".code_decoupler.3.bias")
    extra = [n for n in set(ga) ^ set(gb) if not (n in ga and n.endswith(bias_before_bn))]
    if extra:
        bad.append(("gradient sets differ", sorted(extra)[:8]))

    def cosine(x, y, names):
        a = torch.cat([x[n].reshape(-1) for n in names])
        b = torch.cat([y[n].reshape(-1) for n in names])
        return float(torch.nn.functional.cosine_similarity(a, b, dim=0)), float(b.norm() / a.norm())

    for mod in MODULES:
        names = [n for n in ga if n.startswith(mod + '.') and n in gb and n in gc]
        ck, rk = cosine(ga, gb, names)
        cl, rl = cosine(ga, gc, names)
        info.append((mod, "cos kernel-vs-fp32", round(ck, 4), "cos cudnn_bf16-vs-fp32", round(cl, 4), "norm ratios",
                     round(rk, 3), round(rl, 3)))
        if not (ck > max(0.25, cl - 0.1)) or not (0.7 < rk < 1.4):
            bad.append(info[-1])
    print(info)
    assert not bad, "%s\nall: %s" % (bad, info)
    if latent_DA:
        assert rb['perturbed_image'].requires_grad is False and rb['perturbed_seg'].requires_grad is False


def test_kernel_mode_trains_like_the_library_path(env):
    """Ten optimizer steps on a fixed batch: the loss must FALL as it does with cuDNN's bf16 path (a stale packed
    weight or a wrong-signed gradient shows up here, not in single-step comparisons)."""
    pkg, _ = env
    curves = {}
    for mode in ("bf16", "kernel"):
        try:
            pkg.conv_blocks.set_precision(mode)
            solver, img, lab, noise = _solver_and_batch(pkg)
            random.seed(3); np.random.seed(3); torch.manual_seed(3)
            curves[mode] = [float(pkg.cooperative_step(solver, img, lab, CFG_I, CFG_S, noise=noise)['loss'])
                            for _ in range(10)]
        finally:
            pkg.conv_blocks.set_precision("fp32")
    print(curves)
    drop_lib = curves["bf16"][0] - curves["bf16"][-1]
    drop_ker = curves["kernel"][0] - curves["kernel"][-1]
    assert drop_lib > 0.5, curves
    assert drop_ker > 0.7 * drop_lib, curves
    assert abs(curves["kernel"][-1] - curves["bf16"][-1]) < 0.08 * curves["bf16"][-1], curves


def test_predict_on_a_10_slice_stack_matches_fp32(env):
    """BASELINE.json configs[4]: inference-only FTN + STN refinement on a 10 x 256 x 256 stack (predict, n_iter=2) --
    kernel mode runs the BN-folded all-kernel forward (fastpath 'eval'); the label map must agree with the fp32 modules."""
    pkg, _ = env
    solver, _, _, _ = _solver_and_batch(pkg)
    img, _, _ = weights.synthetic_batch(10, 256, 256, seed=9)
    img = img.cuda()
    try:
        pkg.conv_blocks.set_precision("fp32")
        want = solver.predict(img, softmax=True, n_iter=2)
        pkg.conv_blocks.set_precision("kernel")
        n0 = pkg._lib.LAUNCHES["count"]
        got = solver.predict(img, softmax=True, n_iter=2)
        launched = pkg._lib.LAUNCHES["count"] - n0
    finally:
        pkg.conv_blocks.set_precision("fp32")
    assert launched > 50, "predict did not run on the kernels"
    assert got.shape == want.shape == (10, 4, 256, 256) and got.dtype == torch.float32
    agree = float((got.argmax(1) == want.argmax(1)).float().mean())
    assert agree > 0.9, agree            # synthetic random weights: many pixels sit on near-ties between classes
    assert float((got - want).abs().mean()) < 3e-2   # bf16 through FTN + 2x STN on random weights (measured 2.1e-2)
    assert solver.training is False


@pytest.mark.parametrize("name", ["segmentation_decoder", "image_decoder", "image_encoder"])
def test_direct_grad_accumulation_equals_autograd_accumulation(env, name):
    """trainpath.accumulate_into_grads(): two backward passes ADDED by the kernels into pre-existing .grad tensors give
    what autograd's per-parameter accumulation gives (fp32 atomics in a different order: 1e-5 of the largest entry)."""
    pkg, nets = env
    from cooperative_training_and_latent_space_data_augmentation_b200 import trainpath
    net = nets[name]
    gen = torch.Generator(device="cuda").manual_seed(4)
    if name == "image_encoder":
        x = torch.rand(4, 1, 64, 48, device="cuda", generator=gen)
    else:
        x = torch.relu(torch.randn(4, 128, 4, 3, device="cuda", generator=gen))
    pkg.conv_blocks.set_precision("kernel")
    state0 = {n: b.clone() for n, b in net.named_buffers()}

    def two_passes(direct):
        for n, b in net.named_buffers():
            b.copy_(state0[n])
        for p in net.parameters():
            p.grad = torch.zeros_like(p) if direct else None
        for rep in range(2):
            outs = net(x * (1.0 + 0.1 * rep))
            outs = outs if isinstance(outs, tuple) else (outs,)
            g = torch.Generator(device="cuda").manual_seed(11 + rep)
            loss = sum((o.float() * torch.randn(o.shape, device="cuda", generator=g) * 1e-3).sum() for o in outs)
            if direct:
                with trainpath.accumulate_into_grads():
                    loss.backward()
            else:
                loss.backward()
        return {n: (p.grad.clone() if p.grad is not None else None) for n, p in net.named_parameters()}

    try:
        ref, got = two_passes(False), two_passes(True)
    finally:
        pkg.conv_blocks.set_precision("fp32")
        for n, b in net.named_buffers():
            b.copy_(state0[n])
        for p in net.parameters():
            p.grad = None
    for n, r in ref.items():
        k = got[n]
        if r is None:                      # a conv bias in front of a train-mode BatchNorm: no gradient is produced
            assert float(k.abs().max()) == 0.0, n
            continue
        assert float((k - r).abs().max()) <= 1e-5 * float(r.abs().max()) + 1e-12, (n, float((k - r).abs().max()))
