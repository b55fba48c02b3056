"""GPU: the CUDA-graph replay of the cooperative step (training.GraphedCooperativeTrainer) against the same step issued
launch by launch (training.CooperativeTrainer).

Two independent training runs cannot be compared tightly: the weight-gradient kernels accumulate with fp32 atomics
(order varies run to run), Adam turns that noise into +-lr on near-zero gradients, and a near-tie of a later step's
top-k selection then flips a mask entry -- two EAGER runs fork the same way (tools/graph_divergence.py,
profiles/r1_graph_divergence_session8.txt).  So every step is compared FROM IDENTICAL STATE: before each step the
follower's parameters, Adam moments / step counters and BatchNorm buffers are overwritten with the leader's and the
host / Philox generator states are aligned; what differs is only how the step is issued.  Then

  * the host generators (python `random`, numpy) are consumed identically -- their states after the step are EQUAL
  * k, the Philox offset and the first-sample index reach the kernels through device memory: masks and perturbed
    examples of a replayed step equal the eager ones (forward and the saliency pass have no order-dependent
    arithmetic beyond fp64 atomics rounded once to fp32): bar 1e-3 relative per sample, losses 1e-4
  * parameters after the step agree to the atomics' noise (1e-3 of the largest update)
"""
import random

import numpy as np
import pytest
import torch

from oracle import weights

pytestmark = pytest.mark.gpu

RANDOM_CFG = ({"loss_name": "mse", "mask_type": "random", "max_threshold": 0.5, "random_threshold": True,
               "if_soft": True},
              {"loss_name": "ce", "mask_type": "random", "max_threshold": 0.5, "random_threshold": True,
               "if_soft": True})
FIXED_CFG = ({"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.5, "random_threshold": True,
              "if_soft": True},
             {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.5, "random_threshold": True,
              "if_soft": True})


def _make(pkg, trainer_cls, cfgs, **kw):
    torch.manual_seed(0)
    solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4, learning_rate=1e-5)
    for name, m in solver.model.items():
        m.load_state_dict(weights.synthetic_state_dict(m, 5, prefix=name + "."))
    return trainer_cls(solver, 4, seed=3, image_cfg=cfgs[0], seg_cfg=cfgs[1], **kw)


def _copy_state(dst, src):
    """parameters, optimizer state and BatchNorm buffers of `src`'s solver into `dst`'s (in place)."""
    fa, fb = src.solver.flat_adam, dst.solver.flat_adam
    with torch.no_grad():
        for name in ("flat_params", "exp_avg", "exp_avg_sq", "steps"):
            getattr(fb, name).copy_(getattr(fa, name))
        for k in src.solver.model:
            for bs, bd in zip(src.solver.model[k].buffers(), dst.solver.model[k].buffers()):
                bd.copy_(bs)
    from cooperative_training_and_latent_space_data_augmentation_b200 import fastpath
    fastpath.weights_changed()


def _lockstep(pkg, leader, follower, steps, prefetch=False):
    """Runs `steps` steps; before each one the follower is put into the leader's state.  Returns per-step records."""
    img, lab, noise = weights.synthetic_batch(4, 64, 64, seed=2)
    noise = noise.cuda()
    img_f, lab_f = (img.pin_memory(), lab.pin_memory()) if prefetch else (img.cuda(), lab.cuda())
    img, lab = img.cuda(), lab.cuda()
    rng = pkg.model_util.native_rng()
    records = []
    for step in range(steps):
        torch.cuda.synchronize()
        _copy_state(follower, leader)
        host0, off0 = (random.getstate(), np.random.get_state()), rng.offset
        a = leader.step(img, lab, noise)
        la = {k: float(v) for k, v in a.items() if k.startswith('loss')}
        pa = (a['perturbed_image'].float().clone(), a['perturbed_seg'].float().clone())
        wa = leader.solver.flat_adam.flat_params.clone()
        host_a, off_a = (random.getstate(), np.random.get_state()), rng.offset
        random.setstate(host0[0]); np.random.set_state(host0[1]); rng.offset = off0
        b = follower.step(img_f, lab_f, noise)
        if prefetch:
            follower.prefetch(img_f, lab_f)     # next step's H2D copy on the side stream, consumed by the next step()
        lb = {k: float(v) for k, v in b.items() if k.startswith('loss')}
        pb = (b['perturbed_image'].float().clone(), b['perturbed_seg'].float().clone())
        torch.cuda.synchronize()
        wb = follower.solver.flat_adam.flat_params.clone()
        host_b, off_b = (random.getstate(), np.random.get_state()), rng.offset
        records.append(dict(la=la, lb=lb, pa=pa, pb=pb, wa=wa, wb=wb, host_a=host_a, host_b=host_b, off=(off_a, off_b)))
    return records


def _check(records, lr=1e-5):
    for step, r in enumerate(records):
        assert r["host_a"][0] == r["host_b"][0], step                                  # python generator
        assert all(np.array_equal(x, y) for x, y in zip(r["host_a"][1], r["host_b"][1])), step   # numpy generator
        assert r["off"][0] == r["off"][1], (step, r["off"])                            # Philox draw counter
        for key, want in r["la"].items():
            assert np.isfinite(r["lb"][key])
            assert abs(r["lb"][key] - want) <= 1e-4 * max(1.0, abs(want)), (step, key, r["lb"][key], want)
        for a, b in zip(r["pa"], r["pb"]):
            d = (a - b).flatten(1).norm(dim=1) / a.flatten(1).norm(dim=1).clamp_min(1e-6)
            assert float(d.max()) < 1e-3, (step, d.tolist())
        # Adam moves a weight by at most ~lr per step; the two issue modes agree to the atomics' noise
        assert float((r["wa"] - r["wb"]).abs().max()) <= 2.5 * lr, (step, float((r["wa"] - r["wb"]).abs().max()))
        assert float((r["wa"] - r["wb"]).abs().mean()) <= 2e-2 * lr, (step, float((r["wa"] - r["wb"]).abs().mean()))


@pytest.fixture()
def pkg():
    import cooperative_training_and_latent_space_data_augmentation_b200 as pkg
    pkg.conv_blocks.set_precision("kernel")
    yield pkg
    pkg.conv_blocks.set_precision("fp32")
    pkg.set_rng_mode("torch")


@pytest.mark.parametrize("cfgs", [FIXED_CFG, RANDOM_CFG], ids=["channel+spatial", "random"])
def test_graph_replay_matches_eager_steps(pkg, cfgs):
    steps = 8
    leader = _make(pkg, pkg.CooperativeTrainer, cfgs)
    follower = _make(pkg, pkg.GraphedCooperativeTrainer, cfgs, eager_steps=2)
    n0 = pkg._lib.LAUNCHES["count"]
    records = _lockstep(pkg, leader, follower, steps)
    assert pkg._lib.LAUNCHES["count"] - n0 > 1000 * steps, "replays did not account their kernels"
    assert len(follower.captured) >= 1
    if cfgs is RANDOM_CFG:
        assert len(follower.captured) >= 2, "the random mask type should have met several type combinations"
    _check(records)
    assert follower.solver.flat_adam.steps.tolist() == [float(steps)] * 5


def test_prefetched_inputs_give_the_same_steps(pkg):
    """GraphedCooperativeTrainer.prefetch: the batch copied host -> device on the side stream under the previous step is
    the batch the next step trains on."""
    steps = 5
    leader = _make(pkg, pkg.GraphedCooperativeTrainer, FIXED_CFG, eager_steps=2)
    follower = _make(pkg, pkg.GraphedCooperativeTrainer, FIXED_CFG, eager_steps=2)
    records = _lockstep(pkg, leader, follower, steps, prefetch=True)
    assert follower._prefetched is not None and follower._stage is not None      # the last prefetch is pending
    _check(records)


def test_step_params_reach_the_kernels(pkg):
    """ctl_saliency_mask_apply_dyn / ctl_channel_dropout_dyn read (k, offset, first sample) from device memory:
    bit-exact against the by-value entry points for several parameter sets on ONE recorded row."""
    ops = pkg.ops
    gen = torch.Generator(device="cuda").manual_seed(1)
    z = torch.relu(torch.randn(6, 32, 7, 9, device="cuda", generator=gen))
    g = 1e-5 * torch.randn(6, 32, 7, 9, device="cuda", generator=gen)
    for mode, n in ((ops.MODE_CHANNEL, 32), (ops.MODE_SPATIAL, 63)):
        sp = ops.StepParams("cuda")
        rng = ops.NativeRNG(seed=9, first_sample=12)
        ops.saliency_mask_apply(g, z, mode, 3, soft=True, rng=rng, step_params=sp)
        for k, offset, first in ((3, 0, 12), (0, 5, 0), (n - 1, 2, 40)):
            r2 = ops.NativeRNG(seed=9, first_sample=first)
            r2.offset = offset
            want = ops.saliency_mask_apply(g, z, mode, k, soft=True, rng=r2)
            r3 = ops.NativeRNG(seed=9, first_sample=first)
            r3.offset = offset
            sp.begin()
            sp.fill(0, k, r3)
            sp.upload()
            # same recorded launch, new device-resident values
            row = sp.dev[0]
            s = torch.empty_like(want[2]); zt = torch.empty_like(want[0]); mask = torch.empty_like(want[1])
            N, C, H, W = z.shape
            pkg._lib.check(pkg._lib.load().ctl_saliency_mask_apply_dyn(
                g.data_ptr(), 0, z.data_ptr(), 0, N, C, H * W, mode, 1, 0, 9, row.data_ptr(), s.data_ptr(),
                mask.data_ptr(), 0, zt.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
            assert torch.equal(zt, want[0]) and torch.equal(mask, want[1]) and torch.equal(s, want[2])
    sp = ops.StepParams("cuda")
    rng = ops.NativeRNG(seed=4, first_sample=3)
    rng.offset = 7
    got = ops.channel_dropout(z, 0.5, rng=rng, want_keep=True, step_params=sp)
    r2 = ops.NativeRNG(seed=4, first_sample=3)
    r2.offset = 7
    want = ops.channel_dropout(z, 0.5, rng=r2, want_keep=True)
    assert all(torch.equal(a, b) for a, b in zip(got, want))
    with pytest.raises(IndexError):
        ops.saliency_mask_apply(g, z, ops.MODE_CHANNEL, 32, soft=False, step_params=ops.StepParams("cuda"))
