"""GPU, through the C ABI: the multi-tensor Adam (ctl_adam_flat / optim.FlatAdam), the fused squared-error loss, the
on-device confusion matrix / scores and the graph-replayed inference engine, each against its oracle:
torch.optim.Adam (the reference's own dependency) and oracle/metrics_oracle.py (pinned on reference fixtures).
Tolerances: Adam 2e-6 relative on the parameters after 5 steps (fp32 arithmetic, pow in fp64 vs torch's);
squared error 1e-6; confusion matrices bit-exact (integers); scores 1e-12 (fp64, same formula)."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as mx
from oracle import weights

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture()
def pkg():
    import cooperative_training_and_latent_space_data_augmentation_b200 as p
    return p


def test_adam_flat_matches_torch_adam(pkg):
    gen = torch.Generator(device="cuda").manual_seed(0)
    bounds = [(0, 1001), (1024, 1024 + 4099), (5184, 5184 + 7)]
    n = 5192
    p = torch.randn(n, device="cuda", generator=gen)
    p_init = p.clone()
    ref_p = [p[b:e].clone().requires_grad_(True) for b, e in bounds]
    opts = [torch.optim.Adam([q], lr=1e-3, foreach=False, fused=False) for q in ref_p]
    g = torch.zeros(n, device="cuda")
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    steps = torch.zeros(3, device="cuda")
    world = 4.0
    for it in range(5):
        mask = 0b111 if it != 2 else 0b101                   # step 2: the middle optimizer sits out
        for s, (b, e) in enumerate(bounds):
            g[b:e] = torch.randn(e - b, device="cuda", generator=gen) * (10.0 ** (s - 2))
            if (mask >> s) & 1:
                ref_p[s].grad = (g[b:e] / world).clone()
                opts[s].step()
        g_before = g.clone()
        pkg.ops.adam_flat(p, g, m, v, bounds, steps, 1e-3, grad_scale=1.0 / world, zero_grad=(it % 2 == 1), seg_mask=mask)
        for s, (b, e) in enumerate(bounds):
            if (mask >> s) & 1 and it % 2 == 1:
                assert float(g[b:e].abs().sum()) == 0.0      # cleared in the same pass
            else:
                assert torch.equal(g[b:e], g_before[b:e])
    assert steps.tolist() == [5.0, 4.0, 5.0]
    for s, (b, e) in enumerate(bounds):
        np.testing.assert_allclose(p[b:e].cpu().numpy(), ref_p[s].detach().cpu().numpy(), rtol=2e-6, atol=1e-7)
        st = opts[s].state[ref_p[s]]
        want_m, want_v = st["exp_avg"].cpu().numpy(), st["exp_avg_sq"].cpu().numpy()
        np.testing.assert_allclose(m[b:e].cpu().numpy(), want_m, rtol=1e-5, atol=1e-6 * np.abs(want_m).max())
        np.testing.assert_allclose(v[b:e].cpu().numpy(), want_v, rtol=1e-5, atol=1e-6 * np.abs(want_v).max())
    # untouched padding between the segments
    assert torch.equal(p[1001:1024], p_init[1001:1024]) and float(m[1001:1024].abs().sum()) == 0.0
    # fp64 oracle on one segment, one step from scratch
    p1 = torch.randn(64, device="cuda", generator=gen); g1 = torch.randn(64, device="cuda", generator=gen)
    want, _, _ = mx.adam_step(p1.cpu().numpy(), g1.cpu().numpy(), np.zeros(64), np.zeros(64), 1, lr=1e-4)
    pkg.ops.adam_flat(p1, g1, torch.zeros(64, device="cuda"), torch.zeros(64, device="cuda"), [(0, 64)],
                      torch.zeros(1, device="cuda"), 1e-4)
    np.testing.assert_allclose(p1.cpu().numpy(), want, rtol=1e-6, atol=1e-8)


def test_solver_optimizers_step_like_five_torch_adams(pkg):
    """solver.optimize_all_params() (one launch) and solver.optimize_params(name) against torch.optim.Adam per
    sub-network (advanced...model.py:774-789); the checkpoint the views write loads into torch.optim.Adam."""
    solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4, learning_rate=1e-4)
    assert set(solver.optimizers) == set(solver.model) and solver.flat_adam.attached()
    assert sum(p.numel() for p in solver.parameters()) == 2528953
    clones = {k: [p.detach().clone().requires_grad_(True) for p in m.parameters()] for k, m in solver.model.items()}
    refs = {k: torch.optim.Adam(ps, lr=1e-4, foreach=True) for k, ps in clones.items()}
    gen = torch.Generator(device="cuda").manual_seed(1)
    for it in range(3):
        solver.reset_all_optimizers()
        assert float(solver.flat_adam.flat_grads.abs().sum()) == 0.0
        for k, m in solver.model.items():
            for p, c in zip(m.parameters(), clones[k]):
                gr = torch.randn(p.shape, device="cuda", generator=gen) * 1e-3
                p.grad.add_(gr)                              # accumulate into the flat view, as backward does
                c.grad = gr.clone()
        if it < 2:
            solver.optimize_all_params()
            for o in refs.values():
                o.step()
        else:
            solver.optimize_params('shape_encoder')
            refs['shape_encoder'].step()
    for k, m in solver.model.items():
        for (n, p), c in zip(m.named_parameters(), clones[k]):
            np.testing.assert_allclose(p.detach().cpu().numpy(), c.detach().cpu().numpy(), rtol=3e-6, atol=1e-8,
                                       err_msg=k + "." + n)
    sd = solver.optimizers['shape_encoder'].state_dict()
    assert float(sd["state"][0]["step"]) == 3.0
    fresh = torch.optim.Adam(clones['shape_encoder'], lr=1.0)
    fresh.load_state_dict(sd)
    # snapshot round trip through the reference-shaped methods
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = solver.save_snapshots(d, epoch=7)
        other = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4, learning_rate=1e-4)
        assert other.load_snapshots(path) == 7
        assert torch.equal(other.flat_adam.exp_avg, solver.flat_adam.exp_avg)
        assert torch.equal(other.flat_adam.steps, solver.flat_adam.steps)
        assert torch.equal(other.flat_adam.flat_params, solver.flat_adam.flat_params) and other.flat_adam.attached()


def test_squared_error_matches_torch(pkg):
    gen = torch.Generator(device="cuda").manual_seed(2)
    for shape in ((64, 1, 224, 224), (3, 1, 7, 9)):
        a = torch.rand(shape, device="cuda", generator=gen).requires_grad_(True)
        b = torch.rand(shape, device="cuda", generator=gen)
        want = 0.5 * torch.nn.functional.mse_loss(a, b)
        (gw,) = torch.autograd.grad(want * 3.0, a)
        a2 = a.detach().clone().requires_grad_(True)
        got = pkg.losses.half_mse_loss(a2, b)
        (gg,) = torch.autograd.grad(got * 3.0, a2)
        np.testing.assert_allclose(float(got), float(want), rtol=1e-6)
        np.testing.assert_allclose(gg.cpu().numpy(), gw.cpu().numpy(), rtol=1e-5, atol=1e-12)
    # the workspace is left clean: a second call gives the same value
    assert float(pkg.losses.half_mse_loss(a2, b)) == float(got)


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_confusion_matrix_and_scores_match_reference_fixture(pkg, case):
    f = np.load(os.path.join(GOLDEN, "metrics_scores.npz"))
    n = int(f[case + "_n"])
    m = pkg.runningScore(n)
    for gt, pr in zip(f[case + "_gt"], f[case + "_pred"]):
        gt = gt.astype(np.int64)
        gt[gt == 255] = 255                                  # out-of-range true labels stay out of range
        m.update(gt, pr.astype(np.int64))                    # numpy in, like the reference's callers
    np.testing.assert_array_equal(m.confusion_matrix, f[case + "_hist"])
    scores, cls_iu = m.get_scores()
    assert list(scores.keys()) == list(f[case + "_score_keys"])
    np.testing.assert_allclose(np.array(list(scores.values())), f[case + "_scores"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(np.array([cls_iu[i] for i in range(n)]), f[case + "_cls_iu"], rtol=1e-12, equal_nan=True)
    m.reset()
    assert m.confusion_matrix.sum() == 0


def test_confusion_from_logits_fuses_the_argmax(pkg):
    gen = torch.Generator(device="cuda").manual_seed(3)
    logits = torch.randn(10, 4, 256, 256, device="cuda", generator=gen)
    logits[0, :, :4, :4] = 1.0                                # exact ties: first maximum wins, like torch.max(1)[1]
    gt = torch.randint(0, 4, (10, 256, 256), device="cuda", generator=gen)
    gt[1, :3] = -1
    m = pkg.runningScore(4)
    labels = m.update_from_logits(gt, logits, want_labels=True)
    want_labels = logits.max(1)[1]
    assert labels.dtype == torch.uint8 and torch.equal(labels.long(), want_labels)
    ora = mx.RunningScoreOracle(4)
    ora.update(gt.cpu().numpy(), want_labels.cpu().numpy())
    np.testing.assert_array_equal(m.confusion_matrix, ora.confusion_matrix)
    s, iu = ora.get_scores()
    got, _ = m.get_scores()
    np.testing.assert_allclose(np.array(list(got.values())), s, rtol=1e-12)
    assert torch.equal(pkg.ops.argmax_labels(logits).long(), want_labels)


def test_graphed_predictor_equals_eager_predict(pkg):
    """BASELINE.json configs[4]: 10 x 256 x 256 stacks.  The graph-replayed chunk (frozen BN affines, fused arg-max +
    confusion matrix) must give exactly what solver.predict + arg-max give eagerly on the same kernels."""
    pkg.conv_blocks.set_precision("kernel")
    solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4)
    for k, mod in solver.model.items():
        mod.load_state_dict(weights.synthetic_state_dict(mod, 7, prefix=k + "."))
    pred = pkg.GraphedPredictor(solver, (10, 1, 256, 256), n_iter=2)
    ora = mx.RunningScoreOracle(4)
    for seed in (1, 2):
        img, lab, _ = weights.synthetic_batch(23, 256, 256, seed=seed)     # 10 + 10 + 3 slices: a tail chunk too
        out = pred.predict_stack(img.pin_memory(), lab.cuda())
        torch.cuda.synchronize()
        want = solver.predict(img.cuda(), n_iter=2).max(1)[1]
        assert out.shape == (23, 256, 256) and out.dtype == torch.uint8
        assert torch.equal(out.long(), want.cpu())
        ora.update(lab.numpy(), want.cpu().numpy())
    np.testing.assert_array_equal(pred.metric.confusion_matrix, ora.confusion_matrix)
    got, _ = pred.scores()
    np.testing.assert_allclose(np.array(list(got.values())), ora.get_scores()[0], rtol=1e-12)
    # solver.evaluate: the reference-shaped entry (advanced...model.py:643-664) with the metric on the device
    solver.running_metric = solver.set_running_metric()
    img, lab, _ = weights.synthetic_batch(4, 64, 64, seed=5)
    logits = solver.evaluate(img.cuda(), lab.numpy(), n_iter=2)
    assert torch.equal(solver.cur_eval_predicts.long(), logits.max(1)[1])
    assert solver.running_metric.confusion_matrix.sum() == 4 * 64 * 64
