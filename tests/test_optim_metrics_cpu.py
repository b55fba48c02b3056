"""CPU: (1) the metric / Adam oracle (oracle/metrics_oracle.py) pinned on fixtures produced by the UNMODIFIED reference's
runningScore (tests/golden/metrics_scores.npz, oracle/make_golden.py gen_metrics) and on torch.optim.Adam itself;
(2) the host logic of optim.FlatAdam: layout of the flat buffers, parameters / gradients re-pointed as views,
torch-format optimizer state round trip -- everything except the kernel launch, which needs a GPU
(tests/test_optim_metrics_gpu.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import metrics_oracle as mx

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_running_score_oracle_matches_reference_fixture(case):
    f = np.load(os.path.join(GOLDEN, "metrics_scores.npz"))
    n = int(f[case + "_n"])
    m = mx.RunningScoreOracle(n)
    for gt, pr in zip(f[case + "_gt"], f[case + "_pred"]):
        m.update(gt.astype(np.int64), pr.astype(np.int64))
    np.testing.assert_array_equal(m.confusion_matrix, f[case + "_hist"])
    scores, iu = m.get_scores()
    np.testing.assert_array_equal(scores, f[case + "_scores"])          # same arithmetic, same order: bit-exact
    np.testing.assert_array_equal(iu, f[case + "_cls_iu"])
    assert [k.strip() for k in f[case + "_score_keys"]] == ['Overall Acc:', 'Mean Acc :', 'FreqW Acc :', 'Mean IoU :']


def test_adam_oracle_matches_torch_adam():
    torch.manual_seed(0)
    p = torch.randn(257, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([p], lr=1e-3)
    pn, m, v = p.detach().numpy().copy(), np.zeros(257), np.zeros(257)
    for t in range(1, 6):
        g = torch.randn(257, dtype=torch.float64)
        p.grad = g.clone()
        opt.step()
        pn, m, v = mx.adam_step(pn, g.numpy(), m, v, t, lr=1e-3)
        np.testing.assert_allclose(pn, p.detach().numpy(), rtol=1e-12, atol=1e-14)


def _toy():
    torch.manual_seed(3)
    a = nn.Sequential(nn.Conv2d(1, 3, 3), nn.BatchNorm2d(3), nn.Conv2d(3, 5, 1))
    b = nn.Sequential(nn.Linear(7, 2))
    return {"a": a, "b": b}


def test_flat_adam_layout_and_views():
    from cooperative_training_and_latent_space_data_augmentation_b200.optim import FlatAdam
    mods = _toy()
    before = {k: {n: t.clone() for n, t in m.state_dict().items()} for k, m in mods.items()}
    flat = FlatAdam(mods, lr=1e-3)
    assert flat.attached()
    # values and state_dict keys untouched; parameters / gradients are views of the flat buffers
    for k, m in mods.items():
        for n, t in m.state_dict().items():
            assert torch.equal(t, before[k][n]), (k, n)
    covered = torch.zeros(flat.numel, dtype=torch.bool)
    for p, o in zip(flat.params, flat.offsets):
        assert p.data_ptr() == flat.flat_params.data_ptr() + 4 * o
        assert p.grad is not None and p.grad.data_ptr() == flat.flat_grads.data_ptr() + 4 * o and p.grad.shape == p.shape
        if p.dim() > 1:
            assert o % 4 == 0                       # weight tensors on 16-byte boundaries
        assert not covered[o:o + p.numel()].any()
        covered[o:o + p.numel()] = True
    assert float(flat.flat_params[~covered].abs().sum()) == 0.0
    for (b, e), name in zip(flat.bounds, flat.names):
        assert b % 64 == 0 and e - b >= sum(p.numel() for p in mods[name].parameters())
    # autograd accumulates INTO the views; zero_grad keeps them attached
    x = torch.randn(2, 1, 8, 8)
    mods["a"](x).sum().backward()
    assert flat.attached() and float(flat.flat_grads.abs().sum()) > 0
    g0 = flat.flat_grads.clone()
    flat.view("a").zero_grad()
    b0, e0 = flat.bounds[0]
    assert float(flat.flat_grads[b0:e0].abs().sum()) == 0.0 and torch.equal(flat.flat_grads[e0:], g0[e0:])
    # something replaced a gradient (zero_grad(set_to_none=True) of a foreign caller): reattach folds it back
    mods["b"][0].weight.grad = None
    assert not flat.attached()
    flat.reattach()
    assert flat.attached()
    # writes through the flat buffer are writes to the module
    flat.flat_params[flat.offsets[0]] = 42.0
    assert float(mods["a"][0].weight.view(-1)[0]) == 42.0


def test_flat_adam_state_dict_round_trip_in_torch_format():
    from cooperative_training_and_latent_space_data_augmentation_b200.optim import FlatAdam
    mods = _toy()
    # a torch.optim.Adam state as a reference checkpoint holds it (advanced...model.py:676-677)
    ref_opt = torch.optim.Adam(mods["a"].parameters(), lr=3e-4)
    mods["a"](torch.randn(2, 1, 8, 8)).sum().backward()
    ref_opt.step()
    ref_opt.step()
    sd = ref_opt.state_dict()
    flat = FlatAdam(mods, lr=1e-4)
    view = flat.view("a")
    view.load_state_dict(sd)
    assert float(flat.steps[0]) == 2.0 and float(flat.steps[1]) == 0.0
    assert view.param_groups[0]["lr"] == 3e-4
    out = view.state_dict()
    assert sorted(out["state"]) == sorted(sd["state"])
    for i, st in sd["state"].items():
        assert torch.equal(out["state"][i]["exp_avg"], st["exp_avg"])
        assert torch.equal(out["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(out["state"][i]["step"]) == float(st["step"])
    # and torch.optim.Adam accepts what the view writes
    fresh = torch.optim.Adam(mods["a"].parameters(), lr=1.0)
    fresh.load_state_dict(out)
    assert fresh.param_groups[0]["lr"] == 3e-4
    # a never-stepped sub-network has an empty state, like a fresh torch optimizer
    assert flat.view("b").state_dict()["state"] == {}
    # loading happens IN PLACE: the buffers a captured CUDA graph would update are still the same memory
    ptr = flat.exp_avg.data_ptr()
    view.load_state_dict(out)
    assert flat.exp_avg.data_ptr() == ptr


def test_confusion_and_adam_entry_points_validate_on_the_host():
    import ctypes
    from cooperative_training_and_latent_space_data_augmentation_b200 import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(256)
    p = ctypes.cast(buf, ctypes.c_void_p)
    bounds = (ctypes.c_int64 * 4)(0, 10, 8, 20)              # second segment starts inside the first
    assert lib.ctl_adam_flat(p, p, p, p, bounds, 2, 3, p, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, 0, None) == _lib.CTL_ERR_INVALID
    bounds = (ctypes.c_int64 * 2)(2, 10)                     # not a multiple of 4
    assert lib.ctl_adam_flat(p, p, p, p, bounds, 1, 1, p, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, 0, None) == _lib.CTL_ERR_INVALID
    assert lib.ctl_adam_flat(p, p, p, p, bounds, 1, 1, p, 1e-3, 1.5, 0.999, 1e-8, 0.0, 1.0, 0, None) == _lib.CTL_ERR_INVALID
    assert lib.ctl_confusion_update(p, p, p, 1, 4, 16, p, None, None) == _lib.CTL_ERR_INVALID      # logits AND labels
    assert lib.ctl_confusion_update(p, None, p, 1, 4, 16, None, None, None) == _lib.CTL_ERR_INVALID  # gt without hist
    assert lib.ctl_sse_fwd(p, p, 0, 1.0, p, p, None) == _lib.CTL_ERR_INVALID
